"""cpvs_b200 -- host-side mirror of the reference's C++ API over the CUDA library's C ABI.

The product is ``libcpvs_b200.so`` (hand-written sm_100a kernels, ``cpvs_b200/csrc``) behind
``include/cpvs_b200.h``. This module only binds that ABI with ctypes and puts the reference's names
back on top -- ``MinMaxHierarchy``, ``CompressedShadow.create`` / ``traverse`` / ``getDAG``,
``CompressedShadowContainer.set`` / ``copyToGPU`` / ``evaluate`` (reference ``src/MinMaxHierarchy.h``,
``src/CompressedShadow.h``, ``src/CompressedShadowContainer.h``) -- so tests and benches read like the
reference's own. C++ callers use ``include/cpvs/*.h`` instead.

There is no CPU path: if the library is missing or no B200 is present, calls raise.
"""
import ctypes
import os

import numpy as np

# See capi.cu (moreHardwareQueues): the library's streams should not share hardware queues. Read by the driver when the
# process's CUDA context is created, so it is set here, at import, and again when the library is loaded.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcpvs_b200.so")

OK, EINVAL, ENOMEM, ECUDA, EOVERFLOW, EINTERNAL = range(6)
MEM_HOST, MEM_DEVICE = 0, 1
SHADOW, VISIBLE, PARTIAL = 0, 1, 2
MAX_LEVELS = 32
NUM_PHASES = 10
PHASE_NAMES = ["count", "expand", "leaves", "leaf_table", "leaf_insert", "leaf_resolve", "inner_merge", "bases", "emit_inner", "emit_leaves"]
GRID_CELL_SHADOWED = 0x0FFFFFFF
GRID_CELL_VISIBLE = 0x0FFFFFFE


class CpvsError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("cpvs_b200 error %d: %s" % (code, message))
        self.code = code


class CtxStats(ctypes.Structure):
    _fields_ = [("predicted_builds", ctypes.c_uint64), ("exact_builds", ctypes.c_uint64), ("overflow_rebuilds", ctypes.c_uint64),
                ("reemissions", ctypes.c_uint64)]


class ShadowInfo(ctypes.Structure):
    _fields_ = [("num_levels", ctypes.c_uint32), ("leafmasks", ctypes.c_uint32), ("total_visibility", ctypes.c_uint32),
                ("predicted", ctypes.c_uint32), ("words", ctypes.c_uint64), ("svo_nodes", ctypes.c_uint64 * MAX_LEVELS),
                ("dag_nodes", ctypes.c_uint64 * MAX_LEVELS), ("dag_words", ctypes.c_uint64 * MAX_LEVELS),
                ("build_ms", ctypes.c_float), ("phase_ms", ctypes.c_float * NUM_PHASES)]


# name -> (restype, argtypes); also the list tests check against include/cpvs_b200.h
_VP, _I, _U32, _U64, _I64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int64
_PP = ctypes.POINTER(ctypes.c_void_p)
SIGNATURES = {
    "cpvs_ctx_create": (_I, [_I, _PP]),
    "cpvs_ctx_destroy": (_I, [_VP]),
    "cpvs_ctx_set_stream": (_I, [_VP, _VP]),
    "cpvs_ctx_get_stream": (_VP, [_VP]),
    "cpvs_ctx_reserve": (_I, [_VP, _U64]),
    "cpvs_ctx_trim": (_I, [_VP]),
    "cpvs_ctx_synchronize": (_I, [_VP]),
    "cpvs_ctx_launch_count": (_U64, [_VP]),
    "cpvs_ctx_set_prediction": (_I, [_VP, _I, _U32]),
    "cpvs_ctx_get_stats": (_I, [_VP, _VP]),
    "cpvs_last_error": (ctypes.c_char_p, []),
    "cpvs_version": (ctypes.c_char_p, []),
    "cpvs_minmax_build": (_I, [_VP, _VP, _I, _I, _PP]),
    "cpvs_minmax_build_tiled": (_I, [_VP, _VP, _I, _I, _U32, _PP]),
    "cpvs_minmax_destroy": (_I, [_VP]),
    "cpvs_minmax_num_levels": (_I, [_VP]),
    "cpvs_minmax_size": (_I, [_VP]),
    "cpvs_minmax_level": (_I, [_VP, _I, _VP]),
    "cpvs_minmax_level_device": (_VP, [_VP, _I]),
    "cpvs_minmax_childmask": (_I, [_VP, _U32, _U32, _U32, _U32, _U32, ctypes.POINTER(_U32)]),
    "cpvs_minmax_timing": (_I, [_VP, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
    "cpvs_shadow_create": (_I, [_VP, _VP, _U32, _U32, _I, _PP]),
    "cpvs_shadow_create_async": (_I, [_VP, _VP, _U32, _U32, _I, _PP]),
    "cpvs_shadow_wait": (_I, [_VP]),
    "cpvs_shadow_create_from_depth": (_I, [_VP, _VP, _I, _I, _U32, _U32, _I, _PP]),
    "cpvs_shadow_destroy": (_I, [_VP]),
    "cpvs_shadow_info_get": (_I, [_VP, ctypes.POINTER(ShadowInfo)]),
    "cpvs_shadow_copy_dag": (_I, [_VP, _VP]),
    "cpvs_shadow_dag_device": (_VP, [_VP]),
    "cpvs_shadow_lookup_ndc": (_I, [_VP, _VP, _I64, _I, _I, _VP]),
    "cpvs_container_create": (_I, [_VP, _U32, _PP]),
    "cpvs_container_destroy": (_I, [_VP]),
    "cpvs_container_set": (_I, [_VP, _VP, _U32, _U32, _U32]),
    "cpvs_container_set_dag": (_I, [_VP, _VP, _U64, _I, _U32, _I, _U32, _U32, _U32]),
    "cpvs_container_finalize": (_I, [_VP]),
    "cpvs_container_info": (_I, [_VP, ctypes.POINTER(_U64), ctypes.POINTER(_U32), ctypes.POINTER(_U32), ctypes.POINTER(_U32)]),
    "cpvs_container_copy": (_I, [_VP, _VP, _VP]),
    "cpvs_container_lookup_ndc": (_I, [_VP, _VP, _I64, _I, _VP]),
    "cpvs_container_evaluate": (_I, [_VP, _VP, _U32, _U32, _I, _VP, _VP]),
    "cpvs_container_set_filter_size": (_I, [_VP, _U32]),
    "cpvs_container_evaluate_surface": (_I, [_VP, _U64, _U64, _U32, _U32, _VP]),
    "cpvs_container_save": (_I, [_VP, ctypes.c_char_p]),
    "cpvs_container_load": (_I, [_VP, ctypes.c_char_p, _PP]),
    "cpvs_depth_generate": (_I, [_VP, _I, _I, _I, _I, _I, _VP]),
    "cpvs_container_assemble": (_I, [_VP, _U32, _U32, _I, _VP, _PP]),
    "cpvs_grid_worker_create": (_I, [_VP, _VP, _PP]),
    "cpvs_grid_worker_destroy": (_I, [_VP]),
    "cpvs_grid_worker_estimate": (_I, [_VP, _VP, _I, _VP]),
    "cpvs_grid_worker_release": (_I, [_VP, _VP, _I]),
    "cpvs_grid_worker_build": (_I, [_VP, _VP, _I]),
    "cpvs_grid_worker_build_from": (_I, [_VP, _VP, _VP]),
    "cpvs_grid_worker_num_cells": (_I, [_VP]),
    "cpvs_grid_worker_cells": (_I, [_VP, _VP, _I]),
    "cpvs_grid_worker_device_ms": (ctypes.c_float, [_VP]),
    "cpvs_grid_worker_depth_ms": (ctypes.c_float, [_VP]),
    "cpvs_grid_worker_export": (_I, [_VP, _VP, _VP, _I]),
    "cpvs_grid_worker_copy_cells": (_I, [_VP, _VP, _U64, _VP, _I]),
    "cpvs_ipc_open": (_I, [_VP, _I, _PP]),
    "cpvs_ipc_close": (_I, [_I, _VP]),
    "cpvs_grid_assign": (_I, [_VP, _I, _I, _VP, _VP]),
    "cpvs_grid_build": (_I, [_VP, _I, _VP, _I, _PP]),
    "cpvs_grid_destroy": (_I, [_VP]),
    "cpvs_grid_stats_get": (_I, [_VP, _VP]),
    "cpvs_grid_container": (_VP, [_VP, _I]),
    "cpvs_grid_lookup_ndc": (_I, [_VP, _VP, _I64, _VP]),
}

_lib = None


def load_library():
    """Loads libcpvs_b200.so (built by ``python -m cpvs_b200.build``); raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CpvsError(ECUDA, "%s not built -- run `python -m cpvs_b200.build`; there is no CPU fallback" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _check(rc):
    if rc != OK:
        raise CpvsError(rc, load_library().cpvs_last_error().decode())


def _as_ptr(array_or_ptr):
    """(pointer, mem kind) of a numpy array (host) or a torch CUDA tensor / raw int (device)."""
    if isinstance(array_or_ptr, np.ndarray):
        return array_or_ptr.ctypes.data, MEM_HOST
    if isinstance(array_or_ptr, int):
        return array_or_ptr, MEM_DEVICE
    if hasattr(array_or_ptr, "data_ptr"):  # torch tensor
        return array_or_ptr.data_ptr(), MEM_DEVICE if array_or_ptr.is_cuda else MEM_HOST
    raise TypeError("expected numpy array, torch tensor or device pointer")


class Context:
    """One per GPU: a stream plus the stream-ordered scratch pool."""

    def __init__(self, device=0, stream=None):
        self._lib = load_library()
        h = ctypes.c_void_p()
        _check(self._lib.cpvs_ctx_create(device, ctypes.byref(h)))
        self.handle = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def set_stream(self, stream):
        """``stream``: raw cudaStream_t as int (e.g. ``torch.cuda.current_stream().cuda_stream``) or None."""
        _check(self._lib.cpvs_ctx_set_stream(self.handle, ctypes.c_void_p(stream or 0)))

    def reserve(self, nbytes):
        """Pre-grows the stream-ordered pool the finished DAGs are allocated from (see cpvs_ctx_reserve)."""
        _check(self._lib.cpvs_ctx_reserve(self.handle, int(nbytes)))

    def trim(self):
        """Releases the memory the context keeps for recycling (see cpvs_ctx_trim)."""
        _check(self._lib.cpvs_ctx_trim(self.handle))

    def synchronize(self):
        _check(self._lib.cpvs_ctx_synchronize(self.handle))

    @property
    def launch_count(self):
        return int(self._lib.cpvs_ctx_launch_count(self.handle))

    def set_prediction(self, enabled=True, headroom_shift=3):
        """Sizing of builds from the previous build of the same shape (see cpvs_ctx_set_prediction)."""
        _check(self._lib.cpvs_ctx_set_prediction(self.handle, int(bool(enabled)), int(headroom_shift)))

    def stats(self):
        st = CtxStats()
        _check(self._lib.cpvs_ctx_get_stats(self.handle, ctypes.byref(st)))
        return {name: int(getattr(st, name)) for name, _ in CtxStats._fields_}

    def close(self):
        if self.handle:
            self._lib.cpvs_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


SCENES = {"plane": 0, "city": 2, "terrain_dev": 3}  # CPVS_SCENE_*: the scenes with a device generator


def generate_depth(kind, n, out, tile=(0, 0), tiles_per_side=1, ctx=None):
    """Writes window ``tile`` of the synthetic scene ``kind`` into ``out`` (torch CUDA float32 [n, n] or a raw
    device pointer) on the context's stream: the device-resident depth source of SURVEY.md 8f.3, standing in
    for the reference's render + read-back (``src/ShadowMap.cpp:23-30``). Same bytes as ``cpvs_b200.synth``."""
    ctx = ctx or default_context()
    ptr, mem = _as_ptr(out)
    if mem != MEM_DEVICE:
        raise TypeError("generate_depth writes device memory")
    _check(ctx._lib.cpvs_depth_generate(ctx.handle, SCENES[kind], n, tile[0], tile[1], tiles_per_side, ctypes.c_void_p(ptr)))
    return out


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class MinMaxHierarchy:
    """Reference ``MinMaxHierarchy`` (src/MinMaxHierarchy.h:23-72), built on the GPU."""

    def __init__(self, depth, ctx=None, n=None, zTileNum=1):
        """``zTileNum``: the slicing the hierarchy's builds will use (a speed hint, see cpvs_minmax_build_tiled)."""
        self.ctx = ctx or default_context()
        self._lib = self.ctx._lib
        if isinstance(depth, np.ndarray):
            depth = np.ascontiguousarray(depth, dtype=np.float32)
            if depth.ndim != 2 or depth.shape[0] != depth.shape[1]:
                raise CpvsError(EINVAL, "depth map must be square")  # assert of src/MinMaxHierarchy.cpp:12
            n = depth.shape[0]
        elif n is None:
            n = int(depth.shape[0])
        self._keepalive = depth
        ptr, mem = _as_ptr(depth)
        h = ctypes.c_void_p()
        _check(self._lib.cpvs_minmax_build_tiled(self.ctx.handle, ctypes.c_void_p(ptr), n, mem, zTileNum, ctypes.byref(h)))
        self.handle = h
        self.n = n

    def getNumLevels(self):
        return int(self._lib.cpvs_minmax_num_levels(self.handle))

    def getLevel(self, level):
        side = self.n >> level
        out = np.empty((side, side) if level == 0 else (side, side, 2), np.float32)
        _check(self._lib.cpvs_minmax_level(self.handle, level, out.ctypes.data))
        return out

    def createChildmask(self, level, x, y, z, zTileNum=1):
        """``cs::createChildmask(minMax, level, ivec3(x, y, z))`` (reference src/CompressedShadowUtil.cpp:20-54)."""
        out = _U32()
        _check(self._lib.cpvs_minmax_childmask(self.handle, level, x, y, z, zTileNum, ctypes.byref(out)))
        return out.value

    def timing(self):
        """(total ms, fused base kernel ms) of the build, from CUDA events."""
        total, base = ctypes.c_float(), ctypes.c_float()
        _check(self._lib.cpvs_minmax_timing(self.handle, ctypes.byref(total), ctypes.byref(base)))
        return total.value, base.value

    def getMin(self, level, x, y):
        lvl = self.getLevel(level)
        return float(lvl[y, x] if level == 0 else lvl[y, x, 0])

    def getMax(self, level, x, y):
        lvl = self.getLevel(level)
        return float(lvl[y, x] if level == 0 else lvl[y, x, 1])

    def close(self):
        if getattr(self, "handle", None):
            self._lib.cpvs_minmax_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CompressedShadow:
    """Reference ``CompressedShadow`` (src/CompressedShadow.h:20-126)."""

    SHADOW, VISIBLE, PARTIAL = SHADOW, VISIBLE, PARTIAL

    def __init__(self, ctx, handle, wait=True, keep=None):
        self.ctx = ctx
        self._lib = ctx._lib
        self.handle = handle
        self._info = None
        self._keep = keep  # the hierarchy of a build in flight
        if wait:
            self.wait()

    @classmethod
    def create(cls, minmax, zTileIndex=0, zTileNum=1, leafmasks=True, ctx=None, wait=True):
        """``CompressedShadow::create(minMax, zTileIndex, zTileNum)`` (src/CompressedShadow.cpp:49-59).
        ``wait=False``: return while the build is in flight (``cpvs_shadow_create_async``); ``wait()`` or any accessor finishes it."""
        ctx = ctx or minmax.ctx
        h = ctypes.c_void_p()
        fn = ctx._lib.cpvs_shadow_create if wait else ctx._lib.cpvs_shadow_create_async
        _check(fn(ctx.handle, minmax.handle, zTileIndex, zTileNum, int(leafmasks), ctypes.byref(h)))
        return cls(ctx, h, wait, None if wait else minmax)

    def wait(self):
        if self._info is None:
            _check(self._lib.cpvs_shadow_wait(self.handle))
            info = ShadowInfo()
            _check(self._lib.cpvs_shadow_info_get(self.handle, ctypes.byref(info)))
            self._info = info
            self._keep = None
        return self

    @property
    def info(self):
        return self.wait()._info

    def getNumLevels(self):
        return int(self.info.num_levels)

    def getTotalVisibility(self):
        return int(self.info.total_visibility)

    def getDAG(self):
        out = np.empty(int(self.info.words), np.uint32)
        _check(self._lib.cpvs_shadow_copy_dag(self.handle, out.ctypes.data))
        return out

    @property
    def dag_device_ptr(self):
        return int(self._lib.cpvs_shadow_dag_device(self.handle) or 0)

    def phase_ms(self):
        return {name: float(self.info.phase_ms[i]) for i, name in enumerate(PHASE_NAMES)}

    def level_counts(self):
        """(SVO nodes, DAG nodes, DAG words) per level, index = level."""
        nl = self.getNumLevels()
        return (np.array(self.info.svo_nodes[:nl - 1], np.uint64), np.array(self.info.dag_nodes[:nl - 1], np.uint64),
                np.array(self.info.dag_words[:nl - 1], np.uint64))

    def traverse(self, ndc, tryLeafmasks=True, out=None):
        """``traverse(vec3 ndc, bool tryLeafmasks)`` (src/CompressedShadow.cpp:404-463), batched."""
        if isinstance(ndc, np.ndarray):
            ndc = np.ascontiguousarray(ndc, dtype=np.float32).reshape(-1, 3)
            count = len(ndc)
            if out is None:
                out = np.empty(count, np.uint8)
        else:
            count = int(ndc.shape[0])
        ptr, mem = _as_ptr(ndc)
        optr, omem = _as_ptr(out)
        assert mem == omem
        _check(self._lib.cpvs_shadow_lookup_ndc(self.handle, ctypes.c_void_p(ptr), count, mem, int(tryLeafmasks), ctypes.c_void_p(optr)))
        return out

    def close(self):
        if getattr(self, "handle", None):
            self._lib.cpvs_shadow_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CompressedShadowContainer:
    """Reference ``CompressedShadowContainer`` (src/CompressedShadowContainer.h:18-92) + ``shader/traverse.cs``."""

    def __init__(self, length_or_shadow, ctx=None):
        if isinstance(length_or_shadow, CompressedShadow):  # ctor of src/CompressedShadowContainer.h:28-32
            shadow, length = length_or_shadow, 1
            ctx = ctx or shadow.ctx
        else:
            shadow, length = None, int(length_or_shadow)
        self.ctx = ctx or default_context()
        self._lib = self.ctx._lib
        h = ctypes.c_void_p()
        _check(self._lib.cpvs_container_create(self.ctx.handle, length, ctypes.byref(h)))
        self.handle = h
        self.length = length
        if shadow is not None:
            self.set(shadow, 0, 0, 0)

    def set(self, shadow, x, y, z):
        _check(self._lib.cpvs_container_set(self.handle, shadow.handle, x, y, z))

    def save(self, path):
        """Writes the finalized container (grid + combined DAG) to ``path``."""
        _check(self._lib.cpvs_container_save(self.handle, os.fsencode(path)))

    @classmethod
    def load(cls, path, ctx=None):
        """A finalized container from a file written by :meth:`save`; ready for lookups."""
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        self._lib = self.ctx._lib
        h = ctypes.c_void_p()
        _check(self._lib.cpvs_container_load(self.ctx.handle, os.fsencode(path), ctypes.byref(h)))
        self.handle = h
        self.length = round(self.info()["grid_cells"] ** (1.0 / 3.0))
        return self

    def set_dag(self, words, num_levels, leafmasks, x, y, z):
        """A cell built on another GPU / rank: hand over its finished DAG words."""
        if isinstance(words, np.ndarray):
            words = np.ascontiguousarray(words, dtype=np.uint32)
            count = words.size
        else:
            count = int(words.numel())
        self._keep = words
        ptr, mem = _as_ptr(words)
        _check(self._lib.cpvs_container_set_dag(self.handle, ctypes.c_void_p(ptr), count, mem, num_levels, int(leafmasks), x, y, z))

    def copyToGPU(self):
        """``copyToGPU`` (src/CompressedShadowContainer.cpp:31-46): combined DAG + top-level grid."""
        _check(self._lib.cpvs_container_finalize(self.handle))

    moveToGPU = copyToGPU

    def setFilterSize(self, size):
        _check(self._lib.cpvs_container_set_filter_size(self.handle, size))

    def info(self):
        words, cells, dl, gl = _U64(), _U32(), _U32(), _U32()
        _check(self._lib.cpvs_container_info(self.handle, ctypes.byref(words), ctypes.byref(cells), ctypes.byref(dl), ctypes.byref(gl)))
        return {"dag_words": words.value, "grid_cells": cells.value, "dag_levels": dl.value, "grid_levels": gl.value}

    def dag_and_grid(self):
        info = self.info()
        dag = np.empty(info["dag_words"], np.uint32)
        grid = np.empty(info["grid_cells"], np.uint32)
        _check(self._lib.cpvs_container_copy(self.handle, dag.ctypes.data, grid.ctypes.data))
        return dag, grid

    def lookup_ndc(self, ndc, out=None):
        """``traverse()`` of shader/traverse.cs:75-133 on NDC points over the whole virtual volume."""
        if isinstance(ndc, np.ndarray):
            ndc = np.ascontiguousarray(ndc, dtype=np.float32).reshape(-1, 3)
            count = len(ndc)
            if out is None:
                out = np.empty(count, np.uint8)
        else:
            count = int(ndc.shape[0])
        ptr, mem = _as_ptr(ndc)
        optr, _ = _as_ptr(out)
        _check(self._lib.cpvs_container_lookup_ndc(self.handle, ctypes.c_void_p(ptr), count, mem, ctypes.c_void_p(optr)))
        return out

    def evaluate(self, positionsWS, lightViewProj, visibilities=None):
        """``evaluate(positionsWS, lightViewProj, visibilities)`` (src/CompressedShadowContainer.cpp:93-124)."""
        m = np.ascontiguousarray(lightViewProj, dtype=np.float32).reshape(16)
        if isinstance(positionsWS, np.ndarray):
            positionsWS = np.ascontiguousarray(positionsWS, dtype=np.float32)
            if visibilities is None:
                visibilities = np.empty(positionsWS.shape[:2], np.uint8)
        height, width = int(positionsWS.shape[0]), int(positionsWS.shape[1])
        ptr, mem = _as_ptr(positionsWS)
        optr, _ = _as_ptr(visibilities)
        _check(self._lib.cpvs_container_evaluate(self.handle, ctypes.c_void_p(ptr), width, height, mem, m.ctypes.data, ctypes.c_void_p(optr)))
        return visibilities

    def close(self):
        if getattr(self, "handle", None):
            self._lib.cpvs_container_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
