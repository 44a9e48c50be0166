"""Tile grids through the C++ driver (``cpvs_grid_*`` in include/cpvs_b200.h, csrc/grid.cu): the reference's
``DeferredRenderer::renderWithTiles`` / ``createShadowTiles`` / ``precomputeShadows`` (reference
``src/DeferredRenderer.cpp:150-235``) over the GPUs of one box -- one host thread per GPU inside the library, host-side gather
of sizes, peer copies for the replication, no NCCL. This module only binds it with ctypes.

``Grid.build`` runs the whole thing in this process; ``GridWorker`` is the per-GPU half for callers that run one process per
GPU (``cpvs_b200.gridbuild``)."""
import ctypes

import numpy as np

from . import CpvsError, EINVAL, SCENES, _check, load_library, CompressedShadowContainer, Context

MAX_DEVICES = 16
FETCH_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_float))
NEXT_TILE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32))


class GridDesc(ctypes.Structure):
    _fields_ = [("length", ctypes.c_uint32), ("tile", ctypes.c_int32), ("leafmasks", ctypes.c_int32), ("scene", ctypes.c_int32),
                ("fetch", FETCH_FN), ("user", ctypes.c_void_p)]


class GridCell(ctypes.Structure):
    _fields_ = [("index", ctypes.c_uint32), ("num_levels", ctypes.c_uint32), ("words", ctypes.c_uint64), ("root_mask", ctypes.c_uint32),
                ("device", ctypes.c_int32), ("words_device", ctypes.c_void_p), ("svo_nodes", ctypes.c_uint64), ("dag_nodes", ctypes.c_uint64)]


class CellPart(ctypes.Structure):
    _fields_ = [("words", ctypes.c_uint64), ("root_mask", ctypes.c_uint32), ("device", ctypes.c_int32), ("words_device", ctypes.c_void_p)]


class GridStats(ctypes.Structure):
    _fields_ = [("devices", ctypes.c_uint32), ("cells", ctypes.c_uint32), ("one_word_cells", ctypes.c_uint32), ("moved_tiles", ctypes.c_uint32),
                ("dag_words", ctypes.c_uint64), ("svo_nodes", ctypes.c_uint64), ("dag_nodes", ctypes.c_uint64), ("launches", ctypes.c_uint64),
                ("build_ms_max", ctypes.c_float), ("build_ms", ctypes.c_float * MAX_DEVICES), ("depth_ms", ctypes.c_float * MAX_DEVICES),
                ("tiles", ctypes.c_uint32 * MAX_DEVICES),
                ("build_wall_ms", ctypes.c_float), ("gather_ms", ctypes.c_float), ("replicate_ms", ctypes.c_float), ("wall_ms", ctypes.c_float)]


def make_desc(length, tile, kind=None, fetch=None, leafmasks=True):
    """``kind``: a scene with a device generator (``cpvs_b200.SCENES``); otherwise ``fetch(x, y, out)`` fills the float32
    [tile, tile] array ``out`` with depth tile (x, y). Returns (desc, keepalive)."""
    desc = GridDesc()
    desc.length, desc.tile, desc.leafmasks = length, tile, int(bool(leafmasks))
    keep = None
    if kind is not None:
        if kind not in SCENES:
            raise CpvsError(EINVAL, "scene %r has no device generator" % (kind,))
        desc.scene = SCENES[kind]
        desc.fetch = FETCH_FN()
    else:
        def trampoline(_user, x, y, out):
            try:
                fetch(int(x), int(y), np.ctypeslib.as_array(out, shape=(tile, tile)))
                return 0
            except Exception:  # noqa: BLE001 -- reported through the status code
                return 1
        keep = FETCH_FN(trampoline)
        desc.scene = -1
        desc.fetch = keep
    return desc, keep


class _BorrowedContainer(CompressedShadowContainer):
    """A container owned by a Grid."""

    def __init__(self, handle, length, device):
        self._lib = load_library()
        self.handle = ctypes.c_void_p(handle)
        self.length = length
        self.device = device
        self.ctx = None

    def close(self):
        self.handle = None


class Grid:
    def __init__(self, handle, length, devices, keep):
        self._lib = load_library()
        self.handle = handle
        self.length = length
        self.devices = list(devices)
        self._keep = keep

    @classmethod
    def build(cls, devices, length, tile, kind=None, fetch=None, leafmasks=True, replicate=True):
        lib = load_library()
        desc, keep = make_desc(length, tile, kind, fetch, leafmasks)
        devs = (ctypes.c_int * len(devices))(*devices)
        h = ctypes.c_void_p()
        _check(lib.cpvs_grid_build(devs, len(devices), ctypes.byref(desc), int(bool(replicate)), ctypes.byref(h)))
        return cls(h, length, devices, keep)

    def stats(self):
        st = GridStats()
        _check(self._lib.cpvs_grid_stats_get(self.handle, ctypes.byref(st)))
        out = {name: getattr(st, name) for name, _ in GridStats._fields_ if name not in ("build_ms", "tiles", "depth_ms")}
        out["build_ms"] = [float(v) for v in st.build_ms[:st.devices]]
        out["depth_ms"] = [float(v) for v in st.depth_ms[:st.devices]]
        out["tiles"] = [int(v) for v in st.tiles[:st.devices]]
        return out

    def container(self, index=0):
        h = self._lib.cpvs_grid_container(self.handle, index)
        if not h:
            raise CpvsError(EINVAL, "grid has no container %d" % index)
        return _BorrowedContainer(h, self.length, self.devices[index])

    def lookup_ndc(self, ndc):
        ndc = np.ascontiguousarray(ndc, dtype=np.float32).reshape(-1, 3)
        out = np.empty(len(ndc), np.uint8)
        _check(self._lib.cpvs_grid_lookup_ndc(self.handle, ndc.ctypes.data, len(ndc), out.ctypes.data))
        return out

    def close(self):
        if getattr(self, "handle", None):
            self._lib.cpvs_grid_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class GridWorker:
    """The per-GPU half of a grid build: estimates, builds and keeps the cells of the xy tiles it is given."""

    def __init__(self, ctx, length, tile, kind=None, fetch=None, leafmasks=True):
        self._lib = load_library()
        self.ctx = ctx
        self.length = length
        self.desc, self._keep = make_desc(length, tile, kind, fetch, leafmasks)
        h = ctypes.c_void_p()
        _check(self._lib.cpvs_grid_worker_create(ctx.handle, ctypes.byref(self.desc), ctypes.byref(h)))
        self.handle = h

    @staticmethod
    def _pairs(tiles):
        flat = [int(v) for xy in tiles for v in xy]
        return (ctypes.c_uint32 * max(1, len(flat)))(*flat), len(flat) // 2

    def estimate(self, tiles):
        arr, n = self._pairs(tiles)
        cost = (ctypes.c_uint64 * max(1, n))()
        _check(self._lib.cpvs_grid_worker_estimate(self.handle, arr, n, cost))
        return [int(c) for c in cost[:n]]

    def release(self, tiles):
        arr, n = self._pairs(tiles)
        _check(self._lib.cpvs_grid_worker_release(self.handle, arr, n))

    def build(self, tiles):
        arr, n = self._pairs(tiles)
        _check(self._lib.cpvs_grid_worker_build(self.handle, arr, n))

    def build_from(self, next_tile):
        """Builds the tiles ``next_tile()`` hands out -- (x, y), or None when there are no more (a shared queue). It is called one
        tile ahead of the build (``cpvs_grid_worker_build_from``)."""
        error = []

        def pull(_user, x, y):
            try:
                t = next_tile()
            except Exception as e:  # noqa: BLE001 -- must not unwind through the C frames
                error.append(e)
                return 0
            if t is None:
                return 0
            x[0], y[0] = int(t[0]), int(t[1])
            return 1

        cb = NEXT_TILE_FN(pull)
        _check(self._lib.cpvs_grid_worker_build_from(self.handle, cb, None))
        if error:
            raise error[0]

    def cells(self):
        n = self._lib.cpvs_grid_worker_num_cells(self.handle)
        arr = (GridCell * max(1, n))()
        got = self._lib.cpvs_grid_worker_cells(self.handle, arr, max(1, n))
        if got < 0:
            _check(EINVAL)
        return list(arr[:got])

    def device_ms(self):
        return float(self._lib.cpvs_grid_worker_device_ms(self.handle))

    def depth_ms(self):
        return float(self._lib.cpvs_grid_worker_depth_ms(self.handle))

    def export(self):
        """(64-byte CUDA IPC handle of a block holding all finished cells, first word of every cell in ``cells()`` order)."""
        n = self._lib.cpvs_grid_worker_num_cells(self.handle)
        handle = (ctypes.c_ubyte * 64)()
        offsets = (ctypes.c_uint64 * max(1, n))()
        got = self._lib.cpvs_grid_worker_export(self.handle, handle, offsets, max(1, n))
        if got < 0:
            _check(EINVAL)
        return bytes(handle), [int(o) for o in offsets[:got]]

    def copy_cells(self, out=None):
        """All finished cells copied into the uint32 numpy array ``out`` (host; None: sizes only).
        Returns (first word of every cell in ``cells()`` order, total words)."""
        n = self._lib.cpvs_grid_worker_num_cells(self.handle)
        offsets = (ctypes.c_uint64 * max(1, n))()
        ptr, cap = (out.ctypes.data, out.size) if out is not None else (None, 0)
        got = self._lib.cpvs_grid_worker_copy_cells(self.handle, ptr, cap, offsets, max(1, n))
        if got < 0:
            _check(EINVAL)
        offs = [int(o) for o in offsets[:got]]
        return offs

    def close(self):
        if getattr(self, "handle", None):
            self._lib.cpvs_grid_worker_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def assign(costs, workers, owner_in=None):
    """``cpvs_grid_assign``: longest tile first to the least loaded worker."""
    lib = load_library()
    n = len(costs)
    c = (ctypes.c_uint64 * max(1, n))(*costs)
    out = (ctypes.c_int * max(1, n))()
    oin = (ctypes.c_int * max(1, n))(*owner_in) if owner_in is not None else None
    _check(lib.cpvs_grid_assign(c, n, workers, oin, out))
    return [int(v) for v in out[:n]]


def assemble(ctx, length, num_levels, leafmasks, parts):
    """``cpvs_container_assemble`` from a list of (words, root_mask, device, device pointer) in container order."""
    lib = load_library()
    arr = (CellPart * len(parts))()
    for i, (words, mask, device, ptr) in enumerate(parts):
        arr[i].words, arr[i].root_mask, arr[i].device, arr[i].words_device = int(words), int(mask), int(device), ctypes.c_void_p(int(ptr) if ptr else None)
    h = ctypes.c_void_p()
    _check(lib.cpvs_container_assemble(ctx.handle, length, num_levels, int(bool(leafmasks)), arr, ctypes.byref(h)))
    cont = CompressedShadowContainer.__new__(CompressedShadowContainer)
    cont.ctx, cont._lib, cont.handle, cont.length = ctx, lib, h, length
    return cont


def ipc_open(handle, device):
    """Maps another process's exported block (``GridWorker.export``) on ``device``; returns the device pointer."""
    lib = load_library()
    buf = (ctypes.c_ubyte * 64)(*handle)
    ptr = ctypes.c_void_p()
    _check(lib.cpvs_ipc_open(buf, device, ctypes.byref(ptr)))
    return int(ptr.value)


def ipc_close(device, ptr):
    _check(load_library().cpvs_ipc_close(device, ctypes.c_void_p(ptr)))
