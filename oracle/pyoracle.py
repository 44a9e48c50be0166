"""TEST INFRASTRUCTURE ONLY: ctypes access to the CPU oracle.

Two back ends with one interface:
  * ``port``            oracle/_build/libcpvs_oracle.so  -- this repo's restatement (oracle_port.cpp)
  * ``ref``/``ref_noleaf`` oracle/_ref/libcpvs_ref*.so   -- the unmodified reference, compiled from
                        /root/reference by oracle/Makefile (only where that directory exists; the
                        prebuilt libraries travel to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_DIR = "/root/reference"
_PORT = os.path.join(HERE, "_build", "libcpvs_oracle.so")
_REF = os.path.join(HERE, "_ref", "libcpvs_ref.so")
_REF_NOLEAF = os.path.join(HERE, "_ref", "libcpvs_ref_noleaf.so")


def build(ref=True):
    """Compile the port (always) and, where /root/reference exists, the reference itself."""
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if ref and os.path.isdir(REFERENCE_DIR):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", "REF=" + REFERENCE_DIR])
        if os.path.exists(os.path.join(HERE, "..", "cpvs_b200", "libcpvs_b200.so")):
            # the reference's gtests against this repo's C++ facade + CUDA library (run on the GPU box)
            subprocess.check_call(["make", "-s", "-C", HERE, "facade-tests", "REF=" + REFERENCE_DIR])


def have_ref():
    return os.path.exists(_REF) and os.path.exists(_REF_NOLEAF)


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data)


class _Backend:
    def __init__(self, path, prefix):
        self.lib = ctypes.CDLL(path)
        self.prefix = prefix
        self.is_port = prefix == "orc_"
        for name, res in [("minmax_create", ctypes.c_void_p), ("shadow_create", ctypes.c_void_p),
                          ("shadow_words", ctypes.c_long), ("minmax_level", ctypes.c_long),
                          ("svo", ctypes.c_long), ("container_create", ctypes.c_void_p),
                          ("container_words", ctypes.c_long)]:
            if hasattr(self.lib, prefix + name):
                getattr(self.lib, prefix + name).restype = res

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)


_backends = {}


def backend(kind="port"):
    if kind not in _backends:
        if kind == "port":
            if not os.path.exists(_PORT):
                build(ref=False)
            _backends[kind] = _Backend(_PORT, "orc_")
        elif kind == "ref":
            _backends[kind] = _Backend(_REF, "ref_")
        elif kind == "ref_noleaf":
            _backends[kind] = _Backend(_REF_NOLEAF, "ref_")
        else:
            raise ValueError(kind)
    return _backends[kind]


class MinMax:
    """MinMaxHierarchy (reference src/MinMaxHierarchy.h:23-72)."""

    def __init__(self, depth, kind="port"):
        self.b = backend(kind)
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        self.n = depth.shape[0]
        assert depth.shape == (self.n, self.n)
        self.h = ctypes.c_void_p(self.b.fn("minmax_create")(_vp(depth), self.n))

    def num_levels(self):
        return int(self.b.fn("minmax_num_levels")(self.h))

    def level(self, level):
        side = self.n >> level
        out = np.empty((side, side) if level == 0 else (side, side, 2), np.float32)
        self.b.fn("minmax_level")(self.h, level, _vp(out))
        return out

    def childmask(self, level, x, y, z, z_tile_num=1):
        return int(self.b.fn("create_childmask")(self.h, level, x, y, z, z_tile_num)) & 0xFFFFFFFF

    def svo(self, z_tile=0, z_num=1, leafmasks=True):
        """Uncompressed SVO words and the reference's levelOffsets vector."""
        nl = self.num_levels()
        offs = np.zeros(nl - 1, np.uint32)
        args = (self.h, z_tile, z_num) + ((int(leafmasks),) if self.b.is_port else ())
        words = self.b.fn("svo")(*args, None, _vp(offs))
        out = np.empty(words, np.uint32)
        self.b.fn("svo")(*args, _vp(out), _vp(offs))
        return out, offs

    def __del__(self):
        if getattr(self, "h", None):
            self.b.fn("minmax_destroy")(self.h)
            self.h = None


class Shadow:
    """CompressedShadow (reference src/CompressedShadow.h:20-126)."""

    def __init__(self, mm, z_tile=0, z_num=1, leafmasks=True):
        self.b = mm.b
        self.mm = mm
        if self.b.is_port:
            self.h = ctypes.c_void_p(self.b.fn("shadow_create")(mm.h, z_tile, z_num, int(leafmasks)))
        else:
            assert bool(self.b.lib.ref_leafmasks_compiled()) == bool(leafmasks), "use kind='ref_noleaf'"
            self.h = ctypes.c_void_p(self.b.fn("shadow_create")(mm.h, z_tile, z_num))

    def num_levels(self):
        return int(self.b.fn("shadow_num_levels")(self.h))

    def total_visibility(self):
        return int(self.b.fn("shadow_total_visibility")(self.h))

    def dag(self):
        out = np.empty(self.b.fn("shadow_words")(self.h), np.uint32)
        self.b.fn("shadow_copy_dag")(self.h, _vp(out))
        return out

    def level_counts(self):
        assert self.b.is_port
        nl = self.num_levels()
        svo = np.zeros(nl - 1, np.uint64)
        uniq = np.zeros(nl - 1, np.uint64)
        self.b.fn("shadow_level_counts")(self.h, _vp(svo), _vp(uniq))
        return svo, uniq

    def traverse(self, ndc, try_leafmasks=True, threads=1):
        ndc = np.ascontiguousarray(ndc, dtype=np.float32).reshape(-1, 3)
        out = np.empty(len(ndc), np.uint8)
        if threads > 1 and not self.b.is_port:
            self.b.fn("shadow_traverse_mt")(self.h, _vp(ndc), len(ndc), int(try_leafmasks), _vp(out), threads)
        else:
            self.b.fn("shadow_traverse")(self.h, _vp(ndc), len(ndc), int(try_leafmasks), _vp(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.b.fn("shadow_destroy")(self.h)
            self.h = None


class Container:
    """CompressedShadowContainer + shader/traverse.cs (port only: the reference's needs GL)."""

    def __init__(self, length):
        self.b = backend("port")
        self.length = length
        self.h = ctypes.c_void_p(self.b.fn("container_create")(length))

    def set(self, shadow, x, y, z):
        self.b.fn("container_set")(self.h, shadow.h, x, y, z)

    def finalize(self):
        self.b.fn("container_finalize")(self.h)

    def dag_and_grid(self):
        dag = np.empty(self.b.fn("container_words")(self.h), np.uint32)
        grid = np.empty(self.length ** 3, np.uint32)
        self.b.fn("container_copy")(self.h, _vp(dag), _vp(grid))
        return dag, grid

    def lookup_ndc(self, ndc):
        ndc = np.ascontiguousarray(ndc, dtype=np.float32).reshape(-1, 3)
        out = np.empty(len(ndc), np.uint8)
        self.b.fn("container_lookup_ndc")(self.h, _vp(ndc), len(ndc), _vp(out))
        return out

    def evaluate(self, positions, matrix):
        positions = np.ascontiguousarray(positions, dtype=np.float32)
        h, w = positions.shape[:2]
        m = np.ascontiguousarray(matrix, dtype=np.float32).reshape(16)
        out = np.empty((h, w), np.uint8)
        self.b.fn("container_evaluate")(self.h, _vp(positions), w, h, _vp(m), _vp(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.b.fn("container_destroy")(self.h)
            self.h = None


def merge_level(level_words, node_size, kind="port"):
    """cs::mergeLevel (reference src/CompressedShadowUtil.h:154-182): (kept, merged words, mapping)."""
    b = backend(kind)
    level_words = np.ascontiguousarray(level_words, dtype=np.uint32)
    merged = np.zeros_like(level_words)
    mapping = np.zeros(level_words.size // node_size, np.uint32)
    kept = b.fn("merge_level")(_vp(level_words), level_words.size, node_size, _vp(merged), _vp(mapping))
    return int(kept), merged, mapping


def ref_time_build(depth, z_tile=0, z_num=1):
    """(ms pyramid, ms create, words) of the unmodified reference on this host."""
    b = backend("ref")
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    a, c, w = ctypes.c_double(), ctypes.c_double(), ctypes.c_long()
    b.lib.ref_time_build(_vp(depth), depth.shape[0], z_tile, z_num, ctypes.byref(a), ctypes.byref(c), ctypes.byref(w))
    return a.value, c.value, w.value
