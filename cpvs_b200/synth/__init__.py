"""Synthetic depth maps and lookup points -- host-side workload generators.

Thin ctypes wrapper over ``synth.cpp`` (built in-tree by ``__graft_entry__.build()``). The maps
replace what the reference renders and reads back in ``ShadowMap::createImageF`` (reference
``src/ShadowMap.cpp:23-30``). No CUDA and no oracle code in here.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcpvs_synth.so")
_lib = None

KINDS = {"plane": 0, "terrain": 1, "city": 2, "terrain_dev": 3}


def build(force=False):
    src = os.path.join(_HERE, "synth.cpp")
    newest = max(os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "scene.h")))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < newest:
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", _LIB_PATH, "-lpthread"])


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        lib = ctypes.CDLL(_LIB_PATH)
        lib.cpvs_synth_depth.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p]
        lib.cpvs_synth_depth.restype = ctypes.c_int
        lib.cpvs_synth_lookups.argtypes = [ctypes.c_uint32, ctypes.c_long, ctypes.c_void_p]
        lib.cpvs_synth_lookups.restype = None
        lib.cpvs_synth_fnv64.argtypes = [ctypes.c_void_p, ctypes.c_long]
        lib.cpvs_synth_fnv64.restype = ctypes.c_uint64
        _lib = lib
    return _lib


def depth_map(kind, n, tile=(0, 0), tiles_per_side=1, threads=None, out=None):
    """float32 [n, n] depth map; ``tile``/``tiles_per_side`` pick a window of a virtual map."""
    lib = _load()
    if out is None:
        out = np.empty((n, n), np.float32)
    assert out.dtype == np.float32 and out.size == n * n and out.flags.c_contiguous
    threads = threads or min(64, os.cpu_count() or 1)
    rc = lib.cpvs_synth_depth(KINDS[kind], n, tile[0], tile[1], tiles_per_side, threads, out.ctypes.data)
    if rc != 0:
        raise ValueError("bad synth arguments")
    return out


def lookups(count, seed=777):
    """float32 [count, 3] points in [-1, 1]^3."""
    out = np.empty((count, 3), np.float32)
    _load().cpvs_synth_lookups(seed, count, out.ctypes.data)
    return out


def fnv64(words):
    words = np.ascontiguousarray(words, dtype=np.uint32)
    return int(_load().cpvs_synth_fnv64(words.ctypes.data, words.size))
