"""TEST-ONLY: builds tests/hostsim/hostsim.cpp (a g++ build of the device rule code) and loads it with
ctypes.  Used by the CPU-side tests to check chess_core.cuh against the oracle without a GPU."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_hostsim.so")
_SRC = os.path.join(_HERE, "hostsim.cpp")
_CSRC = os.path.join(_HERE, "..", "..", "chessrl_b200", "csrc")
_DEPS = [os.path.join(_CSRC, f) for f in ("chess_core.cuh", "tree_core.cuh", "hash_eval.cuh", "warp_gen.cuh")]


def load():
    newest = max([os.path.getmtime(_SRC)] + [os.path.getmtime(d) for d in _DEPS])
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", _SRC, "-o", _SO])
    lib = ctypes.CDLL(_SO)
    u64p = ctypes.POINTER(ctypes.c_uint64)
    lib.hs_movegen.argtypes = [u64p, ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_int)]
    lib.hs_movegen.restype = ctypes.c_int
    lib.hs_movegen_warp.argtypes = [u64p, ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_int)]
    lib.hs_movegen_warp.restype = ctypes.c_int
    lib.hs_make.argtypes = [u64p, ctypes.c_uint16]
    lib.hs_make.restype = None
    lib.hs_key.argtypes = [u64p, ctypes.c_int]
    lib.hs_key.restype = ctypes.c_uint64
    lib.hs_eval_hash.argtypes = [u64p, ctypes.c_uint64]
    lib.hs_eval_hash.restype = ctypes.c_uint64
    lib.hs_result.argtypes = [u64p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.hs_result.restype = ctypes.c_int
    lib.hs_perft.argtypes = [u64p, ctypes.c_int, ctypes.c_int]
    lib.hs_perft.restype = ctypes.c_uint64
    return lib
