"""TEST INFRASTRUCTURE ONLY (oracle) -- ctypes driver for oracle/_ref/libctcdec_ref.so: the reference's own
externals/ctc_decoders C++ (ctc_beam_search_decoder.cpp:18-187, ctc_greedy_decoder.cpp:4-45) built by
oracle/build_ref.py."""
from __future__ import annotations

import ctypes
import os
from typing import List, Tuple

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libctcdec_ref.so")
_lib = None


def available() -> bool:
    return os.path.isfile(_PATH)


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_PATH)
        lib.ctcref_beam.restype = ctypes.c_int
        lib.ctcref_beam.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.ctcref_greedy.restype = ctypes.c_int
        lib.ctcref_greedy.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib = lib
    return _lib


def beam_search(probs: np.ndarray, beam: int, cutoff_prob: float = 1.0, cutoff_top_n: int = 40) -> List[Tuple[float, List[int]]]:
    p = np.ascontiguousarray(probs, dtype=np.float64)
    T, V = p.shape
    ids = np.zeros((beam, max(T, 1)), dtype=np.int32)
    lens = np.zeros(beam, dtype=np.int32)
    scores = np.zeros(beam, dtype=np.float64)
    n = _load().ctcref_beam(p.ctypes.data, T, V, beam, cutoff_prob, cutoff_top_n, ids.ctypes.data, lens.ctypes.data,
                            scores.ctypes.data)
    return [(float(scores[i]), ids[i, :lens[i]].tolist()) for i in range(n)]


def greedy(probs: np.ndarray) -> List[int]:
    p = np.ascontiguousarray(probs, dtype=np.float64)
    T, V = p.shape
    ids = np.zeros(max(T, 1), dtype=np.int32)
    n = _load().ctcref_greedy(p.ctypes.data, T, V, ids.ctypes.data)
    return ids[:n].tolist()
