"""ctypes loader for the CPU oracle (oracle/pbx_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (pixelbox_b200) must never import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpbx_oracle.so")

_u8p = ctypes.POINTER(ctypes.c_uint8)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_f32p = ctypes.POINTER(ctypes.c_float)


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc only) if the .so is missing or stale."""
    src = os.path.join(_HERE, "pbx_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "clean", "all"])
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.pbx_oracle_cosine_distance.restype = ctypes.c_float
        L.pbx_oracle_cosine_distance.argtypes = [_u8p, ctypes.c_size_t, _u8p, ctypes.c_size_t]
        L.pbx_oracle_cosine_similarity_f32.restype = ctypes.c_float
        L.pbx_oracle_cosine_similarity_f32.argtypes = [_u8p, _u8p, ctypes.c_size_t]
        L.pbx_oracle_udf_cosine_distance.restype = ctypes.c_double
        L.pbx_oracle_udf_cosine_distance.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
        L.pbx_oracle_int_terms.restype = None
        L.pbx_oracle_int_terms.argtypes = [_u8p, _u8p, ctypes.c_size_t, _i32p, _i32p, _i32p]
        L.pbx_oracle_cosine_exact.restype = ctypes.c_double
        L.pbx_oracle_cosine_exact.argtypes = [_u8p, _u8p, ctypes.c_size_t]
        L.pbx_oracle_topk.restype = ctypes.c_int
        L.pbx_oracle_topk.argtypes = [_u8p, _i64p, ctypes.c_uint64, ctypes.c_uint32, _u8p, ctypes.c_uint32,
                                      ctypes.c_double, ctypes.c_int, _i64p, _f32p, _i32p, _i32p, _u32p]
        L.pbx_oracle_all_distances.restype = None
        L.pbx_oracle_all_distances.argtypes = [_u8p, ctypes.c_uint64, ctypes.c_uint32, _u8p, _f32p]
        L.pbx_oracle_byte_distance.restype = ctypes.c_float
        L.pbx_oracle_byte_distance.argtypes = [_u8p, _u8p, ctypes.c_size_t]
        L.pbx_oracle_hamming_distance.restype = ctypes.c_float
        L.pbx_oracle_hamming_distance.argtypes = [_u8p, _u8p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32)]
        L.pbx_oracle_quantize.restype = ctypes.c_uint8
        L.pbx_oracle_quantize.argtypes = [ctypes.c_float]
        L.pbx_oracle_synth_rows.restype = None
        L.pbx_oracle_synth_rows.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, _u8p]
        L.pbx_oracle_max_threads.restype = ctypes.c_int
        L.pbx_oracle_max_threads.argtypes = []
        _lib = L
    return _lib


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint8))


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def cosine_distance(a, b) -> np.float32:
    """Literal src/engine.rs:572-588 (different lengths allowed, zip semantics)."""
    a, b = _u8(a).ravel(), _u8(b).ravel()
    return np.float32(lib().pbx_oracle_cosine_distance(_p(a, _u8p), a.size, _p(b, _u8p), b.size))


def cosine_similarity_f32(a, b) -> np.float32:
    a, b = _u8(a).ravel(), _u8(b).ravel()
    assert a.size == b.size
    return np.float32(lib().pbx_oracle_cosine_similarity_f32(_p(a, _u8p), _p(b, _u8p), a.size))


def udf_cosine_distance(lhs: bytes, rhs: bytes) -> float:
    """The SQLite UDF body, src/engine.rs:613-620 (returns f64)."""
    return float(lib().pbx_oracle_udf_cosine_distance(lhs, len(lhs), rhs, len(rhs)))


def int_terms(q, r):
    q, r = _u8(q).ravel(), _u8(r).ravel()
    assert q.size == r.size
    dot, nq, nr = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    lib().pbx_oracle_int_terms(_p(q, _u8p), _p(r, _u8p), q.size, ctypes.byref(dot), ctypes.byref(nq), ctypes.byref(nr))
    return dot.value, nq.value, nr.value


def cosine_exact(q, r) -> float:
    q, r = _u8(q).ravel(), _u8(r).ravel()
    return float(lib().pbx_oracle_cosine_exact(_p(q, _u8p), _p(r, _u8p), q.size))


def topk(corpus, ids, query, k: int, max_dist: float = 1e3, threads: int = 1):
    """Bare-loop restatement of the similarity query (src/engine.rs:375-383).

    Returns (ids[int64], dist[float32], dot[int32], norm2[int32]) of the <=k rows with
    dist < max_dist, ordered by (dist asc, image_id asc).
    """
    corpus = _u8(corpus)
    n, d = corpus.shape
    query = _u8(query).ravel()
    assert query.size == d
    ids_arr = None if ids is None else np.ascontiguousarray(np.asarray(ids, dtype=np.int64))
    o_ids = np.zeros(k, np.int64)
    o_dist = np.zeros(k, np.float32)
    o_dot = np.zeros(k, np.int32)
    o_n2 = np.zeros(k, np.int32)
    cnt = ctypes.c_uint32(0)
    rc = lib().pbx_oracle_topk(_p(corpus, _u8p), None if ids_arr is None else _p(ids_arr, _i64p), n, d,
                               _p(query, _u8p), k, float(max_dist), int(threads),
                               _p(o_ids, _i64p), _p(o_dist, _f32p), _p(o_dot, _i32p), _p(o_n2, _i32p), ctypes.byref(cnt))
    if rc != 0:
        raise MemoryError("pbx_oracle_topk")
    c = cnt.value
    return o_ids[:c], o_dist[:c], o_dot[:c], o_n2[:c]


def all_distances(corpus, query) -> np.ndarray:
    corpus = _u8(corpus)
    n, d = corpus.shape
    query = _u8(query).ravel()
    out = np.zeros(n, np.float32)
    lib().pbx_oracle_all_distances(_p(corpus, _u8p), n, d, _p(query, _u8p), _p(out, _f32p))
    return out


def byte_distance(a, b) -> np.float32:
    """src/engine.rs:590-592."""
    a, b = _u8(a).ravel(), _u8(b).ravel()
    assert a.size == b.size
    return np.float32(lib().pbx_oracle_byte_distance(_p(a, _u8p), _p(b, _u8p), a.size))


def hamming_distance(a, b):
    """src/engine.rs:594-604 with the u8 sum wrapping as in a release build; returns (f32 distance, true differing bits)."""
    a, b = _u8(a).ravel(), _u8(b).ravel()
    assert a.size == b.size
    bits = ctypes.c_uint32(0)
    d = lib().pbx_oracle_hamming_distance(_p(a, _u8p), _p(b, _u8p), a.size, ctypes.byref(bits))
    return np.float32(d), int(bits.value)


def quantize(f) -> np.ndarray:
    """src/image_hashes/efficientnet.rs:39."""
    f = np.asarray(f, dtype=np.float32).ravel()
    return np.array([lib().pbx_oracle_quantize(float(x)) for x in f], dtype=np.uint8)


def synth_rows(seed: int, first_row: int, nrows: int, d: int) -> np.ndarray:
    out = np.zeros((nrows, d), np.uint8)
    lib().pbx_oracle_synth_rows(seed, first_row, nrows, d, _p(out, _u8p))
    return out


def max_threads() -> int:
    return int(lib().pbx_oracle_max_threads())
