"""ctypes view of oracle/_ref/libmc2ref.so — the UNMODIFIED reference compiled from /root/reference
by oracle/Makefile (`make ref`).  TEST INFRASTRUCTURE ONLY.

available() is False where the library was not built (e.g. a box without /root/reference and without
a prebuilt oracle/_ref/); callers skip or fall back to oracle.port.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libmc2ref.so")
REF_ROOT = os.environ.get("MC2_REFERENCE_ROOT", "/root/reference")
DTYPES = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}


def build():
    """Compile the reference where it lies (needs REF_ROOT); outputs only into oracle/_ref/."""
    subprocess.check_call(["make", "-s", "-j8", "-C", _HERE, "ref", "REF=" + REF_ROOT])


def available():
    return os.path.exists(_LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_LIB)
        _lib.ref_model_load.restype = C.c_void_p
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def max_threads():
    return lib().ref_max_threads()


def encode(text):
    n = len(text)
    base = np.zeros(max(n, 1), dtype=np.int8)
    max_segs = n // 2 + 2
    segs = np.zeros((max_segs, 2), dtype=np.int32)
    nseg, eff = C.c_int(), C.c_long()
    rc = lib().ref_encode(text, C.c_long(n), _p(base), _p(segs), max_segs, C.byref(nseg), C.byref(eff))
    if rc != 0:
        raise ValueError("ref_encode rc=%d" % rc)
    return base[:n], segs[:nseg.value].copy(), eff.value


def get_point(text, k, elem_bytes):
    N = 4 ** k
    h = np.zeros(N, dtype=DTYPES[elem_bytes])
    m1 = np.zeros(4, dtype=np.uint64)
    mag, ln, sd = C.c_uint64(), C.c_uint64(), C.c_double()
    rc = lib().ref_get_point(text, C.c_long(len(text)), k, elem_bytes, _p(h), _p(m1), C.byref(mag), C.byref(ln),
                             C.byref(sd))
    if rc != 0:
        raise ValueError("ref_get_point rc=%d" % rc)
    return dict(hist=h, mers1=m1, mag=mag.value, len=ln.value, stddev=sd.value)


def kmer_table(codes, first, last, k, elem_bytes, init=1):
    N = 4 ** k
    v = np.zeros(N, dtype=DTYPES[elem_bytes])
    ret = C.c_int()
    codes = np.ascontiguousarray(codes, dtype=np.int8)
    rc = lib().ref_kmer_table(_p(codes), first, last, k, elem_bytes, C.c_uint64(init), _p(v), C.byref(ret))
    if rc != 0:
        raise ValueError("ref_kmer_table rc=%d" % rc)
    return v, ret.value


def raw_single(flag, p, q, mag_p=0, mag_q=0, len_p=1, len_q=1, k=1):
    """mag_* = 0 -> the reference sums the bins itself; else a stale pseudo-magnitude (quirk Q4)."""
    out = C.c_double()
    rc = lib().ref_raw_single(C.c_uint64(flag), p.dtype.itemsize, k, C.c_uint64(p.size), _p(p), _p(q),
                              C.c_uint64(mag_p), C.c_uint64(mag_q), C.c_uint64(len_p), C.c_uint64(len_q),
                              C.byref(out))
    if rc != 0:
        raise ValueError("ref_raw_single rc=%d" % rc)
    return out.value


def distance(p, q, mag_p=0, mag_q=0):
    out = C.c_uint64()
    rc = lib().ref_distance(p.dtype.itemsize, C.c_uint64(p.size), _p(p), _p(q), C.c_uint64(mag_p), C.c_uint64(mag_q),
                            C.byref(out))
    if rc != 0:
        raise ValueError("ref_distance rc=%d" % rc)
    return out.value


def distance_d(p, center):
    center = np.ascontiguousarray(center, dtype=np.float64)
    out = C.c_double()
    rc = lib().ref_distance_d(p.dtype.itemsize, C.c_uint64(p.size), _p(p), _p(center), C.byref(out))
    if rc != 0:
        raise ValueError("ref_distance_d rc=%d" % rc)
    return out.value


class RefModel:
    """Predictor<T>(weights file) + a Trainer<T> carrying its classifier (never destroyed)."""

    def __init__(self, weights_text, elem_bytes, cutoff):
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            f.write(weights_text)
            path = f.name
        try:
            self.h = lib().ref_model_load(path.encode(), elem_bytes, C.c_double(cutoff))
        finally:
            os.unlink(path)
        if not self.h:
            raise ValueError("ref_model_load failed")
        self.elem_bytes = elem_bytes
        self.n_singles = None

    def score_pairs(self, H, mag, ln, ia, ib, mode=0, threads=1, n_singles=0):
        H = np.ascontiguousarray(H)
        n, N = H.shape
        mag = None if mag is None else np.ascontiguousarray(mag, dtype=np.uint64)
        ln = np.ascontiguousarray(ln, dtype=np.uint64)
        ia = np.ascontiguousarray(ia, dtype=np.uint64)
        ib = np.ascontiguousarray(ib, dtype=np.uint64)
        m = ia.size
        score = np.zeros(m)
        dist = np.zeros(m)
        close = np.zeros(m, dtype=np.uint8)
        cache = np.zeros((m, n_singles)) if n_singles else None
        sec = C.c_double()
        rc = lib().ref_score_pairs(C.c_void_p(self.h), mode, C.c_uint64(N), C.c_uint64(n), _p(H), _p(mag), _p(ln),
                                   C.c_uint64(m), _p(ia), _p(ib), _p(score), _p(dist), _p(close), _p(cache), threads,
                                   C.byref(sec))
        if rc != 0:
            raise ValueError("ref_score_pairs rc=%d" % rc)
        return dict(score=score, dist=dist, close=close, cache=cache, seconds=sec.value)

    def get_close(self, H, mag, ln, q, cand, threads=1):
        H = np.ascontiguousarray(H)
        n, N = H.shape
        mag = None if mag is None else np.ascontiguousarray(mag, dtype=np.uint64)
        ln = np.ascontiguousarray(ln, dtype=np.uint64)
        cand = np.ascontiguousarray(cand, dtype=np.uint64)
        best, bd, ismin = C.c_int64(), C.c_double(), C.c_int()
        marks = np.zeros(cand.size, dtype=np.uint8)
        rc = lib().ref_get_close(C.c_void_p(self.h), C.c_uint64(N), _p(H), _p(mag), _p(ln), C.c_uint64(q),
                                 C.c_uint64(cand.size), _p(cand), C.byref(best), C.byref(bd), C.byref(ismin),
                                 _p(marks), threads)
        if rc != 0:
            raise ValueError("ref_get_close rc=%d" % rc)
        return best.value, bd.value, bool(ismin.value), marks

    def filter_members(self, H, mag, ln, c, members):
        H = np.ascontiguousarray(H)
        n, N = H.shape
        mag = None if mag is None else np.ascontiguousarray(mag, dtype=np.uint64)
        ln = np.ascontiguousarray(ln, dtype=np.uint64)
        members = np.ascontiguousarray(members, dtype=np.uint64)
        keep = np.zeros(members.size, dtype=np.uint8)
        rc = lib().ref_filter(C.c_void_p(self.h), C.c_uint64(N), _p(H), _p(mag), _p(ln), C.c_uint64(c),
                              C.c_uint64(members.size), _p(members), _p(keep))
        if rc != 0:
            raise ValueError("ref_filter rc=%d" % rc)
        return keep

    def merge(self, H, mag, ln, rows, cur, begin, last, threads=1):
        H = np.ascontiguousarray(H)
        n, N = H.shape
        mag = None if mag is None else np.ascontiguousarray(mag, dtype=np.uint64)
        ln = np.ascontiguousarray(ln, dtype=np.uint64)
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = C.c_long()
        rc = lib().ref_merge(C.c_void_p(self.h), C.c_uint64(N), _p(H), _p(mag), _p(ln), C.c_uint64(rows.size), _p(rows),
                             C.c_long(cur), C.c_long(begin), C.c_long(last), C.byref(out), threads)
        if rc != 0:
            raise ValueError("ref_merge rc=%d" % rc)
        return out.value


def mean_closest(H, members):
    H = np.ascontiguousarray(H)
    n, N = H.shape
    members = np.ascontiguousarray(members, dtype=np.uint64)
    best, bd = C.c_int64(), C.c_double()
    mean = np.zeros(N)
    dist = np.zeros(members.size)
    rc = lib().ref_mean_closest(H.dtype.itemsize, C.c_uint64(N), _p(H), _p(members), C.c_uint64(members.size), C.byref(best),
                                C.byref(bd), _p(mean), _p(dist))
    if rc != 0:
        raise ValueError("ref_mean_closest rc=%d" % rc)
    return best.value, bd.value, mean, dist


def largest_count(texts, k):
    """Runner::run's width detection over raw sequences (list of bytes) -> Largest count"""
    n = len(texts)
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(t) for t in texts])
    out = C.c_uint64()
    rc = lib().ref_largest_count(b"".join(texts), _p(off), C.c_uint64(n), k, C.byref(out))
    if rc != 0:
        raise ValueError("ref_largest_count rc=%d" % rc)
    return out.value


def count_batch(texts, k, elem_bytes, threads=1, want_hist=True):
    """Loader<T>::get_point over raw sequences (list of bytes), omp over sequences -> (hist or None, seconds)"""
    n = len(texts)
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(t) for t in texts])
    blob = b"".join(texts)
    hist = np.zeros((n, 4 ** k), dtype=DTYPES[elem_bytes]) if want_hist else None
    sec = C.c_double()
    rc = lib().ref_count_batch(blob, _p(off), C.c_uint64(n), k, elem_bytes, _p(hist), threads, C.byref(sec))
    if rc != 0:
        raise ValueError("ref_count_batch rc=%d" % rc)
    return hist, sec.value
