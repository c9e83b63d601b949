"""ctypes view of oracle/libmc2oracle.so (mc2_oracle.c) — TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmc2oracle.so")

MAX_SINGLES, MAX_COMBOS, MAX_COMBO_IDX = 16, 16, 4

FEAT = {
    "manhattan": 1 << 2,
    "euclidean": 1 << 3,
    "normalized_vectors": 1 << 5,
    "jefferey_divergence": 1 << 7,
    "pearson": 1 << 9,
    "intersection": 1 << 13,
    "emd": 1 << 18,
    "length_difference": 1 << 21,
    "kulczynski2": 1 << 27,
    "simratio": 1 << 28,
    "jensen_shannon": 1 << 29,
}
FAST = ["manhattan", "euclidean", "normalized_vectors", "pearson", "intersection", "emd", "length_difference",
        "kulczynski2", "simratio"]
SLOW = FAST + ["jefferey_divergence", "jensen_shannon"]
DTYPES = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}


class CModel(C.Structure):
    _fields_ = [
        ("n_singles", C.c_int),
        ("single_flag", C.c_uint64 * MAX_SINGLES),
        ("single_min", C.c_double * MAX_SINGLES),
        ("single_max", C.c_double * MAX_SINGLES),
        ("n_combos", C.c_int),
        ("combo_kind", C.c_int * MAX_COMBOS),
        ("combo_nidx", C.c_int * MAX_COMBOS),
        ("combo_idx", (C.c_int * MAX_COMBO_IDX) * MAX_COMBOS),
        ("weight", C.c_double * (MAX_COMBOS + 1)),
        ("bias", C.c_double),
    ]


class CPoint(C.Structure):
    _fields_ = [("bins", C.c_void_p), ("mag", C.c_uint64), ("len", C.c_uint64)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.mc2o_distance.restype = C.c_uint64
        _lib.mc2o_distance_d.restype = C.c_double
        _lib.mc2o_strip_acgt.restype = C.c_long
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Model:
    """Feature set + GLM weights in the reference's weights.txt terms (Predictor.cpp:82-185)."""

    def __init__(self, singles, combos, weights, bias=0.0, k=None, ident=None, datatype=None, mode=1,
                 feature_set=405021228, max_features=4):
        # singles: list of (flag, min, max); combos: list of (kind_code, flags); weights: C+1 doubles
        self.singles = [(int(f), float(a), float(b)) for f, a, b in singles]
        self.combos = [(int(kc), int(fl)) for kc, fl in combos]
        self.weights = [float(w) for w in weights]
        self.bias = float(bias)
        self.k, self.ident, self.datatype, self.mode = k, ident, datatype, mode
        self.feature_set, self.max_features = feature_set, max_features
        assert len(self.weights) == len(self.combos) + 1

    def lookup(self):
        return [f for f, _, _ in self.singles]

    def combo_indices(self):
        look = self.lookup()
        out = []
        for _, flags in self.combos:
            idx = [look.index(1 << b) for b in range(64) if flags >> b & 1]  # ascending flag bit
            out.append(idx)
        return out

    def cmodel(self):
        m = CModel()
        m.n_singles = len(self.singles)
        for i, (f, lo, hi) in enumerate(self.singles):
            m.single_flag[i], m.single_min[i], m.single_max[i] = f, lo, hi
        m.n_combos = len(self.combos)
        for c, ((kind, _), idx) in enumerate(zip(self.combos, self.combo_indices())):
            m.combo_kind[c] = kind
            m.combo_nidx[c] = len(idx)
            for j, ix in enumerate(idx):
                m.combo_idx[c][j] = ix
        for i, w in enumerate(self.weights):
            m.weight[i] = w
        m.bias = self.bias
        return m

    # --- weights.txt (writer mirrors Predictor::save / write_to; reader mirrors the file ctor / read_from)
    def to_text(self):
        def g(x):
            return "%.15g" % x
        s = "k: %d\nmode: %d\nmax_features: %d\nID: %s\nDatatype: %s\nfeature_set: %d\n" % (
            self.k, self.mode, self.max_features, g(self.ident), self.datatype, self.feature_set)
        s += "\nn_combos: %d\n%s\n" % (len(self.combos), g(self.weights[0]))
        for (kind, flags), w in zip(self.combos, self.weights[1:]):
            s += "%d %d %s\n" % (kind, flags, g(w))
        s += "\nn_singles: %d\n" % len(self.singles)
        for f, lo, hi in self.singles:
            s += "%d %s %s\n" % (f, g(lo), g(hi))
        return s

    @staticmethod
    def from_text(text):
        tok = text.split()
        it = iter(tok)

        def nxt():
            return next(it)
        hdr = {}
        for key in ("k:", "mode:", "max_features:", "ID:", "Datatype:", "feature_set:"):
            assert nxt() == key, key
            hdr[key] = nxt()
        assert nxt() == "n_combos:"
        nc = int(nxt())
        weights = [float(nxt())]
        combos = []
        lookup = []
        for _ in range(nc):
            kind, flags, w = int(nxt()), int(nxt()), float(nxt())
            combos.append((kind, flags))
            weights.append(w)
            for b in range(64):  # add_feature order, Feature.cpp:102-127
                if flags >> b & 1 and (1 << b) not in lookup:
                    lookup.append(1 << b)
        assert nxt() == "n_singles:"
        ns = int(nxt())
        norm = {}
        for _ in range(ns):
            f, lo, hi = int(nxt()), float(nxt()), float(nxt())
            norm[f] = (lo, hi)
        singles = [(f,) + norm[f] for f in lookup]
        return Model(singles, combos, weights, 0.0, int(hdr["k:"]), float(hdr["ID:"]), hdr["Datatype:"],
                     int(hdr["mode:"]), int(hdr["feature_set:"]), int(hdr["max_features:"]))


def encode(text):
    """raw sequence text (bytes) -> (codes int8[len], segs int32[nseg,2], effective_size)"""
    n = len(text)
    base = np.zeros(max(n, 1), dtype=np.int8)
    max_segs = n // 2 + 2
    segs = np.zeros((max_segs, 2), dtype=np.int32)
    nseg, eff = C.c_int(), C.c_long()
    rc = lib().mc2o_encode(text, C.c_long(n), _p(base), _p(segs), max_segs, C.byref(nseg), C.byref(eff))
    if rc != 0:
        raise ValueError("mc2o_encode rc=%d" % rc)
    return base[:n], segs[:nseg.value].copy(), eff.value


def count(codes, segs, k, elem_bytes):
    N = 4 ** k
    hist = np.zeros(N, dtype=DTYPES[elem_bytes])
    m1 = np.zeros(4, dtype=np.uint64)
    novf = C.c_int()
    segs = np.ascontiguousarray(segs, dtype=np.int32)
    codes = np.ascontiguousarray(codes, dtype=np.int8)
    rc = lib().mc2o_count(_p(codes), _p(segs), len(segs), k, elem_bytes, _p(hist), _p(m1), C.byref(novf))
    if rc != 0:
        raise ValueError("mc2o_count rc=%d" % rc)
    return hist, m1, novf.value


def largest_count(codes, segs, k):
    """Runner::run's width detection for one sequence: 1 + largest k-mer multiplicity (u64 table, init 1)"""
    out = C.c_uint64()
    segs = np.ascontiguousarray(segs, dtype=np.int32)
    codes = np.ascontiguousarray(codes, dtype=np.int8)
    rc = lib().mc2o_largest_count(_p(codes), _p(segs), len(segs), k, C.byref(out))
    if rc != 0:
        raise ValueError("mc2o_largest_count rc=%d" % rc)
    return out.value


def width_for(largest):
    return lib().mc2o_width_for(C.c_uint64(largest))


def point_stats(hist):
    mag, sd = C.c_uint64(), C.c_double()
    lib().mc2o_point_stats(_p(hist), C.c_uint64(hist.size), hist.dtype.itemsize, C.byref(mag), C.byref(sd))
    return mag.value, sd.value


def get_point(text, k, elem_bytes):
    """Loader<T>::get_point(ChromosomeOneDigit*) on raw text -> dict"""
    codes, segs, eff = encode(text)
    hist, m1, novf = count(codes, segs, k, elem_bytes)
    mag, sd = point_stats(hist)
    return dict(hist=hist, mers1=m1, mag=mag, len=eff, stddev=sd, n_overflow=novf, codes=codes, segs=segs)


def _pt(h, mag, ln):
    return CPoint(h.ctypes.data, int(mag), int(ln))


def raw_single(flag, p, q, mag_p=None, mag_q=None, len_p=1, len_q=1):
    assert p.dtype == q.dtype and p.size == q.size
    mp = int(p.sum(dtype=np.uint64)) if mag_p is None else mag_p
    mq = int(q.sum(dtype=np.uint64)) if mag_q is None else mag_q
    a, b = _pt(p, mp, len_p), _pt(q, mq, len_q)
    out = C.c_double()
    rc = lib().mc2o_raw_single(C.c_uint64(flag), p.dtype.itemsize, C.c_uint64(p.size), C.byref(a), C.byref(b),
                               C.byref(out))
    if rc != 0:
        raise ValueError("mc2o_raw_single rc=%d" % rc)
    return out.value


def score_pairs(model, H, mag, ln, ia, ib, threads=1, want_cache=True):
    """Trainer::classify-style scoring of pairs (ia[j], ib[j]) -> dict(score, dist, close, cache, seconds)"""
    H = np.ascontiguousarray(H)
    n, N = H.shape
    mag = np.ascontiguousarray(mag, dtype=np.uint64)
    ln = np.ascontiguousarray(ln, dtype=np.uint64)
    ia = np.ascontiguousarray(ia, dtype=np.uint64)
    ib = np.ascontiguousarray(ib, dtype=np.uint64)
    m = ia.size
    cm = model.cmodel()
    score = np.zeros(m)
    dist = np.zeros(m)
    close = np.zeros(m, dtype=np.uint8)
    cache = np.zeros((m, cm.n_singles)) if want_cache else None
    sec = C.c_double()
    rc = lib().mc2o_score_pairs(C.byref(cm), H.dtype.itemsize, C.c_uint64(N), _p(H), _p(mag), _p(ln), C.c_uint64(m),
                                _p(ia), _p(ib), _p(score), _p(dist), _p(close), _p(cache), threads, C.byref(sec))
    if rc != 0:
        raise ValueError("mc2o_score_pairs rc=%d" % rc)
    return dict(score=score, dist=dist, close=close, cache=cache, seconds=sec.value)


def predict_pair(model, p, q, mag_p, mag_q, len_p, len_q):
    cm = model.cmodel()
    a, b = _pt(p, mag_p, len_p), _pt(q, mag_q, len_q)
    out = C.c_double()
    rc = lib().mc2o_predict_pair(C.byref(cm), p.dtype.itemsize, C.c_uint64(p.size), C.byref(a), C.byref(b), C.byref(out))
    if rc != 0:
        raise ValueError("mc2o_predict_pair rc=%d" % rc)
    return out.value


def get_close(model, H, mag, ln, q, cand, cutoff):
    H = np.ascontiguousarray(H)
    n, N = H.shape
    mag = np.ascontiguousarray(mag, dtype=np.uint64)
    ln = np.ascontiguousarray(ln, dtype=np.uint64)
    cand = np.ascontiguousarray(cand, dtype=np.uint64)
    cm = model.cmodel()
    best, bd, ismin = C.c_int64(), C.c_double(), C.c_int()
    marks = np.zeros(cand.size, dtype=np.uint8)
    rc = lib().mc2o_get_close(C.byref(cm), H.dtype.itemsize, C.c_uint64(N), _p(H), _p(mag), _p(ln), C.c_uint64(q),
                              C.c_uint64(cand.size), _p(cand), C.c_double(cutoff), C.byref(best), C.byref(bd),
                              C.byref(ismin), _p(marks))
    if rc != 0:
        raise ValueError("mc2o_get_close rc=%d" % rc)
    return best.value, bd.value, bool(ismin.value), marks


def filter_members(model, H, mag, ln, c, members, ident):
    H = np.ascontiguousarray(H)
    n, N = H.shape
    mag = np.ascontiguousarray(mag, dtype=np.uint64)
    ln = np.ascontiguousarray(ln, dtype=np.uint64)
    members = np.ascontiguousarray(members, dtype=np.uint64)
    cm = model.cmodel()
    keep = np.zeros(members.size, dtype=np.uint8)
    rc = lib().mc2o_filter(C.byref(cm), H.dtype.itemsize, C.c_uint64(N), _p(H), _p(mag), _p(ln), C.c_uint64(c),
                           C.c_uint64(members.size), _p(members), C.c_double(ident), _p(keep))
    if rc != 0:
        raise ValueError("mc2o_filter rc=%d" % rc)
    return keep


def merge(model, H, mag, ln, rows, cur, begin, last, ident):
    H = np.ascontiguousarray(H)
    n, N = H.shape
    mag = np.ascontiguousarray(mag, dtype=np.uint64)
    ln = np.ascontiguousarray(ln, dtype=np.uint64)
    rows = np.ascontiguousarray(rows, dtype=np.uint64)
    cm = model.cmodel()
    out = C.c_long()
    rc = lib().mc2o_merge(C.byref(cm), H.dtype.itemsize, C.c_uint64(N), _p(H), _p(mag), _p(ln), _p(rows),
                          C.c_long(cur), C.c_long(begin), C.c_long(last), C.c_double(ident), C.byref(out))
    if rc != 0:
        raise ValueError("mc2o_merge rc=%d" % rc)
    return out.value


def distance(p, q, mag_p=None, mag_q=None):
    mp = int(p.sum(dtype=np.uint64)) if mag_p is None else mag_p
    mq = int(q.sum(dtype=np.uint64)) if mag_q is None else mag_q
    a, b = _pt(p, mp, 1), _pt(q, mq, 1)
    return lib().mc2o_distance(p.dtype.itemsize, C.c_uint64(p.size), C.byref(a), C.byref(b))


def distance_d(p, center):
    center = np.ascontiguousarray(center, dtype=np.float64)
    return lib().mc2o_distance_d(p.dtype.itemsize, C.c_uint64(p.size), _p(p), _p(center))


def mean_closest(H, members):
    """K3: (best position in members, its distance_d, mean[N], dist[n])"""
    H = np.ascontiguousarray(H)
    n, N = H.shape
    members = np.ascontiguousarray(members, dtype=np.uint64)
    best, bd = C.c_int64(), C.c_double()
    mean = np.zeros(N)
    dist = np.zeros(members.size)
    rc = lib().mc2o_mean_closest(H.dtype.itemsize, C.c_uint64(N), _p(H), _p(members), C.c_uint64(members.size), C.byref(best),
                                 C.byref(bd), _p(mean), _p(dist))
    if rc != 0:
        raise ValueError("mc2o_mean_closest rc=%d" % rc)
    return best.value, bd.value, mean, dist


def count_batch(codes, seq_off, segs, seg_off, k, elem_bytes, threads=1):
    """codes: int8 concatenated; seq_off uint64[n+1]; segs int32[total,2] (sequence-relative); seg_off uint64[n+1]"""
    n = len(seq_off) - 1
    N = 4 ** k
    hist = np.zeros((n, N), dtype=DTYPES[elem_bytes])
    sec = C.c_double()
    codes = np.ascontiguousarray(codes, dtype=np.int8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
    segs = np.ascontiguousarray(segs, dtype=np.int32)
    seg_off = np.ascontiguousarray(seg_off, dtype=np.uint64)
    rc = lib().mc2o_count_batch(_p(codes), _p(seq_off), _p(segs), _p(seg_off), C.c_uint64(n), k, elem_bytes, _p(hist),
                                threads, C.byref(sec))
    if rc != 0:
        raise ValueError("mc2o_count_batch rc=%d" % rc)
    return hist, sec.value
