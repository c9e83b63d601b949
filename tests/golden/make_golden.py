"""Generate tests/golden/golden_v1.npz from the UNMODIFIED reference compiled in place (oracle/_ref, built by
`make -C oracle ref` from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

The .npz holds inputs (raw sequences) and the reference's own outputs for them:
  * Loader<T>::get_point histograms / 1-mers / mag / length / stddev   (k=5 u8, k=3 u16, k=2 u32, k=4 u64, k=6 u8)
  * all 11 in-scope raw singles for a fixed pair list                    (Feature<T>::xxx static functions)
  * Trainer::classify scores, first-combo "dist", close flags, normalised caches for two weight files
  * Trainer::get_close / filter / merge results
  * DivergencePoint::distance / distance_d
Reference build flags are recorded in the file (oracle/Makefile REFFLAGS).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from meshclust2_b200 import synth  # noqa: E402

FLAGS = {"manhattan": 1 << 2, "euclidean": 1 << 3, "normalized_vectors": 1 << 5, "jefferey_divergence": 1 << 7,
         "pearson": 1 << 9, "intersection": 1 << 13, "emd": 1 << 18, "length_difference": 1 << 21,
         "kulczynski2": 1 << 27, "simratio": 1 << 28, "jensen_shannon": 1 << 29}


def main():
    assert ref.available(), "build oracle/_ref first: make -C oracle ref"
    rng = np.random.default_rng(20261017)
    seqs, tids = synth.make_set(96, 1000, 8, 0.10, seed=99)
    # adversarial additions: N runs (short bridged, long splitting), IUPAC, lower case, tiny sequences
    extra = [
        b"ACGTACGTTTGACCANNNACGTGGGTACCATGACGTACGATCGATCGTAGCTAGCTAGCATCGATCGAT",
        b"ACGTACGATTGACCACGTGGGTACCTTGACGTACGATCGATCCTAGCTAGGTAGCATCGATCGTT",
        seqs[0][:300] + b"N" * 9 + seqs[0][300:600] + b"N" * 10 + seqs[0][600:],
        seqs[1][:100].lower() + b"RYMKSWHBVD" + seqs[1][100:400] + b"N" * 55 + seqs[1][400:415] + b"N" * 30 + seqs[1][415:],
        b"A" * 700,                       # saturates u8 at k<=3
        b"ACGTACGTACGTACGTACGTAN",        # trailing N
        b"NNNNNNNNNNACGTACGTACGTACGTACGTACGTACGTACGTNA",  # lone trailing base after N (quirk Q7)
        b"ACGT",
        b"ACGTTGCAACGTTGCAAC",            # <= 20: no merging
    ]
    seqs = list(seqs) + extra
    n = len(seqs)
    out = {"n": n, "ref_flags": "-fopenmp -O3 -march=x86-64-v3 -std=c++11 -include cstdint -include limits",
           "ref_version": "MeShClust2 2.3.0 (/root/reference)"}
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    out["text"] = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    out["text_off"] = off
    for k, eb in ((5, 1), (3, 2), (2, 4), (4, 8), (6, 1), (3, 1)):
        pts = [ref.get_point(s, k, eb) for s in seqs]
        tag = "k%d_eb%d" % (k, eb)
        out["hist_" + tag] = np.stack([p["hist"] for p in pts])
        out["mers1_" + tag] = np.stack([p["mers1"] for p in pts])
        out["mag_" + tag] = np.array([p["mag"] for p in pts], dtype=np.uint64)
        out["len_" + tag] = np.array([p["len"] for p in pts], dtype=np.uint64)
        out["stddev_" + tag] = np.array([p["stddev"] for p in pts])
    # codes + segments of every sequence (ChromosomeOneDigit state)
    codes, segs, seg_off, eff = [], [], [0], []
    for s in seqs:
        c, sg, e = ref.encode(s)
        codes.append(c)
        segs.append(sg.reshape(-1, 2))
        seg_off.append(seg_off[-1] + len(sg))
        eff.append(e)
    out["codes"] = np.concatenate(codes)
    out["segs"] = np.concatenate(segs).astype(np.int32)
    out["seg_off"] = np.array(seg_off, dtype=np.int64)
    out["eff"] = np.array(eff, dtype=np.int64)

    # raw singles on a fixed pair list, per width
    m = 160
    ia = rng.integers(0, n, m)
    ib = rng.integers(0, n, m)
    out["pair_ia"], out["pair_ib"] = ia, ib
    names = list(FLAGS)
    out["single_names"] = np.array(names)
    for k, eb in ((5, 1), (3, 2), (2, 4), (4, 8)):
        tag = "k%d_eb%d" % (k, eb)
        H, ln = out["hist_" + tag], out["len_" + tag]
        raw = np.full((m, len(names)), np.nan)
        for j in range(m):
            for c, nm in enumerate(names):
                la, lb = int(ln[ia[j]]), int(ln[ib[j]])
                if nm == "length_difference" and (la == 0 or lb == 0):
                    continue
                raw[j, c] = ref.raw_single(FLAGS[nm], H[ia[j]], H[ib[j]], 0, 0, la, lb, k=k)
        out["raw_" + tag] = raw
    # stale pseudo-magnitudes (quirk Q4) on the u8 k=5 set
    H, ln, mag = out["hist_k5_eb1"], out["len_k5_eb1"], out["mag_k5_eb1"].copy()
    mag[::3] += 17
    mag[1::7] -= 5
    out["stale_mag_k5_eb1"] = mag
    raw = np.full((m, len(names)), np.nan)
    for j in range(m):
        for c, nm in enumerate(names):
            la, lb = int(ln[ia[j]]), int(ln[ib[j]])
            if nm == "length_difference" and (la == 0 or lb == 0):
                continue
            raw[j, c] = ref.raw_single(FLAGS[nm], H[ia[j]], H[ib[j]], int(mag[ia[j]]), int(mag[ib[j]]), la, lb, k=5)
    out["raw_stale_k5_eb1"] = raw

    # classifier outputs for the two pinned weight files (only sequences with non-zero length are scorable)
    # (a histogram of all ones has zero variance -> pearson NaN -> the reference throws; kept out of the scored set)
    ok = np.nonzero(ln >= 30)[0]
    ja = ok[rng.integers(0, len(ok), 400)]
    jb = ok[rng.integers(0, len(ok), 400)]
    out["score_ia"], out["score_ib"] = ja, jb
    for wname in ("weights_cfg1_id90", "weights_appendixD_id90"):
        txt = open(os.path.join(ROOT, "tests", "golden", wname + ".txt")).read()
        rm = ref.RefModel(txt, 1, 0.9)
        ns = 5
        r = rm.score_pairs(H, None, ln, ja, jb, mode=0, n_singles=ns)
        out[wname + "_score"], out[wname + "_dist"] = r["score"], r["dist"]
        out[wname + "_close"], out[wname + "_cache"] = r["close"], r["cache"]
        r2 = rm.score_pairs(H, out["stale_mag_k5_eb1"], ln, ja, jb, mode=0, n_singles=ns)
        out[wname + "_stale_score"], out[wname + "_stale_close"] = r2["score"], r2["close"]
        # get_close / filter / merge on a few queries
        gq, gbest, gdist, gmin, gmarks, fkeep, mres = [], [], [], [], [], [], []
        cand = ok
        for q in ok[::9]:
            b, d, ismin, marks = rm.get_close(H, None, ln, int(q), cand)
            gq.append(q), gbest.append(b), gdist.append(d), gmin.append(ismin), gmarks.append(marks)
            fkeep.append(rm.filter_members(H, None, ln, int(q), cand))
            rows = cand[(np.arange(8) + int(q)) % len(cand)]
            mres.append(rm.merge(H, None, ln, rows, 0, 1, 7))
        out[wname + "_gc_q"] = np.array(gq)
        out[wname + "_gc_best"] = np.array(gbest)
        out[wname + "_gc_dist"] = np.array(gdist)
        out[wname + "_gc_ismin"] = np.array(gmin)
        out[wname + "_gc_marks"] = np.stack(gmarks)
        out[wname + "_filter_keep"] = np.stack(fkeep)
        out[wname + "_merge"] = np.array(mres)
    out["cand"] = ok
    # distance / distance_d
    out["distance_k5_eb1"] = np.array([ref.distance(H[a], H[b]) for a, b in zip(ia, ib)], dtype=np.uint64)
    centers = (H[ia[:40]].astype(np.float64) + H[ib[:40]].astype(np.float64) + H[ia[40:80]].astype(np.float64)) / 3.0
    out["distance_d_centers"] = centers
    out["distance_d_k5_eb1"] = np.array([ref.distance_d(H[ib[j]], centers[j]) for j in range(40)])
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
