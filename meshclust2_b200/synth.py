"""Seeded synthetic mutated-template DNA for the BASELINE.json configs.

templates = i.i.d. uniform ACGT of length L +/- 5 %; variants = template with per-base event rate
r ~ U(0, r_max), split 80 % substitution / 10 % deletion / 10 % insertion.
Everything is numpy `default_rng(seed)`; sequences are returned as lists of ASCII `bytes`.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_template(rng, length):
    return _ACGT[rng.integers(0, 4, size=int(length))]


def mutate(rng, tmpl, rate):
    """Point-mutate a uint8 ASCII array: 80 % substitutions, 10 % deletions, 10 % insertions."""
    n = tmpl.size
    ev = rng.random(n) < rate
    kind = rng.random(n)
    sub = ev & (kind < 0.8)
    dele = ev & (kind >= 0.8) & (kind < 0.9)
    ins = ev & (kind >= 0.9)
    out = tmpl.copy()
    # substitution: shift to a different base
    codes = np.zeros(256, dtype=np.uint8)
    codes[_ACGT] = np.arange(4, dtype=np.uint8)
    c = codes[out[sub]]
    out[sub] = _ACGT[(c + rng.integers(1, 4, size=c.size)) % 4]
    keep = ~dele
    # insertion: add one random base after the position
    reps = keep.astype(np.int64) + ins.astype(np.int64)
    res = np.repeat(out, reps)
    if ins.any():
        # positions of the inserted copies = last copy of each run where ins & keep, or the only copy where ins & ~keep
        ends = np.cumsum(reps) - 1
        pos = ends[ins]
        res[pos] = _ACGT[rng.integers(0, 4, size=pos.size)]
    return res


def make_set(n, length, n_templates, r_max, seed, len_jitter=0.05, ancestor_div=None, variant_rmax=None):
    """Return (list[bytes] sequences, template_id int32[n]).

    ancestor_div=(lo,hi): templates are derived from ONE ancestor at lo..hi divergence (16S-like, cfg2);
    then variants at 0..variant_rmax (defaults to r_max)."""
    rng = np.random.default_rng(seed)
    if ancestor_div is None:
        lens = np.round(length * (1 + len_jitter * (2 * rng.random(n_templates) - 1))).astype(np.int64)
        templates = [random_template(rng, L) for L in lens]
    else:
        anc = random_template(rng, length)
        lo, hi = ancestor_div
        templates = [mutate(rng, anc, lo + (hi - lo) * rng.random()) for _ in range(n_templates)]
    vr = r_max if variant_rmax is None else variant_rmax
    seqs, tids = [], np.zeros(n, dtype=np.int32)
    for i in range(n):
        t = i % n_templates
        tids[i] = t
        seqs.append(mutate(rng, templates[t], vr * rng.random()).tobytes())
    return seqs, tids


CONFIGS = {
    # name: kwargs for make_set + k + elem_bytes
    "cfg1": dict(n=2000, length=1000, n_templates=300, r_max=0.08, seed=7, k=5, elem_bytes=1),
    "cfg2": dict(n=10000, length=1500, n_templates=200, r_max=0.03, seed=42, ancestor_div=(0.05, 0.25), k=5,
                 elem_bytes=1),
    "cfg3": dict(n=100000, length=1000, n_templates=1000, r_max=0.08, seed=3, k=5, elem_bytes=1),
    "cfg5": dict(n=1000000, length=1000, n_templates=10000, r_max=0.08, seed=5, k=5, elem_bytes=1),
}


BLOCK = 1024


def make_range(n, length, n_templates, r_max, seed, lo=0, hi=None, len_jitter=0.05, ancestor_div=None,
               variant_rmax=None):
    """Sequences [lo, hi) of an n-sequence set, generated block-wise so that any rank can produce its own shard and
    the union over ranks is independent of how the range was split: templates come from `seed`, the variants of
    block b (sequences b*1024 .. b*1024+1023) from default_rng([seed, 1 + b])."""
    hi = n if hi is None else hi
    rng = np.random.default_rng([seed, 0])
    if ancestor_div is None:
        lens = np.round(length * (1 + len_jitter * (2 * rng.random(n_templates) - 1))).astype(np.int64)
        templates = [random_template(rng, L) for L in lens]
    else:
        anc = random_template(rng, length)
        a, b = ancestor_div
        templates = [mutate(rng, anc, a + (b - a) * rng.random()) for _ in range(n_templates)]
    vr = r_max if variant_rmax is None else variant_rmax
    seqs, tids = [], []
    for blk in range(lo // BLOCK, (hi + BLOCK - 1) // BLOCK):
        brng = np.random.default_rng([seed, 1 + blk])
        for i in range(blk * BLOCK, min((blk + 1) * BLOCK, n)):
            t = i % n_templates
            s = mutate(brng, templates[t], vr * brng.random())
            if lo <= i < hi:
                seqs.append(s.tobytes())
                tids.append(t)
    return seqs, np.array(tids, dtype=np.int32)


def make_config_range(name, lo=0, hi=None, n=None):
    cfg = dict(CONFIGS[name])
    k, eb = cfg.pop("k"), cfg.pop("elem_bytes")
    if n is not None:
        cfg["n"] = n
    seqs, tids = make_range(lo=lo, hi=hi, **cfg)
    return seqs, tids, k, eb


def make_config(name, n=None):
    cfg = dict(CONFIGS[name])
    k, eb = cfg.pop("k"), cfg.pop("elem_bytes")
    if n is not None:
        cfg["n"] = n
    seqs, tids = make_set(**cfg)
    return seqs, tids, k, eb


def make_single_file(n_files, contigs, contig_len, seed, n_templates=None):
    """cfg4: each record = `contigs` contigs of `contig_len` joined by 50 N, as `--single-file` does
    (ChromListMaker.cpp:138-142)."""
    rng = np.random.default_rng(seed)
    n_templates = n_templates or max(1, n_files // 50)
    templ = [[random_template(rng, contig_len) for _ in range(contigs)] for _ in range(n_templates)]
    gap = np.full(50, ord("N"), dtype=np.uint8)
    out = []
    for i in range(n_files):
        t = templ[i % n_templates]
        parts = []
        r = 0.05 * rng.random()
        for j, c in enumerate(t):
            if j:
                parts.append(gap)
            parts.append(mutate(rng, c, r))
        out.append(np.concatenate(parts).tobytes())
    return out


def to_fasta(seqs, tids=None, width=70):
    lines = []
    for i, s in enumerate(seqs):
        hdr = ">seq%d" % i + ("" if tids is None else " template_%d" % tids[i])
        lines.append(hdr)
        s = s.decode()
        lines.extend(s[j:j + width] for j in range(0, len(s), width))
    return "\n".join(lines) + "\n"
