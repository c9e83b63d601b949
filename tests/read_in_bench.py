"""Histogram stage of the relinked meshclust2 on a large single FASTA (default: the cfg3 shape, 100k x 1 kb): the reference's
own `read_in_points` timestamp (FASTA parse + one Loader<T>::get_point per sequence, serial for a single file) next to the
relinked binary's (same parse + ONE K1 batch on the device).  Both stop after training (--dump).
usage: python tests/read_in_bench.py [n_sequences] [threads]"""
import os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshclust2_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
threads = sys.argv[2] if len(sys.argv) > 2 else "16"
seqs, tids, k, eb = synth.make_config_range("cfg3", 0, n)
tmp = tempfile.mkdtemp()
fasta = os.path.join(tmp, "in.fa")
open(fasta, "w").write(synth.to_fasta(seqs, tids))
for name in ("meshclust2", "meshclust2_b200"):  # the relinked binary takes the device reader above MC2_K1_MIN_BASES (default 32 Mi)
    wd = os.path.join(tmp, name); os.makedirs(wd)
    t0 = time.time()
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", name), "--id", "0.9", "--threads", threads, "--sample", "300",
                        "--num-templates", "60", "--dump", os.path.join(wd, "w.txt"), fasta], cwd=wd, capture_output=True, text=True,
                       env=dict(os.environ, MC2_TIMING="1"))
    dt = time.time() - t0
    m = re.search(r"timestamp read_in_points ([0-9.]+)", r.stdout)
    for l in r.stderr.splitlines():
        if "timing" in l:
            print("    " + l)
    print("%-16s n=%d threads=%s: rc=%d read_in_points %s s, wall %.2f s" % (name, n, threads, r.returncode, m.group(1) if m else "?", dt))
    if r.returncode != 0:
        print(r.stdout[-800:], r.stderr[-800:])
