"""End-to-end meshclust2 on the BASELINE configs[0] / configs[1] shapes: the unmodified reference binary (oracle/_ref/
meshclust2) next to the same binary with src/cluster/Trainer.cpp swapped for integration/Trainer_b200.cpp (GPU through the C
ABI).  Prints wall-clock, the reference's own stage timestamps and whether the two CLSTR outputs hold the same clusters.
The relinked binary runs with the batched update stage (default) and, with a third argument "percall", also with
MC2_NO_BATCH=1 (one device call per center).
usage: python tests/e2e_cluster.py [cfg1|cfg2] [threads] [percall]"""
import os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshclust2_b200 import synth

def parse_clstr(path):
    clusters, cur = [], None
    for line in open(path):
        if line.startswith(">Cluster"):
            cur = set(); clusters.append(cur)
        else:
            m = re.search(r">(\S+)", line)
            if m:
                cur.add(m.group(1).rstrip("."))
    return {frozenset(c) for c in clusters if c}

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
threads = sys.argv[2] if len(sys.argv) > 2 else "1"
seqs, tids, k, eb = synth.make_config_range(cfg, 0, None)
tmp = tempfile.mkdtemp()
fasta = os.path.join(tmp, "in.fa")
open(fasta, "w").write(synth.to_fasta(seqs, tids))
extra = {"cfg1": ["--sample", "2000", "--num-templates", "300"], "cfg2": []}[cfg]
res = {}
arms = [("meshclust2", "meshclust2", {}), ("meshclust2_b200", "meshclust2_b200", {})]
if len(sys.argv) > 3 and sys.argv[3] == "percall":
    arms.append(("meshclust2_b200 percall", "meshclust2_b200", {"MC2_NO_BATCH": "1"}))
for name, binary, env in arms:
    wd = os.path.join(tmp, name.replace(" ", "_")); os.makedirs(wd)
    out = os.path.join(tmp, name.replace(" ", "_") + ".clstr")
    t0 = time.time()
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", binary), "--id", "0.9", "--threads", threads] + extra + [fasta, "--output", out],
                       cwd=wd, capture_output=True, text=True, env=dict(os.environ, **env))
    dt = time.time() - t0
    stamps = [l for l in r.stdout.splitlines() if "timestamp" in l.lower() or "clock" in l.lower()]
    print("%-24s %s n=%d threads=%s: rc=%d wall %.2f s" % (name, cfg, len(seqs), threads, r.returncode, dt))
    for l in stamps[-8:]:
        print("    " + l.strip())
    if r.returncode != 0:
        print(r.stdout[-1500:], r.stderr[-1500:])
        sys.exit(1)
    res[name] = (parse_clstr(out), open(os.path.join(wd, "weights.txt")).read())
a = res["meshclust2"]
for name in list(res)[1:]:
    b = res[name]
    print("%s: clusters reference %d, b200 %d, identical sets: %s; weights identical: %s" % (name, len(a[0]), len(b[0]), a[0] == b[0], a[1] == b[1]))
    if a[1] != b[1]:
        print("    (the two runs trained different models: with --threads > 1 the reference's own OpenMP training is not "
              "reproducible run to run, SURVEY section 4 -- compare clusters at --threads 1)")
