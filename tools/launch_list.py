"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list into a markdown table.
usage: python tools/launch_list.py launches.csv "title / command line" > profiles/rN_launches.md"""
import collections, csv, re, sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = re.sub(r"\(.*$", "", r[ix["Kernel Name"]]).strip()
    tot[name][0] += 1
    tot[name][1] += ms
allms = sum(v[1] for v in tot.values())
print("# " + (sys.argv[2] if len(sys.argv) > 2 else "ncu launch list"))
print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.2f%% |" % (name, n, ms, 100 * ms / allms))
