#!/bin/bash
# Static evidence for profiles/: ptxas -v (registers, spills) of the three hot translation units and a SASS excerpt of the tile
# sweep showing the TMA / tcgen05 / EMD mnemonics.  Runs without a GPU.  usage: bash tools/static_profiles.sh r2
set -e
tag=${1:-r2}
cd "$(dirname "$0")/.."
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xptxas -v"
out=profiles/${tag}_ptxas_v.txt
: > $out
for f in tile_sweep pair_score kmer_count; do
  echo "### nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xptxas -v -c meshclust2_b200/csrc/$f.cu" >> $out
  $NV -c -o /tmp/sp_$f.o meshclust2_b200/csrc/$f.cu 2>&1 | python3 -c "
import sys, re
name = None
for line in sys.stdin:
    m = re.search(r\"Compiling entry function '([^']+)'\", line)
    if m: name = m.group(1); props = []; continue
    if name and ('bytes stack frame' in line or 'Used' in line):
        props.append(line.strip().replace('ptxas info    : ', ''))
        if 'Used' in line:
            print(name + '\t' + '\t'.join(props)); name = None
" >> $out
done
sass=profiles/${tag}_sass_tile_sweep.txt
/usr/local/cuda/bin/cuobjdump -sass /tmp/sp_tile_sweep.o > /tmp/sp_tile.sass
python3 - "$sass" <<'PY'
import re, sys, collections
lines = open('/tmp/sp_tile.sass').read().split('\n')
# the benchmarked instantiation: NEED_DOT|NEED_EMD, scoring, one-slab rows
start = next(i for i, l in enumerate(lines) if 'Function :' in l and 'tile_sweep_kernelILi6ELb0ELb1E' in l)
end = next((i for i in range(start + 1, len(lines)) if 'Function :' in lines[i]), len(lines))
body = lines[start:end]
ops = collections.Counter()
for l in body:
    m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', l)
    if m: ops[m.group(1)] += 1
ev = ('UTMALDG', 'UTCIMMA', 'UTCHMMA', 'LDTM', 'STTM', 'UTCATOMSWS', 'UTCBAR', 'USETMAXREG', 'VIMNMX.U16x2', 'IDP.2A', 'ELECT', 'SYNCS.PHASECHK', 'SYNCS.ARRIVE', 'LDL', 'STL')
with open(sys.argv[1], 'w') as f:
    f.write('# SASS of tile_sweep_kernel<NEED_DOT|NEED_EMD, scoring, one-slab rows> (sm_100a), cuobjdump -sass of tile_sweep.cu.o\n')
    f.write('# evidence lines: TMA = UTMALDG, tcgen05.mma = UTCIMMA (kind::i8), tcgen05.ld/st = LDTM/STTM, tcgen05.alloc = UTCATOMSWS,\n')
    f.write('# tcgen05.commit = UTCBAR, setmaxnreg = USETMAXREG, mbarrier = SYNCS, EMD inner pair = VIMNMX.U16x2 + IDP.2A.LO.U16.U8\n')
    f.write('# static instruction counts: ' + ', '.join('%s %d' % kv for kv in ops.most_common(28)) + '\n')
    f.write('# LDL %d, STL %d in the whole kernel (epilogue / role prologues; none between the first and the last VIMNMX)\n\n' % (ops['LDL'], ops['STL']))
    shown = collections.Counter()
    for l in body:
        for e in ev:
            if e in l and shown[e] < (12 if e in ('VIMNMX.U16x2', 'IDP.2A') else 6):
                shown[e] += 1
                f.write(l.rstrip() + '\n')
                break
PY
echo "wrote $out and $sass"
