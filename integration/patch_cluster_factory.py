"""Build-time edit of the reference's src/cluster/ClusterFactory.cpp for the relinked meshclust2 (oracle/Makefile, `make
integrated`): the two per-center loops of ClusterFactory<T>::MS and its merge() call first offer the whole pass to
mc2_batched_update / mc2_batched_merge (integration/update_batch_b200.h) and run unchanged when those decline.

usage: patch_cluster_factory.py <reference ClusterFactory.cpp> <output .cpp>

Nothing of the reference is stored in this repository: the script reads the source where it lies and writes the edited
copy under oracle/_ref/ (git-ignored).  It fails loudly if the reference text is not what the hunks expect."""
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
text = open(src).read()


def sub_once(pattern, repl, text, what):
    out, n = re.subn(pattern, repl, text, flags=re.M)
    if n != 1:
        sys.exit("patch_cluster_factory: expected exactly one %s, found %d" % (what, n))
    return out


# hunk 0: the declarations
text = sub_once(r'^#include "Center.h"\n', '#include "Center.h"\n#include "update_batch_b200.h"\n', text, "Center.h include")

# hunk 1: the update loop + merge of every iteration (ClusterFactory.cpp:639-643)
text = sub_once(
    r'^(?P<i>[ \t]*)#pragma omp parallel for\n'
    r'(?P<loop>[ \t]*for \(int j = 0; j < part\.size\(\); j\+\+\) \{\n'
    r'[ \t]*mean_shift_update\(part, j, trn, delta\);\n'
    r'[ \t]*\}\n)'
    r'[ \t]*merge\(part, trn, delta, bandwidth\);\n',
    lambda m: (m.group("i") + "if (!mc2_batched_update(part, trn, delta)) {\n"
               + m.group("i") + "#pragma omp parallel for\n" + m.group("loop")
               + m.group("i") + "}\n"
               + m.group("i") + "if (!mc2_batched_merge(part, trn, delta)) {\n"
               + m.group("i") + "\tmerge(part, trn, delta, bandwidth);\n"
               + m.group("i") + "}\n"),
    text, "update loop followed by merge()")

# hunk 2: the final pass with delta = 0 (ClusterFactory.cpp:648-651)
text = sub_once(
    r'^(?P<i>[ \t]*)#pragma omp parallel for\n'
    r'(?P<loop>[ \t]*for \(int j = 0; j < part\.size\(\); j\+\+\) \{\n'
    r'[ \t]*mean_shift_update\(part, j, trn, 0\);\n'
    r'[ \t]*\}\n)',
    lambda m: (m.group("i") + "if (!mc2_batched_update(part, trn, 0)) {\n"
               + m.group("i") + "#pragma omp parallel for\n" + m.group("loop")
               + m.group("i") + "}\n"),
    text, "final update loop")

open(dst, "w").write(text)
