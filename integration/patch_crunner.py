"""Build-time edit of the reference's src/cluster/CRunner.cpp for the relinked meshclust2 (oracle/Makefile, `make
integrated`): Runner::get_points first offers the sequences of each FASTA file to mc2_batched_get_points
(integration/get_points_b200.h: one K1 batch on the device) and runs its own per-sequence Loader<T>::get_point loop when that
declines.

usage: patch_crunner.py <reference CRunner.cpp> <output .cpp>

Nothing of the reference is stored in this repository: the script reads the source where it lies and writes the edited
copy under oracle/_ref/ (git-ignored).  It fails loudly if the reference text is not what the hunks expect."""
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
text = open(src).read()


def sub_once(pattern, repl, text, what):
    out, n = re.subn(pattern, repl, text, flags=re.M)
    if n != 1:
        sys.exit("patch_crunner: expected exactly one %s, found %d" % (what, n))
    return out


text = sub_once(r'^#include "CRunner.h"\n', '#include "CRunner.h"\n#include "get_points_b200.h"\n', text, "CRunner.h include")

# the per-sequence loop of get_points (CRunner.cpp:523-528)
text = sub_once(
    r'^(?P<i>[ \t]*)(?P<loop>for \(auto elt : \*chromList\) \{\n'
    r'[ \t]*ChromosomeOneDigitDna\* chrom = dynamic_cast<ChromosomeOneDigitDna\*>\(elt\);\n'
    r'[ \t]*Point<T>\* pt = Loader<T>::get_point\(chrom, _id, k\);\n'
    r'#pragma omp critical\n'
    r'[ \t]*points\.push_back\(pt\);\n'
    r'[ \t]*\}\n)',
    lambda m: (m.group("i") + "if (!mc2_batched_get_points<T>(*chromList, _id, k, points)) {\n"
               + m.group("i") + m.group("loop")
               + m.group("i") + "}\n"),
    text, "per-sequence get_point loop")

# the per-file body (CRunner.cpp:521-522): a multi-record FASTA first goes to the device reader
text = sub_once(
    r'^(?P<i>[ \t]*)(?P<maker>ChromListMaker maker\(files\.at\(i\), is_single_file\);\n)',
    lambda m: (m.group("i") + "if (mc2_batched_read_points<T>(files.at(i), is_single_file, _id, k, points)) {\n"
               + "#pragma omp critical\n"
               + m.group("i") + "\tprog++;\n"
               + m.group("i") + "\tcontinue;\n"
               + m.group("i") + "}\n"
               + m.group("i") + m.group("maker")),
    text, "ChromListMaker construction in get_points")

# the per-file body of Runner::run's width detection (CRunner.cpp:61-63)
text = sub_once(
    r'^(?P<i>[ \t]*)(?P<maker>ChromListMaker maker\(f, is_single_file\);\n)',
    lambda m: (m.group("i") + "{\n"
               + m.group("i") + "\tuint64_t mc2_largest = 0;\n"
               + m.group("i") + "\tif (mc2_batched_largest_count(f, is_single_file, k, mc2_largest)) {\n"
               + "#pragma omp critical\n"
               + m.group("i") + "\t\t{\n"
               + m.group("i") + "\t\t\tif (mc2_largest > largest_count) {\n"
               + m.group("i") + "\t\t\t\tlargest_count = mc2_largest;\n"
               + m.group("i") + "\t\t\t}\n"
               + m.group("i") + "\t\t\tprogress++;\n"
               + m.group("i") + "\t\t}\n"
               + m.group("i") + "\t\tcontinue;\n"
               + m.group("i") + "\t}\n"
               + m.group("i") + "}\n"
               + m.group("i") + m.group("maker")),
    text, "ChromListMaker construction in Runner::run")

# the per-file body of Runner::find_k (CRunner.cpp:484-493)
text = sub_once(
    r'^(?P<i>[ \t]*)(?P<maker>ChromListMaker maker\(all_files\.at\(i\), is_single_file\);\n)',
    lambda m: (m.group("i") + "{\n"
               + m.group("i") + "\tunsigned long long mc2_sum = 0, mc2_n = 0;\n"
               + m.group("i") + "\tif (mc2_batched_effective_length(all_files.at(i), is_single_file, mc2_sum, mc2_n)) {\n"
               + m.group("i") + "\t\tunsigned long long mc2_l = mc2_sum / mc2_n;\n"
               + "#pragma omp atomic\n"
               + m.group("i") + "\t\tlength += mc2_l;\n"
               + m.group("i") + "\t\tcontinue;\n"
               + m.group("i") + "\t}\n"
               + m.group("i") + "}\n"
               + m.group("i") + m.group("maker")),
    text, "ChromListMaker construction in Runner::find_k")

open(dst, "w").write(text)
