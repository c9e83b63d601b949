#!/usr/bin/env python3
"""Build-time edit of the reference's src/fastcar/FC_Runner.cpp for the relinked fastcar (oracle/Makefile: fastcar_b200):
work() (FC_Runner.cpp:427-470) first offers its block to the device; everything else is the reference's text.
usage: patch_fc_runner.py <reference FC_Runner.cpp> <output .cpp>"""
import sys

src, dst = sys.argv[1], sys.argv[2]
s = open(src).read()
inc = '#include "FC_Runner.h"'
assert s.count(inc) == 1, "include anchor not found"
s = s.replace(inc, inc + '\n#include "fastcar_b200.h"')
anchor = """	if (pts.empty()) {
		return;
	}
	uint8_t mode = pred->get_mode();"""
assert s.count(anchor) == 1, "work() anchor not found"
s = s.replace(anchor, """	if (pts.empty()) {
		return;
	}
	if (mc2_batched_work<T>(queries, pts, similarity, pred, delim, out, num_pred_pos, format, format_header)) {
		return;
	}
	uint8_t mode = pred->get_mode();""")
open(dst, "w").write(s)
