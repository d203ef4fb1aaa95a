#!/usr/bin/env python3
"""Writes tests/golden/op_counts.json: the number of 254-bit Montgomery products each device
primitive executes, counted by running the SAME headers on the host (tests/hostsim,
-DRB_HOST_SIM).  bench.py and DESIGN.md build the per-kernel roofline numerators from these."""
import ctypes, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
so = os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
hs = ctypes.CDLL(so)
names = ["fe_inv", "g1_madd", "g1_dbl", "g2_madd", "g2_dbl", "fp2_inv", "miller_single", "final_exponentiation", "fp12_mul", "fp12_sqr",
         "fp12_cyclotomic_sqr", "fp12_mul_by_line", "fp12_inv", "to_mont", "g1_on_curve", "g2_on_curve", "miller_lines_for", "miller_fixed", "miller_pair", "fp12_mul_by_line_pair", "miller_fixed4", "miller_pair3", "miller_pair_unit", "miller_fixed4_unit", "miller_lines_normalize"]
out = (ctypes.c_ulonglong * len(names))()
hs.hs_op_counts(out)
counts = dict(zip(names, [int(x) for x in out]))
json.dump(counts, open(os.path.join(ROOT, "tests", "golden", "op_counts.json"), "w"), indent=1)
print(json.dumps(counts))
