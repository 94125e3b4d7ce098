#!/usr/bin/env python
"""SASS summary of the hot kernels: `cuobjdump -sass libba_gpu.so` reduced, per kernel, to the instruction count and the
memory / fp64 / synchronisation mnemonics that matter.  Runs without a GPU.
  python scripts/sass_summary.py > profiles/r01_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "3dsmc-bundle-adjustment_b200", "libba_gpu.so")
HOT = re.compile(r"k_linearize|kf_linearize|kt_linearize|k_schur_pass|kf_schur_pass|kt_schur_fused|k_bsr_spmv|k_pcg_sparse_persistent|"
                 r"k_sp_schur|k_schur_pairs|k_obs_W|k_ldlt2_solve|k_chol_|k_backproject|k_pt_blocks|k_cam_blocks|k_point_inverse|"
                 r"k_candidate|k_cost|k_lm_control")
KEEP = re.compile(r"^(LDG|STG|LDS|STS|LD\.|ST\.|DFMA|DMUL|DADD|MUFU|SHFL|BAR|ATOM|RED|LDC|LDCU|ACQBULK|PREEXIT|ERRBAR|MEMBAR|SYNCS|UTMA|UBLK)")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
print("# round 1 -- SASS summary of the hot kernels (cuobjdump -sass libba_gpu.so, sm_100a; scripts/sass_summary.py).")
print("# Per kernel: instruction count and the memory / fp64 mnemonics that matter: LDG.E.128 / LDG.E.ENL2.256 (vector loads),")
print("# STG.E.128, LDS/STS (shared memory), DFMA/DMUL/DADD (fp64 pipe), SHFL (warp reductions), BAR, ATOM/RED (none on fp64),")
print("# ACQBULK / PREEXIT (griddepcontrol.wait / launch_dependents of the programmatic dependent launches).")
name, ops, n = None, collections.Counter(), 0


def flush():
    if name and HOT.search(name):
        print("\n%s\n  instructions: %d\n  %s" % (name, n, ", ".join("%s x%d" % kv for kv in sorted(ops.items()))))


for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        name, ops, n = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        n += 1
        op = m.group(1)
        if KEEP.match(op):
            ops[op] += 1
flush()
