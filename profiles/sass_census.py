#!/usr/bin/env python
"""SASS opcode census of the measured kernels: static counts from `cuobjdump -sass` of the sm_100a objects in
maximilian_b200/build/ (run after `python -m maximilian_b200.build`).

    python profiles/sass_census.py > profiles/r02_sass_mnemonics.txt
"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "maximilian_b200", "build")
KERNELS = [
    ("K1 bank_kernel<saw, svf_lp, no env, out, no mix> (headline)", "bank_k_svf_lp.o", "bank_kernelILi2ELi4ELi0ELb1ELb0ELi0E"),
    ("K1 bank_kernel<saw, svf_lp, no env, out, no mix, MOD = 1> (configs[1] with a per-sample frequency stream)", "bank_k_svf_lp.o", "bank_kernelILi2ELi4ELi0ELb1ELb0ELi1E"),
    ("K2 delay_bank_kernel<saw, none, env, f64 out, no mix> (configs[2])", "delay_k_none.o", "delay_bank_kernelILi2ELi0ELi1ELi1ELb0ELb0E"),
    ("K2 delay_bank_kernel<saw, none, env, f64 out, no mix, modulated> (configs[2] with a per-sample frequency stream)", "delay_km_none.o", "delay_bank_kernelILi2ELi0ELi1ELi1ELb0ELb1E"),
    ("K4s stft_stream_kernel<MFCC only> (configs[3])", "spectral.o", "stft_stream_kernelILb0E"),
    ("K4s stft_stream_kernel<all outputs>", "spectral.o", "stft_stream_kernelILb1E"),
    ("K8 patch_kernel (interpreter)", "patch.o", "patch_kernel"),
]
WATCH = ["UBLKCP", "SYNCS", "FENCE", "CCTL", "FMUL2", "FADD2", "FFMA2", "DMMA", "HMMA", "LDGSTS", "DADD", "DMUL", "DFMA", "FADD", "FMUL", "FFMA",
         "SHFL", "MUFU", "LDS", "STS", "LDG", "STG", "BAR", "BRA"]

print("# SASS opcode census of the measured kernels (cuobjdump -sass of the sm_100a objects in maximilian_b200/build/, static counts).")
print("# Blackwell-specific: UBLKCP = cp.async.bulk (the TMA engine's 1-D bulk copy: K2 moves a warp's 4 KB ring window with one instruction")
print("# each way), SYNCS = mbarrier arrive / try_wait (its completion), FMUL2 / FADD2 / FFMA2 = packed fp32 (K4s: two channels per")
print("# register pair; every FFMA2 is a fma(p, +-1, q), i.e. a separately rounded sum of two rounded products). DMMA = fp64 tensor-core")
print("# MMA (mma.sync.m8n8k4.f64, the MFCC DCT). The recurrences are DADD/DMUL, never contracted (-fmad=false); DFMA/FFMA sit inside")
print("# the correctly rounded division / sqrt / libdevice sequences and in the mix accumulation, where fusing is allowed.")
for title, obj, sym in KERNELS:
    path = os.path.join(BUILD, obj)
    if not os.path.exists(path):
        print(f"\n== {title}\n   (object {obj} not built)")
        continue
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    hit = [b for b in blocks if sym in b.split("\n", 1)[0]]
    if not hit:
        print(f"\n== {title}\n   (symbol {sym} not found)")
        continue
    body = hit[0]
    name = body.split("\n", 1)[0].strip()
    ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body, flags=re.M)
    base = Counter(o.split(".")[0] for o in ops)
    forms = Counter(o for o in ops if o.split(".")[0] in ("UBLKCP", "SYNCS", "DMMA", "LDG", "STG", "LDGSTS", "LDS", "STS"))
    print(f"\n== {title}\n   {name}")
    print(f"   {len(ops)} instructions; " + ", ".join(f"{w} {base.get(w, 0)}" for w in WATCH))
    print("   memory / sync / tensor forms: " + ", ".join(f"{k} x{v}" for k, v in sorted(forms.items())))
