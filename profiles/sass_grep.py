#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of minotaur_b200/libmntr_gpu.so (cuobjdump -sass): the instructions that show which
hardware paths the kernels use.  usage: python profiles/sass_grep.py > profiles/sass_grep.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "minotaur_b200", "libmntr_gpu.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = collections.OrderedDict([
    ("UBLKCP (cp.async.bulk, TMA bulk copy global->shared)", r"\bUBLKCP"),
    ("UTMALDG (tensor-map TMA load)", r"\bUTMALDG"),
    ("SYNCS.* (mbarrier arrive / expect_tx / try_wait)", r"\bSYNCS"),
    ("LDGSTS (cp.async global->shared)", r"\bLDGSTS"),
    ("LDG.E.128 (128-bit global loads)", r"\bLDG\.E\.128|\bLDG\.E\.[A-Z.]*128|\bLD\.E\.128"),
    ("LDG (all global loads)", r"\bLDG\b|\bLDG\."),
    ("STG (global stores)", r"\bSTG"),
    ("ATOMG / RED (global atomics)", r"\bATOMG|\bRED\.|\bATOM\."),
    ("DMUL/DADD/DFMA .RM|.RP (directed fp64)", r"\bD(MUL|ADD|FMA)\.(RM|RP)"),
    ("DMUL/DADD (all fp64 mul/add)", r"\bD(MUL|ADD)\b|\bD(MUL|ADD)\."),
    ("DFMA (fp64 fma: division / sqrt sequences only, --fmad=false)", r"\bDFMA"),
    ("MUFU.RCP64H / MUFU.RSQ64H (fp64 division / sqrt seeds)", r"\bMUFU\.(RCP64H|RSQ64H)"),
    ("UCGABAR (cluster barrier)", r"\bUCGABAR"),
    ("BAR.SYNC (CTA barrier)", r"\bBAR\.SYNC"),
    ("CCTL / prefetch", r"\bCCTL"),
    ("HMMA/IMMA/UTCMMA (tensor cores)", r"\b(HMMA|IMMA|DMMA|UTC[A-Z]*MMA|QGMMA)"),
])
funcs = re.split(r"\n\s*Function : ", sass)
print("SASS mnemonic counts per kernel, %s (sm_100a)\n" % os.path.relpath(so, ROOT))
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"mntr::\(anonymous namespace\)::", "", dem)
    dem = re.sub(r"\(mntr::LinDev.*", "(...)", dem)
    lines = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4,6}\*/\s+[A-Z@]", l)]
    print("%s   [%d SASS instructions]" % (dem[:150], len(lines)))
    for label, pat in pats.items():
        n = sum(1 for l in lines if re.search(pat, l))
        if n: print("    %-70s %6d" % (label, n))
    print()
