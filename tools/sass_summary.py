"""SASS mnemonic counts per kernel of the built objects (cuobjdump -sass): the proof that the scoring kernel is
tcgen05 / TMEM / TMA code (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, UTCBAR =
tcgen05.commit, SYNCS = mbarrier ops) and that the gather kernels issue 128-bit loads.
    python tools/sass_summary.py [object ...]      (default: the scoring, propagation and BPR objects)"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(REPO, "movie-recommender-system-with-gnns_b200", "csrc", "build")
KEYS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "HMMA", "FFMA", "LDG", "LDS", "STS",
        "SHFL", "VOTE", "ATOMG", "RED", "MEMBAR", "CCTL"]
objs = sys.argv[1:] or [os.path.join(BUILD, n) for n in ("score_topk_tc.cu.o", "propagate.cu.o", "bpr_owner.cu.o",
                                                         "epoch_kernel.cu.o")]
for obj in objs:
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    print(f"== {os.path.basename(obj)}")
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        ops, wide = collections.Counter(), 0
        for m in re.finditer(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M):
            ops[m.group(1).split(".")[0]] += 1
            wide += m.group(1).startswith("LDG.E.128")
        line = ", ".join(f"{k} {ops[k]}" for k in KEYS if ops[k])
        print(f"{demangled[:110]}\n    {sum(ops.values())} instructions: {line}" + (f" (LDG.E.128: {wide})" if wide else ""))
