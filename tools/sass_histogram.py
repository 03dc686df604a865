"""Opcode histogram of the tensor-path kernels in libdeeppde_b200.so (cuobjdump -sass): how many tcgen05 MMAs (UTCHMMA),
TMEM loads / stores (LDTM / STTM), bulk async copies (UBLKCP), mbarrier ops (SYNCS) ... each kernel contains.
    python tools/sass_histogram.py > profiles/r02_sass_opcode_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "deeppde_actorcritic_b200", "libdeeppde_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "HMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "UTCBAR", "RED", "FFMA2", "FADD2", "F2FP", "STL", "LDL", "LDS", "STS", "LDG", "STG", "BAR"]
kern, hist = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        hist[kern]["total"] += 1
        for w in WATCH:
            if op.startswith(w):
                hist[kern][w] += 1
                break
try:
    names = subprocess.run(["c++filt"] + list(hist), capture_output=True, text=True, check=True).stdout.splitlines()
except Exception:
    names = list(hist)
print("SASS opcode counts per kernel of libdeeppde_b200.so (sm_100a).  UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk,")
print("SYNCS = mbarrier ops, UTCBAR = tcgen05.commit, LDL/STL = local memory (register spills, relu masks).  HMMA (legacy mma.sync) must be 0.\n")
for (k, h), n in zip(hist.items(), names):
    if "tc_kernel" not in k and "selftest" not in k and "bench" not in k:
        continue
    n = re.sub(r"\(dpb::tc::TcArgs\)", "", n)
    print(f"{n}\n    total {h['total']:6d} | " + "  ".join(f"{w} {h[w]}" for w in WATCH if h[w] or w in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "HMMA")))
