"""SASS evidence: per kernel of the shipped library, counts of the instructions that prove what the path runs on
(DMMA = FP64 tensor cores, UBLKCP = cp.async.bulk / TMA bulk copies, SYNCS = mbarrier, RED/ATOMG = global reductions,
BAR = block barriers).  python scripts/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt"""
import collections, re, subprocess, sys, os
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pangulu_b200/lib/libpangulu_sm100_r64.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
keys = ["DMMA", "DFMA", "FFMA", "UBLKCP", "UTMALDG", "SYNCS", "RED", "ATOMG", "ATOMS", "BAR", "LDG", "STG", "LDS", "STS", "SHFL", "WARPSYNC", "NANOSLEEP", "UTCMMA", "LDTM"]
fn, counts, total = None, {}, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(.*", "", fn)
        counts[fn] = collections.Counter(); total[fn] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        total[fn] += 1
        op = m.group(1)
        for k in keys:
            if op.startswith(k):
                counts[fn][k] += 1
                break
print(f"SASS instruction counts per kernel, {os.path.basename(lib)} (cuobjdump -sass, sm_100a)")
print(f"{'kernel':70s} {'total':>7s} " + " ".join(f"{k:>8s}" for k in keys))
for fn in sorted(counts, key=lambda f: -total[f]):
    print(f"{fn[:70]:70s} {total[fn]:7d} " + " ".join(f"{counts[fn][k]:8d}" for k in keys))
print("arch:", subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout.strip().replace("\n", "; "))
