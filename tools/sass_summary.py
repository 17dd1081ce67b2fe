"""profiles/r0N_sass_summary.txt: per-kernel counts of the Blackwell-native SASS instructions in the built library (run here, no GPU):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "3dmedicalimagesegmentation_b200", "csrc", "libunetr_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
OPS = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA")
cur, cnt, nfun = None, collections.defaultdict(collections.Counter), 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        nfun += 1
        continue
    if cur:
        for op in OPS:
            if re.search(r"\b" + op + r"(\.|\b)", line):
                cnt[cur][op] += 1
names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
dem = dict(zip(cnt, names))
rows = [(k, v) for k, v in cnt.items() if any(v[o] for o in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP"))]
sha = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
print(f"# SASS evidence: cuobjdump -sass csrc/libunetr_b200.so (sm_100a), tree {sha}.  PTX names never appear in SASS:")
print("# tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, cp.async.bulk.tensor load / store = UTMALDG / UTMASTG, cp.async.bulk (1-D) = UBLKCP.")
print("# Counts are static instruction counts per kernel (template instantiations listed separately).\n")
print("| kernel | UTCHMMA | LDTM | UTMALDG | UTMASTG | UBLKCP |\n|---|---|---|---|---|---|")
tot = collections.Counter()
for k, v in sorted(rows, key=lambda kv: (-kv[1]["UTCHMMA"], dem[kv[0]])):
    name = re.sub(r"\(.*", "", dem[k]).replace("b200::", "")[:120]
    print(f"| `{name}` | {v['UTCHMMA']} | {v['LDTM']} | {v['UTMALDG']} | {v['UTMASTG']} | {v['UBLKCP']} |")
    tot.update(v)
print(f"| **total, {len(rows)} kernels** | {tot['UTCHMMA']} | {tot['LDTM']} | {tot['UTMALDG']} | {tot['UTMASTG']} | {tot['UBLKCP']} |")
print(f"\nSTTM (tcgen05.st): {sum(v['STTM'] for v in cnt.values())}.  Legacy tensor path (HMMA from mma.sync / wmma): {sum(v['HMMA'] for v in cnt.values())} instructions in {nfun} kernels.")
