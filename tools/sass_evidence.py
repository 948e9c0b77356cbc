"""Writes the SASS mnemonic counts that prove the tcgen05 / TMEM / TMA path of every kernel in libdrpo_sm100.so (cuobjdump -sass)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "distributional-reachability-policy-optimization_b200", "libdrpo_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pats = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR.2CTA.MULTICAST", "UTCBAR.MULTICAST", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS.PHASECHK",
        "SYNCS.ARRIVE", "MUFU.TANH", "F2FP.RELU.BF16"]
tot, per, fn = collections.Counter(), collections.defaultdict(collections.Counter), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m or fn is None:
        continue
    op = m.group(1)
    for p in pats:
        if op.startswith(p):
            tot[p] += 1; per[fn][p.split(".")[0]] += 1
            break
print("# SASS evidence (cuobjdump -sass libdrpo_sm100.so; build.py: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3)\n")
print("## instruction counts over the whole library")
for p in pats:
    print(f"{p:28s} {tot[p]}")
print("\n## per kernel (kernels that issue tcgen05 MMAs)")
for f in sorted(per):
    c = per[f]
    if c["UTCHMMA"]:
        print(f"{f}  UTCHMMA={c['UTCHMMA']} UTCBAR={c['UTCBAR']} LDTM={c['LDTM']} STTM={c['STTM']} UBLKCP={c['UBLKCP']}")
