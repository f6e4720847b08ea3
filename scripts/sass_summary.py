"""scripts/sass_summary.py -- per-kernel counts of the SASS mnemonics that show what the library runs on (B200_PROFILING.md
"What proves a Blackwell-native kernel"): TMA (UTMALDG), mbarrier (SYNCS), packed fp32 FMA (FFMA2), programmatic dependent
launch (ACQBULK / PREEXIT), cp.async (LDGSTS), 128-bit global stores.  Run here: python scripts/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sift-parallel-optimization_b200", "libsspyr.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
KEYS = ["UTMALDG", "SYNCS", "FFMA2", "FFMA", "ACQBULK", "PREEXIT", "LDGSTS", "STG.E.128", "STG.E.EF.128", "LDG.E.128", "ATOMG", "MEMBAR", "BAR.SYNC"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::|sspyr::", "", cur)
        cur = re.sub(r"\(.*\)$", "", cur)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["instructions"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k == "FFMA" and op == "FFMA"):
                if k == "FFMA" and op.startswith("FFMA2"):
                    continue
                per[cur][k] += 1
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): mnemonic counts per kernel (scripts/sass_summary.py)")
print(f"# {len(per)} kernels; rows: kernels with TMA / mbarrier / FFMA2 / PDL, then totals\n")
tot = collections.Counter()
for name, c in per.items():
    tot.update(c)
show = [(n, c) for n, c in per.items() if any(c[k] for k in ("UTMALDG", "FFMA2", "ACQBULK", "SYNCS")) or "extrema" in n or "ref_fused" in n]
fam = collections.OrderedDict()
for n, c in show:
    key = re.sub(r"<.*", "", n)
    fam.setdefault(key, []).append((n, c))
for key, lst in fam.items():
    agg = collections.Counter()
    for _, c in lst:
        agg.update(c)
    print(f"{key}: {len(lst)} instantiations, {agg['instructions']} instructions")
    print("    " + ", ".join(f"{k} {agg[k]}" for k in KEYS if agg[k]))
    n, c = max(lst, key=lambda nc: nc[1]["UTMALDG"] * 1000000 + nc[1]["instructions"])
    print(f"    e.g. {n[:110]}: " + ", ".join(f"{k} {c[k]}" for k in KEYS if c[k]))
print("\nwhole library: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]) + f", {tot['instructions']} instructions")
