"""scripts/ncu_lines.py <report.ncu-rep> [top] -- stall samples of one ncu capture aggregated per CUDA source line
(needs -lineinfo and --import-source on).  Run here, no GPU needed."""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
cur_file = "?"
agg = defaultdict(lambda: defaultdict(int))
text = {}
line_key = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 10 and r[0] == "Line No":
        hdr = r
        idx = {}
        for i, h in enumerate(hdr):
            idx.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":                       # a CUDA source line row
        line_key = (cur_file, int(r[0]))
        text[line_key] = r[1].strip()
        continue
    if line_key is None:
        continue
    for k in ("# Samples", "stall_barrier", "stall_no_inst", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_membar",
              "stall_math", "stall_selected", "stall_not_selected", "stall_branch_resolving", "stall_sleep", "Instructions Executed"):
        try:
            agg[line_key][k] += int(r[idx[k]] or 0)
        except (KeyError, ValueError):
            pass
tot = sum(v["# Samples"] for v in agg.values())
print(f"total samples {tot}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    parts = " ".join(f"{k[6:]}={v[k]}" for k in v if k.startswith("stall_") and v[k] > 0.05 * v["# Samples"])
    print(f"{100.0 * v['# Samples'] / max(tot, 1):5.1f}%  {key[0]}:{key[1]:<4d} inst={v['Instructions Executed']:<9d} {parts}\n        {text.get(key, '')[:120]}")
