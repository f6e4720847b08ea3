"""scripts/ncu_stalls.py <report.ncu-rep> <out.txt> [launch index] -- warp-stall sampling of one captured launch, per
SASS instruction (ncu --set full --import-source on; run here, no GPU needed): totals per stall reason, the
instructions that collect the most samples with their top reasons, and the unit utilisations that say what the
kernel is bound by when its SMs are active."""
import csv
import io
import subprocess
import sys

UTIL = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__issue_active.avg.per_cycle_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum"]


def ncu(rep, page, *extra):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout


def main(rep, out, which=0):
    raw = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units, launches = raw[0], raw[1], raw[2:]
    secs, cur = [], None
    for r in csv.reader(io.StringIO(ncu(rep, "source", "--print-source", "sass"))):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            secs.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and "hdr" in cur and len(r) == len(cur["hdr"]):
            cur["rows"].append(dict(zip(cur["hdr"], r)))
    s, d = secs[which], dict(zip(hdr, launches[which]))
    stall_cols = [h for h in s["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(x["# Samples"]) for x in s["rows"])
    base = int(s["rows"][0]["Address"], 16)
    with open(out, "w") as f:
        f.write(f"# warp-stall sampling of launch {which} of {rep} (scripts/ncu_stalls.py)\n# {s['name']}\n\n")
        for k in UTIL:
            if k in d:
                f.write(f"{k:70s} {d[k]:>16s} {units[hdr.index(k)]}\n")
        if "sm__cycles_active.avg" in d and "sm__cycles_elapsed.avg" in d:
            a, e = float(d["sm__cycles_active.avg"].replace(",", "")), float(d["sm__cycles_elapsed.avg"].replace(",", ""))
            f.write(f"{'SMs active (cycles_active / cycles_elapsed)':70s} {a / e:16.3f}\n")
        f.write(f"\nsamples: {total}\n")
        agg = sorted(((sum(int(x[h]) for x in s["rows"]), h) for h in stall_cols), reverse=True)
        for v, h in agg:
            if v:
                f.write(f"    {h:28s} {v:7d}  {100.0 * v / total:5.1f} %\n")
        f.write("\ninstructions with the most samples (offset, SASS, samples, warp-level executions, top reasons):\n")
        for x in sorted(s["rows"], key=lambda x: -int(x["# Samples"]))[:40]:
            st = sorted(((int(x[h]), h[6:]) for h in stall_cols if int(x[h]) > 0), reverse=True)[:3]
            f.write(f"    {int(x['Address'], 16) - base:#07x}  {x['Source'].strip()[:58]:58s} {int(x['# Samples']):6d} "
                    f"{x['Instructions Executed']:>9s}  {', '.join(f'{n} {v}' for v, n in st)}\n")
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
