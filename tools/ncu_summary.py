"""Summarise an ncu report (--set full) into a markdown table: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md

    python tools/ncu_summary.py rep.ncu-rep --json profiles/x_ncu.json --N 120 --cols 16 --passes 3
also writes the DRAM bytes of ONE block apply (sum over the first `passes` captured kernels of dram__bytes_read.sum +
dram__bytes_write.sum) with the commit it was taken at: bench.py reads `roofline.traffic` from the newest such file."""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
opts = {sys.argv[i][2:]: sys.argv[i + 1] for i in range(2, len(sys.argv) - 1, 2) if sys.argv[i].startswith("--")}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("not_issued")]


def f(r, k, scale=1.0, nd=2):
    try:
        return f"{float(r[idx[k]]) * scale:.{nd}f}"
    except Exception:
        return "n/a"


print(f"# ncu summary of `{rep.split('/')[-1]}` (ncu --set full --clock-control none, B200)\n")
print("| kernel | time ms | DRAM read GB | DRAM write GB | DRAM GB/s | DRAM % of ncu peak | FP64 pipe % | issue active % | smem wavefronts % | warps active % | regs | top stalls |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:60]
    t = float(r[idx["gpu__time_duration.sum"]])
    tu = units[idx["gpu__time_duration.sum"]]
    t_ms = t * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(tu, 1.0)

    def gb(k):
        v = float(r[idx[k]])
        u = units[idx[k]]
        return v * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "Tbyte": 1e3}.get(u, 1.0)
    rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
    vals = [(float(r[idx[h]] or 0), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in stall]
    tot = sum(v for v, _ in vals) or 1.0
    top = ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(vals, reverse=True)[:4])
    print(f"| `{name}` | {t_ms:.3f} | {rd:.3f} | {wr:.3f} | {(rd + wr) / t_ms * 1e3:.0f} | "
          f"{f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', nd=1)} | "
          f"{f(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', nd=1)} | "
          f"{f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active', nd=1)} | "
          f"{f(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', nd=1)} | "
          f"{f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active', nd=1)} | "
          f"{f(r, 'launch__registers_per_thread', nd=0)} | {top} |")

if "json" in opts:
    npass = int(opts.get("passes", 3))
    tot, names = 0.0, []
    for r in rows[2:2 + npass]:
        def gbv(k):
            v = float(r[idx[k]])
            u = units[idx[k]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
        tot += gbv("dram__bytes_read.sum") + gbv("dram__bytes_write.sum")
        names.append(r[idx["Kernel Name"]].split("(")[0].replace("void ", ""))
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    with open(opts["json"], "w") as f:
        json.dump({"N": int(opts.get("N", 120)), "cols": int(opts.get("cols", 16)), "passes": npass, "dram_bytes_per_apply": tot,
                   "kernels": names, "report": rep.split("/")[-1], "commit": commit}, f, indent=1)
