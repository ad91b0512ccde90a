"""Summarise an ncu report (--set full) into a markdown table: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
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
