"""Text summary of an `ncu --set full --import-source on` report (read on a machine without a GPU): key raw metrics, the stall
reason mix and the hottest SASS lines. Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_xxx.txt ["title"]"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
        "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
lines = [f"# {title}", f"# source: {rep} (ncu --set full --clock-control none --import-source on; cold-cache, serialised replay: shares, not bench values)", ""]
if len(rows) >= 3:
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines.append(f"## kernel: {name[:150]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"{k:75s} {r[i]} {units[i]}")
        lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 3:
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
    T = sum(tot.values()) or 1
    lines.append("## warp stall samples (all warps, whole kernel)")
    lines.append(", ".join(f"{k} {100 * v / T:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 100 > T))
    lines.append("")
    lines.append("## 25 hottest SASS instructions (samples, executions, dominant stall, instruction)")
    top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:25]
    for i in sorted(top):
        r = data[i]
        st = {s: int(r[ix[s]]) for s in stalls}
        lines.append(f"{r[ix['# Samples']]:>6s} {r[ix['Instructions Executed']]:>10s} {max(st, key=st.get):24s} {r[ix['Source']].strip()[:100]}")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out, len(lines), "lines")
