"""Summarise ncu outputs (run here, no GPU needed):  python scripts/summarize_ncu.py <tag>
Reads gpurun_out/launches_<tag>.csv and gpurun_out/{attn,gemm,rows,conv}_<tag>.ncu-rep, writes profiles/<tag>_ncu_summary.md"""
import collections
import csv
import io
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = [f"# ncu summary {tag}\n"]

try:
    lines = [l for l in open(f"gpurun_out/launches_{tag}.csv") if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e6, "us": v / 1e3, "ms": v, "msecond": v, "s": v * 1e3, "second": v * 1e3}[row["Metric Unit"]]
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:72]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    ours = {k: v for k, v in agg.items() if k.startswith("scail::")}
    tot = sum(a[1] for a in ours.values())
    out.append("## launch list (gpu__time_duration.sum, --clock-control none; cold-cache serialised: compare SHARES)\n")
    out.append("command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 python bench.py --steps 1 --warmup 1`"
               " (window covers model init + the first ~23 blocks of a 14B step)\n")
    out.append("| kernel | launches | total ms | share of scail:: kernels |\n|---|---|---|---|")
    for k, a in sorted(ours.items(), key=lambda x: -x[1][1]):
        out.append(f"| {k} | {a[0]} | {a[1]:.2f} | {100 * a[1] / tot:.1f}% |")
    other = sum(a[1] for k, a in agg.items() if not k.startswith("scail::"))
    out.append(f"\nnon-scail (torch init / fills / cat) in window: {other:.2f} ms (weight init dominates; not part of a step)\n")
except FileNotFoundError:
    pass

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg"]
for k in ("attn", "gemm", "rows", "rope", "conv"):
    try:
        txt = subprocess.run(["ncu", "-i", f"gpurun_out/{k}_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            out.append(f"## {k}: `{vals[hdr.index('Kernel Name')][:90]}`  (ncu --set full --clock-control none)\n")
            out.append("| metric | value | unit |\n|---|---|---|")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    out.append(f"| {w} | {vals[i]} | {units[i]} |")
            out.append("")
    except Exception as e:
        pass
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
