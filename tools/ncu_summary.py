#!/usr/bin/env python
"""Compact summary of an .ncu-rep (read here, no GPU needed): duration, DRAM/L2 traffic, tensor-pipe activity, and the
source lines with the most warp-stall samples.  Usage: python tools/ncu_summary.py file.ncu-rep [n_lines]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second", "launch__grid_size", "launch__cluster_size",
    "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    h, units, rows = raw(rep)
    idx = {n: i for i, n in enumerate(h)}
    for r in rows:
        print("## kernel:", r[idx.get("Kernel Name", 4)][:90])
        for k in KEYS:
            if k in idx and r[idx[k]] not in ("", "no data"):
                print(f"  {k:95s} {r[idx[k]]:>16s} {units[idx[k]]}")
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next((i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r), None)
    if hdr is None:
        return
    hh = rows[hdr]
    si, src = hh.index("# Samples"), hh.index("Source")
    stall_cols = [i for i, n in enumerate(hh) if n.startswith("stall_") and "Not Issued" not in n]
    body = [r for r in rows[hdr + 1:] if len(r) > si and r[si].isdigit()]
    tot = sum(int(r[si]) for r in body) or 1
    # stall totals
    agg = {hh[i]: sum(int(r[i]) for r in body if r[i].isdigit()) for i in stall_cols}
    print("## stall samples:", ", ".join(f"{k[6:]}={v * 100 // tot}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
    print(f"## top SASS lines by samples (total {tot}); the previous SASS line is usually the one being waited on")
    order = sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:nlines]
    for i in order:
        r = body[i]
        stalls = sorted(((int(r[c]) if r[c].isdigit() else 0, hh[c][6:]) for c in stall_cols), reverse=True)[:2]
        st = ", ".join(f"{n}={v}" for v, n in stalls if v > 0)
        prev = body[i - 1][src].strip()[:60] if i > 0 else ""
        print(f"  {int(r[si]) * 100.0 / tot:5.1f}%  {r[src].strip()[:70]:70s} [{st}]  <- {prev}")


if __name__ == "__main__":
    main()
