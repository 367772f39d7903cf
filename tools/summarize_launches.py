#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of one bench.py --profile-only run:
per-kernel and per-(kernel, grid) device time, split into denoise step / VAE decode by the marker kernels
(temb_select_kernel starts a step, cfg_sched_kernel + step_advance_kernel end it).
usage: summarize_launches.py launches.csv > profiles/rNN_launches_summary.md"""
import collections
import csv
import io
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v  # -> us
        rows.append((r["Kernel Name"].split("(")[0].replace("void ", ""), r["Grid Size"], r["Block Size"], v))
    return rows


def table(rows, by_grid, top):
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = sum(r[3] for r in rows) or 1.0
    for n, g, b, v in rows:
        k = (n, g) if by_grid else (n,)
        agg[k][0] += 1
        agg[k][1] += v
    out = ["| ms | share | launches | avg us | kernel |", "|---:|---:|---:|---:|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f"| {t / 1e3:.3f} | {100 * t / tot:.1f}% | {n} | {t / n:.1f} | `{' grid='.join(k)}` |")
    return "\n".join(out), tot / 1e3


def main():
    rows = load(sys.argv[1])
    work = [r for r in rows if r[0] not in ("pack_weight_kernel", "gather_f32_kernel", "half_to_float_kernel", "bf16_to_float_kernel")]
    starts = [i for i, r in enumerate(work) if r[0] == "temb_select_kernel"]
    ends = [i for i, r in enumerate(work) if r[0] == "step_advance_kernel"]
    print(f"# ncu launch list summary ({sys.argv[1]})\n")
    print("Per-launch times are cold-cache and serialised (ncu replays each kernel alone): compare SHARES, not absolutes.\n")
    print(f"{len(rows)} launches total, {len(rows) - len(work)} of them one-time weight packing; {len(starts)} denoise steps.\n")
    if starts and ends:
        step = work[starts[-1]:ends[-1] + 1]
        t, tot = table(step, False, 20)
        print(f"## One denoise step (last of the run): {len(step)} launches, {tot:.2f} ms\n\n{t}\n")
        t, _ = table(step, True, 25)
        print(f"### by grid\n\n{t}\n")
        dec = work[ends[-1] + 1:]
        if dec:
            t, tot = table(dec, False, 12)
            print(f"## VAE decode + uint8: {len(dec)} launches, {tot:.2f} ms\n\n{t}\n")
            t, _ = table(dec, True, 12)
            print(f"### by grid\n\n{t}\n")
    else:
        t, tot = table(work, True, 40)
        print(f"## all kernels, {tot:.2f} ms\n\n{t}\n")


if __name__ == "__main__":
    main()
