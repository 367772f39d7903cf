#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list of one
bench.py --profile-only run: per-kernel and per-(kernel, grid) device time (and DRAM traffic when captured), split into
denoise step / VAE decode by the marker kernels (temb_select_kernel starts a step, cfg_sched_kernel + step_advance_kernel
end it).
usage: summarize_launches.py launches.csv [traffic.json] > profiles/rNN_launches_summary.md
       (traffic.json: per-kernel launches / DRAM bytes of the denoise step, merged into profiles/ncu_traffic.json by hand)"""
import collections
import csv
import io
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    by_id = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(lines))):
        m = r.get("Metric Name")
        if m not in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
            continue
        e = by_id.setdefault(r["ID"], [r["Kernel Name"].split("(")[0].replace("void ", ""), r["Grid Size"], r["Block Size"], 0.0, None])
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if m == "gpu__time_duration.sum":
            e[3] = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v if u in ("us", "usecond") else v * 1e6 if u == "s" else v  # -> us
        else:
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
            e[4] = (e[4] or 0.0) + v * mult
    return [tuple(e) for e in by_id.values()]


def table(rows, by_grid, top):
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    tot = sum(r[3] for r in rows) or 1.0
    have_dram = any(r[4] is not None for r in rows)
    for n, g, b, v, d in rows:
        k = (n, g) if by_grid else (n,)
        agg[k][0] += 1
        agg[k][1] += v
        agg[k][2] += d or 0.0
    out = ["| ms | share | launches | avg us |" + (" DRAM MB / launch |" if have_dram else "") + " kernel |",
           "|---:|---:|---:|---:|" + ("---:|" if have_dram else "") + "---|"]
    for k, (n, t, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f"| {t / 1e3:.3f} | {100 * t / tot:.1f}% | {n} | {t / n:.1f} |" + (f" {d / n / 1e6:.1f} |" if have_dram else "") +
                   f" `{' grid='.join(k)}` |")
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
        if len(sys.argv) > 2:
            import json
            agg = collections.defaultdict(lambda: {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
            for n, g, b, v, d in step:
                a = agg[n]
                a["launches"] += 1; a["us"] += v; a["dram_bytes"] += d or 0.0
            json.dump(agg, open(sys.argv[2], "w"), indent=1)
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
