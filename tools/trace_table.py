#!/usr/bin/env python
"""Summarise an SDTF_TRACE=1 log (one '[trace] kind shape us TFLOP/s GB/s' line per operator launch, eager mode):
per (kind, shape) launch count, total and mean time, achieved TFLOP/s / GB/s, and the time that would be recovered at a
target rate.  usage: trace_table.py trace.log [first_line_marker_count]  > profiles/rNN_trace_table.md"""
import collections
import re
import sys

PAT = re.compile(r"\[trace\] (\S+)\s+(.*?)\s+([\d.]+) us\s+([\d.]+) TFLOP/s\s+([\d.]+) GB/s")


def main():
    rows = []
    for l in open(sys.argv[1]):
        m = PAT.search(l)
        if m:
            rows.append((m.group(1), m.group(2).strip(), float(m.group(3)), float(m.group(4)), float(m.group(5))))
    agg = collections.OrderedDict()
    for k, s, us, tf, gb in rows:
        a = agg.setdefault((k, s), [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += us; a[2] += tf * us; a[3] += gb * us
    tot = sum(a[1] for a in agg.values())
    print(f"# SDTF_TRACE operator table ({len(rows)} launches, {tot / 1e3:.2f} ms in operators)\n")
    bykind = collections.defaultdict(float)
    for (k, s), a in agg.items():
        bykind[k] += a[1]
    print("| kind | ms | share |\n|---|---:|---:|")
    for k, v in sorted(bykind.items(), key=lambda kv: -kv[1]):
        print(f"| {k} | {v / 1e3:.3f} | {100 * v / tot:.1f}% |")
    print("\n| kind | shape | launches | total us | mean us | TFLOP/s | GB/s |\n|---|---|---:|---:|---:|---:|---:|")
    for (k, s), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {s} | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {a[2] / a[1]:.0f} | {a[3] / a[1]:.0f} |")


if __name__ == "__main__":
    main()
