#!/usr/bin/env python
"""In-graph cost of each kernel class of one denoise step (batch 8 prompts -> UNet batch 16, 512x512): re-runs the
resident 25-step loop in subprocesses with SDTF_SKIP=<class> and reports step-time differences.  GPU box only."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(skip):
    env = dict(os.environ)
    if skip:
        env["SDTF_SKIP"] = skip
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "1", "--profile-only"],
                       capture_output=True, text=True, env=env, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        return None
    return json.loads(line[-1])["timings"]["loop_ms"] / 25


def main():
    base = run("")
    out = {"step_ms": base}
    for k in ("conv", "attn", "gn", "ln", "misc"):
        t = run(k)
        out[k + "_ms"] = None if t is None else round(base - t, 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
