#!/usr/bin/env python
"""Debug aid: does a UNet call give bit-identical results for a sample evaluated alone and inside a batch?  Runs the
comparison under several kernel-selection switches to localise a batch-dependent code path.  GPU box only."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, numpy as np
sys.path.insert(0, %r)
from minsdtf_b200 import synth
from minsdtf_b200.engine import Engine
from minsdtf_b200.scheduler import timestep_embedding
e = Engine(0)
e.load_state_dict(synth.make_state_dict("unet"), "unet")
h = int(sys.argv[1]); B = 2
lat, ctx = synth.latents(B, h, h, seed=11), synth.context(B, seed=12)
te = np.repeat(timestep_embedding(300)[None], B, axis=0)
mode = sys.argv[2] if len(sys.argv) > 2 else "cmp"
if mode == "both":
    e.unet(lat, te, ctx)
elif mode == "one":
    e.unet(lat[1:], te[1:], ctx[1:])
else:
    both = e.unet(lat, te, ctx); one = e.unet(lat[1:], te[1:], ctx[1:])
    print(json.dumps({"h": h, "max_abs_diff": float(np.abs(both[1:] - one).max()), "equal": bool(np.array_equal(both[1:], one))}))
''' % ROOT

if len(sys.argv) > 1 and sys.argv[1] == "fingerprint":
    # operator-by-operator fingerprints of the last sample (SDTF_TRACE=2): first operator where batch and single differ
    def fps(mode, nb):
        r = subprocess.run([sys.executable, "-c", CHILD, "64", mode], capture_output=True, text=True,
                           env={**os.environ, "SDTF_TRACE": "2", "SDTF_TRACE_BATCH": str(nb)})
        return [l for l in r.stderr.splitlines() if l.startswith("[fp]")]
    a, b = fps("both", 2), fps("one", 1)
    print(len(a), len(b), "fingerprints")
    shown = 0
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y and shown < 6:
            print("op", i, "\n  batch :", x, "\n  single:", y)
            if i: print("  prev  :", a[i - 1])
            shown += 1
    sys.exit(0)
for h in (64, 32):
    for env in ({}, {"SDTF_SPLITK": "0"}, {"SDTF_XATTN": "0"}, {"SDTF_GEMM": "1"}):
        r = subprocess.run([sys.executable, "-c", CHILD, str(h)], capture_output=True, text=True, env={**os.environ, **env})
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        print(env, line[-1] if line else r.stderr[-300:], flush=True)
