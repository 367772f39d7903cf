#!/usr/bin/env python
"""One small launch of every hand-written pipeline kernel, for `compute-sanitizer --tool racecheck|synccheck|memcheck`
(tools/gpu_sanitize.sh): attention variants (attn2h d=40 self, attn2x d=80 self, xattn d=40 / d=80 cross, attn d=160,
vattn d=512), the one-kernel GroupNorm, LayerNorm, and the fused CFG + scheduler step.  Sizes are tiny: the sanitizer
slows kernels by two orders of magnitude."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minsdtf_b200 import _lib  # noqa: E402
from minsdtf_b200._lib import StepCoef  # noqa: E402
from minsdtf_b200.engine import Engine  # noqa: E402


def main():
    e = Engine(0)
    rng = np.random.default_rng(0)
    dl = _lib.DL()
    # (run with SDTF_ATTN_PERSIST=2: the (1, 768, 256, 2, 40) case is then 6 work items on 2 persistent CTAs)
    for (B, nq, nk, heads, d) in [(1, 256, 256, 2, 40), (1, 768, 256, 2, 40), (1, 256, 256, 2, 80), (1, 768, 256, 2, 80), (1, 256, 77, 2, 40), (1, 256, 77, 2, 80), (1, 128, 128, 2, 160),
                                  (1, 256, 256, 1, 512)]:
        q = rng.standard_normal((B, nq, heads * d)).astype(np.float32)
        k = rng.standard_normal((B, nk, heads * d)).astype(np.float32)
        v = rng.standard_normal((B, nk, heads * d)).astype(np.float32)
        out = np.empty_like(q)
        e._check(e._lib.sdtf_test_attention(e._h, dl(q), dl(k), dl(v), heads, dl(out)))
        assert np.isfinite(out).all()
        print("attention", (B, nq, nk, heads, d), "ok", flush=True)
    for (shape, mode) in [((2, 16, 16, 320), 1), ((1, 8, 8, 1280), 0), ((1, 16, 16, 320), 2)]:
        x = rng.standard_normal(shape).astype(np.float32)
        g, b = np.ones(shape[-1], np.float32), np.zeros(shape[-1], np.float32)
        out = np.empty_like(x)
        e._check(e._lib.sdtf_test_norm(e._h, dl(x), dl(g), dl(b), mode, dl(out)))
        assert np.isfinite(out).all()
        print("norm", shape, mode, "ok", flush=True)
    eu, ec, x = (rng.standard_normal((2, 16, 16, 4)).astype(np.float32) for _ in range(3))
    got = e.cfg_sched_step(eu, ec, x, StepCoef(7.5, 0.7, 0.9, 0.1, 0.0, 0, 0))
    assert np.isfinite(got).all()
    print("cfg_sched ok", flush=True)
    if len(sys.argv) > 1 and sys.argv[1] == "pipeline":
        # a whole (tiny) job under the sanitizer: UNet (all four levels, folded upsamplers, folded LayerNorms, split-K),
        # ControlNet + HintNet with in-epilogue injection, CFG step, VAE decode with the d = 512 flash attention
        from minsdtf_b200 import synth
        from minsdtf_b200.scheduler import Scheduler, timestep_embedding
        e.load_state_dict(synth.make_state_dict("unet"), "unet")
        e.load_state_dict(synth.make_controlnet_state_dict(), "controlnet")
        e.load_state_dict(synth.make_state_dict("decoder"), "vae_decoder")
        B, h = 1, 16
        sch = Scheduler(active_tcd=False)
        sch.set_timesteps(25)
        ts = [int(t) for t in sch.timesteps[:2]]
        img = e.denoise(synth.latents(B, h, h), synth.context(B), synth.uncond_context(B), np.stack([timestep_embedding(t) for t in ts]),
                        sch.coefficients(ts, 7.5, 0.7), hint_image=(synth.edge_map(8 * h, 8 * h).astype(np.float32) / 255.0)[None],
                        decode=True, use_cuda_graph=False)
        assert img.shape == (B, 8 * h, 8 * h, 3)
        print("pipeline (unet + controlnet + decode, 2 steps) ok", flush=True)
    e.close()


if __name__ == "__main__":
    main()
