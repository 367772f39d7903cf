#!/usr/bin/env python
"""Full-size runs of the five BASELINE.json configurations through the public API (host buffers in, uint8 images out),
one JSON line each: wall ms per generation (after one warm-up), images/s, device loop / decode ms, launches.
Synthetic weights and inputs (minsdtf_b200.synth); GPU box only.  These are coverage + timing runs — parity of each
path is tested at reduced size in tests/test_gpu_parity.py."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minsdtf_b200 import synth  # noqa: E402
from minsdtf_b200.stable_diffusion import StableDiffusion  # noqa: E402


def timed(name, sd, n_img, fn, reps=2):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    dt = (time.perf_counter() - t0) / reps
    tm = sd.engine.timings()
    assert out.dtype == np.uint8 and np.isfinite(out.astype(np.float32)).all()
    print(json.dumps({"config": name, "images": n_img, "wall_ms": round(dt * 1e3, 2), "images_per_s": round(n_img / dt, 3),
                      "loop_ms": round(tm["loop_ms"], 2), "decode_ms": round(tm["decode_ms"], 2), "launches": tm["kernel_launches"],
                      "shape": list(out.shape), "mean": round(float(out.mean()), 2)}), flush=True)


def main():
    sd = StableDiffusion(img_height=512, img_width=512, synthetic=True)
    sd.load_all(control=True, encoder=True)
    eng = sd.engine
    sd.unconditional_context = synth.uncond_context(1)
    ctx1, ctx8 = synth.context(1), synth.context(8)
    # C1: batch 1, 25 DDIM steps, CFG 7.5 (+ rescale 0.7 as the API default)
    timed("C1 txt2img 512 b1 25 steps", sd, 1, lambda: sd.generate_image(ctx1, batch_size=1, num_steps=25, diffusion_noise=synth.latents(1, 64, 64),
                                                                            unconditional_guidance_scale=7.5, guidance_rescale=0.7))
    # C2: batch 8 (the bench configuration)
    timed("C2 txt2img 512 b8 25 steps", sd, 8, lambda: sd.generate_image(ctx8, batch_size=8, num_steps=25, diffusion_noise=synth.latents(8, 64, 64),
                                                                            unconditional_guidance_scale=7.5, guidance_rescale=0.7))
    # C3: img2img / inpaint, strength 0.8, 50 steps -> 40 UNet steps, VAE encoder + masked blend
    src, msk = synth.smooth_image(512, 512), synth.center_mask(512, 512)
    timed("C3 img2img 512 b1 50x0.8", sd, 1, lambda: sd.generate_image(ctx1, batch_size=1, num_steps=50, diffusion_noise=synth.latents(1, 64, 64),
                                                                         reference_image=src, reference_image_strength=0.8, guidance_rescale=0.7))
    timed("C3 inpaint 512 b1 50x0.8", sd, 1, lambda: sd.generate_image(ctx1, batch_size=1, num_steps=50, diffusion_noise=synth.latents(1, 64, 64),
                                                                         reference_image=src, reference_image_strength=0.8, inpaint_mask=msk,
                                                                         mask_blur_strength=5, guidance_rescale=0.7))
    # C4: ControlNet canny, 25 steps
    edges = synth.edge_map(512, 512)
    timed("C4 controlnet 512 b1 25 steps", sd, 1, lambda: sd.generate_image(ctx1, batch_size=1, num_steps=25, diffusion_noise=synth.latents(1, 64, 64),
                                                                              control_net_image=edges, guidance_rescale=0.7))
    timed("C4 controlnet 512 b8 25 steps", sd, 8, lambda: sd.generate_image(ctx8, batch_size=8, num_steps=25, diffusion_noise=synth.latents(8, 64, 64),
                                                                              control_net_image=edges, guidance_rescale=0.7))
    # C5: 768x768 batch 4, 25 steps; TCD 4 steps at g = 0 and g = 7.5
    sd7 = StableDiffusion(img_height=768, img_width=768, synthetic=True, engine=eng)
    sd7.unconditional_context = synth.uncond_context(1)
    ctx4 = synth.context(4)
    timed("C5 txt2img 768 b4 25 steps", sd7, 4, lambda: sd7.generate_image(ctx4, batch_size=4, num_steps=25, diffusion_noise=synth.latents(4, 96, 96),
                                                                             guidance_rescale=0.7))
    tcd = StableDiffusion(img_height=512, img_width=512, synthetic=True, engine=eng, active_tcd=True)
    tcd.unconditional_context = synth.uncond_context(1)
    for g in (0.0, 7.5):
        def run(g=g):
            np.random.seed(123456)
            return tcd.generate_image(ctx8, batch_size=8, num_steps=4, diffusion_noise=synth.latents(8, 64, 64), unconditional_guidance_scale=g)
        timed(f"C5 TCD 512 b8 4 steps g={g}", tcd, 8, run)
    # text tower: tokens -> context
    sd.load_all()
    tok = synth.prompt_tokens(8)
    sd.encode_text(tok)  # loads the (synthetic) text-tower weights
    t0 = time.perf_counter()
    for _ in range(5):
        c = sd.encode_text(tok)
    print(json.dumps({"config": "text tower b8 (tokens -> context)", "wall_ms": round((time.perf_counter() - t0) / 5 * 1e3, 2), "shape": list(c.shape)}))
    eng.close()


if __name__ == "__main__":
    main()
