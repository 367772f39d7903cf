#!/usr/bin/env python
"""Generates tests/golden/reference_run.npz by running the REFERENCE'S OWN CODE (/root/reference/stable_diffusion,
unmodified) on the torch-backed Keras stand-in (oracle/keras_shim): its model builders, its positional checkpoint loader,
its prompt weighting and its `generate_image` loop, on the seeded synthetic checkpoints / inputs of tests/golden_cases.py.
The GPU box has no /root/reference: tests/test_gpu_reference_golden.py compares the engine with this file there.

    python tools/make_golden_ref.py            # ~2 minutes of CPU
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_cases as G  # noqa: E402
import ref_harness as RH  # noqa: E402


def build(out_path, workdir):
    from minsdtf_b200 import synth
    from minsdtf_b200.scheduler import timestep_embedding
    paths, _ = RH.write_checkpoints(workdir)
    out = {}
    sd = RH.reference_pipeline(paths)
    sdc = RH.reference_pipeline(paths, control=True)
    q = RH.quiet

    # ---- single model calls --------------------------------------------------------------------------------------
    lat, ctx = synth.latents(2, G.h, G.h, seed=1), synth.context(2, 77, seed=2)
    te = np.repeat(timestep_embedding(500)[None], 2, 0)
    out["unet_eps"] = sd.diffusion_model.predict_on_batch([lat, te, ctx])
    img = (G.edges().astype(np.float32) / 255.0)[None]
    hint = q(lambda: sdc.hint_net).predict_on_batch(img)
    out["hint"] = hint[..., ::8]
    res = q(lambda: sdc.control_net).predict_on_batch([lat[:1], te[:1], ctx[:1], hint])
    for i, r in enumerate(res):
        out[f"control_{i}"] = r[..., ::8]
    out["unet_eps_control"] = sdc.diffusion_model.predict_on_batch([lat[:1], te[:1], ctx[:1]] + list(res))
    l2 = synth.latents(1, G.h, G.h, seed=3) * 0.18215 * 3.0
    out["decoded"] = sd.image_decoder.predict_on_batch(l2)
    src = G.source_image().astype(np.float32)[None] / 127.5 - 1.0
    out["encoded"] = q(lambda: sd.image_encoder).predict_on_batch(src)

    # ---- text: tokenizer + attention syntax + windows + textual inversion (long_prompt_weighting.py) ---------------
    out["ctx_prompt"] = sd.encode_text(G.PROMPT)[..., ::4]
    out["ctx_negative"] = sd.encode_text(G.NEGATIVE)[..., ::4]
    out["ctx_long"] = sd.encode_text(G.LONG_PROMPT)[..., ::4]
    import torch
    ti_path = os.path.join(workdir, "ti.pt")
    torch.save({"string_to_param": {"*": torch.from_numpy(G.ti_embedding())}}, ti_path)
    out["ctx_ti"] = sd.encode_text(G.PROMPT, ti_path)[..., ::4]
    out["ctx_empty"] = sd._get_unconditional_context()[..., ::4]

    # ---- the loop (stable_diffusion.py:317-486) ----------------------------------------------------------------------
    def run(pipe, name, fn):
        img = q(fn)
        out[name + "_image"] = img
        out[name + "_latent"] = np.asarray(pipe._image_decoder.last_input, np.float32)

    noise = G.noise()
    ctx1, unc1 = G.contexts()
    run(sd, "txt2img", lambda: sd.generate_image(ctx1, batch_size=1, num_steps=4, diffusion_noise=noise, guidance_rescale=0.7))
    run(sd, "txt2img_norescale", lambda: sd.generate_image(ctx1, batch_size=1, num_steps=4, diffusion_noise=noise))
    run(sd, "img2img", lambda: sd.generate_image(ctx1, batch_size=1, num_steps=10, diffusion_noise=noise, guidance_rescale=0.7,
                                                 reference_image=G.source_image(), reference_image_strength=0.8))
    run(sd, "inpaint", lambda: sd.generate_image(ctx1, batch_size=1, num_steps=10, diffusion_noise=noise, guidance_rescale=0.7,
                                                 reference_image=G.source_image(), reference_image_strength=0.8,
                                                 inpaint_mask=G.mask(), mask_blur_strength=5))
    run(sdc, "controlnet", lambda: sdc.generate_image(ctx1, batch_size=1, num_steps=3, diffusion_noise=noise,
                                                      control_net_image=G.edges()))
    sdt = RH.reference_pipeline(paths, active_tcd=True)
    np.random.seed(123456)
    run(sdt, "tcd", lambda: sdt.generate_image(ctx1, batch_size=1, num_steps=4, diffusion_noise=noise,
                                               unconditional_guidance_scale=0.0))
    # the public entry points with string prompts; `seed=` would draw TF Philox noise, so the noise source is replaced
    sd._get_initial_diffusion_noise = lambda batch_size, seed: noise
    run(sd, "text_to_image", lambda: sd.text_to_image(G.PROMPT, negative_prompt=G.NEGATIVE, batch_size=1, num_steps=4, seed=7))
    run(sd, "text_to_image_long", lambda: sd.text_to_image(G.LONG_PROMPT, batch_size=1, num_steps=3, seed=7))
    calls = []
    run(sd, "inpaint_entry", lambda: sd.inpaint(G.PROMPT, negative_prompt=G.NEGATIVE, batch_size=1, num_steps=10, seed=7,
                                                reference_image=G.source_image(), inpaint_mask=G.mask(), mask_blur_strength=5,
                                                callback=calls.append))
    out["inpaint_entry_callbacks"] = np.asarray(calls, np.int32)
    # LoRA (ckpt_loader.py:2196-2276, 2169-2180): UNet eps and text context of the merged model
    lora_path = os.path.join(workdir, "lora.safetensors")
    synth.save_safetensors(synth.make_lora_state_dict(), lora_path)
    sdl = RH.reference_pipeline(paths, lora_path=lora_path)
    out["lora_unet_eps"] = sdl.diffusion_model.predict_on_batch([lat[:1], te[:1], ctx[:1]])
    out["lora_ctx_prompt"] = sdl.encode_text(G.PROMPT)[..., ::4]
    np.savez_compressed(out_path, **{k: np.asarray(v) for k, v in out.items()})
    return out


if __name__ == "__main__":
    dst = os.path.join(ROOT, "tests", "golden", "reference_run.npz")
    with tempfile.TemporaryDirectory() as d:
        o = build(dst, d)
    print(f"wrote {dst}: {len(o)} arrays, {os.path.getsize(dst) / 1e6:.2f} MB")
