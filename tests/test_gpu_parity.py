"""Parity of the CUDA path (through the C ABI) with the CPU oracle of the reference graphs, at the bars of
BASELINE.json: eps max|d|/max|ref| <= 2e-2 per UNet call, scheduler-step latents <= 1e-3, decoded images >= 30 dB PSNR.
Everything here runs on seeded synthetic weights / inputs (minsdtf_b200.synth); nothing reads /root/reference."""
import os

import numpy as np
import pytest
import torch

from minsdtf_b200 import synth
from minsdtf_b200._lib import StepCoef
from minsdtf_b200.scheduler import Scheduler, timestep_embedding
from oracle import sd15_oracle as O
from oracle.scheduler_oracle import OracleScheduler, cfg_combine

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EPS_BAR = 2e-2


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / np.abs(b).max())


def psnr_u8(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def temb(t, B):
    return np.repeat(timestep_embedding(t)[None], B, axis=0)


@pytest.mark.parametrize("B,h,T,t", [(1, 16, 77, 500), (2, 32, 77, 960), (1, 64, 77, 20), (2, 64, 77, 480), (1, 32, 154, 700),
                                     (1, 96, 77, 500)])  # 96 = the 768x768 configuration (9216-token self-attention)
def test_unet_eps_parity(engine_unet, unet_sd, B, h, T, t):
    lat = synth.latents(B, h, h, seed=11 + h)
    ctx = synth.context(B, T, seed=12 + T)
    ref = O.unet_forward(unet_sd, lat, temb(t, B), ctx)
    got = engine_unet.unet(lat, temb(t, B), ctx)
    assert np.isfinite(got).all()
    r = rel(got, ref)
    print(f"unet B={B} h={h} T={T} t={t}: rel err {r:.4g}")
    assert r <= EPS_BAR, r


def test_unet_rectangular_and_torch_device_tensors(engine_unet, unet_sd):
    """non-square latent (48x32) and CUDA tensors borrowed in place through DLPack"""
    B, h, w = 1, 48, 32
    lat = synth.latents(B, h, w, seed=5)
    ctx = synth.context(B, 77, seed=6)
    ref = O.unet_forward(unet_sd, lat, temb(300, B), ctx)
    got = engine_unet.unet(torch.as_tensor(lat).cuda(), torch.as_tensor(temb(300, B)).cuda(), torch.as_tensor(ctx).cuda())
    assert got.is_cuda
    assert rel(got.cpu().numpy(), ref) <= EPS_BAR


def test_controlnet_and_hintnet_parity(engine_unet, engine_cnet, unet_sd, cnet_sd):
    B, h = 1, 32
    img = (synth.edge_map(8 * h, 8 * h).astype(np.float32) / 255.0)[None]
    hint_ref = O.hintnet_forward(cnet_sd, img)
    hint = engine_cnet.hintnet(img)
    assert rel(hint, hint_ref) <= EPS_BAR
    lat = synth.latents(B, h, h, seed=21)
    ctx = synth.context(B, 77, seed=22)
    te = temb(640, B)
    ctr_ref = O.controlnet_forward(cnet_sd, lat, te, ctx, hint_ref)
    ctr = engine_cnet.controlnet(lat, te, ctx, hint_ref)
    for i, (a, b) in enumerate(zip(ctr, ctr_ref)):
        assert a.shape == b.shape
        assert rel(a, b) <= EPS_BAR, (i, rel(a, b))
    # residual injection into the UNet (diffusion_model.py:230-234), teacher-forced with the oracle's residuals
    ref = O.unet_forward(unet_sd, lat, te, ctx, ctr_ref)
    got = engine_unet.unet(lat, te, ctx, ctr_ref)
    assert rel(got, ref) <= EPS_BAR
    plain = O.unet_forward(unet_sd, lat, te, ctx)
    assert rel(plain, ref) > 1e-3  # the synthetic zero-convs are non-zero: the branch is really exercised


def test_vae_decode_psnr(engine_vae, vae_sd):
    lat = synth.latents(1, 32, 32, seed=31) * 0.18215 * 3.0
    ref = O.vae_decode(vae_sd, lat)
    got = engine_vae.vae_decode(lat)
    assert got.shape == (1, 256, 256, 3)
    u_ref, u_got = O.to_uint8(ref), engine_vae.to_uint8(got)
    p = psnr_u8(u_got, u_ref)
    print(f"vae decode: rel {rel(got, ref):.4g} psnr {p:.2f} dB")
    assert p >= 30.0, p
    # uint8 conversion itself is exact arithmetic (truncation, stable_diffusion.py:483-486)
    assert np.array_equal(engine_vae.to_uint8(ref), u_ref)


def test_vae_encode_parity(engine_vae, vae_sd):
    img = synth.smooth_image(128, 128).astype(np.float32)[None] / 127.5 - 1.0
    ref = O.vae_encode(vae_sd, img)
    got = engine_vae.vae_encode(img)
    assert got.shape == (1, 16, 16, 4)
    assert rel(got, ref) <= EPS_BAR, rel(got, ref)


def test_cfg_scheduler_kernel_vs_reference_golden(engine):
    """fused kernel vs trajectories recorded from the REAL reference scheduler (tests/golden/scheduler.npz)"""
    g = np.load(os.path.join(GOLD, "scheduler.npz"))
    for name, n, tcd in (("ddim25", 25, False), ("ddim4", 4, False), ("tcd4", 4, True)):
        s = Scheduler(active_tcd=tcd)
        s.set_timesteps(n)
        x = g[f"{name}_x0"]
        if tcd:
            np.random.seed(123456)
        for i, t in enumerate(s.timesteps):
            ca, cb, cn = s.step_scalars(int(t))
            noise = np.random.randn(*x.shape).astype(np.float32) if cn != 0.0 else None
            x = engine.cfg_sched_step(None, g[f"{name}_eps"][i], x, StepCoef(0, 0, ca, cb, cn, 0, 0), noise=noise)
            assert np.abs(x - g[f"{name}_out"][i]).max() <= 1e-3, (name, i)


def test_cfg_rescale_and_inpaint_blend_vs_oracle(engine):
    rng = np.random.default_rng(7)
    B, h = 3, 16
    eu, ec, x = (rng.standard_normal((B, h, h, 4)).astype(np.float32) for _ in range(3))
    s = Scheduler(active_tcd=False)
    s.set_timesteps(25)
    o = OracleScheduler(False)
    o.set_timesteps(25)
    t = int(s.timesteps[3])
    o._idx = 3
    for guidance, rescale in ((7.5, 0.0), (7.5, 0.7), (0.0, 0.0)):
        o._idx = 3
        eps = cfg_combine(eu, ec, guidance, rescale) if guidance > 0 else ec
        ref = o.step(eps, t, x)
        ca, cb, cn = s.step_scalars(t)
        got = engine.cfg_sched_step(eu if guidance > 0 else None, ec, x, StepCoef(guidance, rescale, ca, cb, cn, 0, 0))
        assert np.abs(got - ref).max() <= 1e-3, (guidance, rescale)
    # inpaint blend (stable_diffusion.py:469-475)
    mask = rng.random((h, h)).astype(np.float32)
    init = rng.standard_normal((h, h, 4)).astype(np.float32)
    noise = rng.standard_normal((B, h, h, 4)).astype(np.float32)
    o._idx = 3
    ref = o.step(cfg_combine(eu, ec, 7.5, 0.7), t, x)
    orig = o.signal_rates[t] * init[None] + o.noise_rates[t] * noise
    ref = orig * (1 - mask[None, ..., None]) + ref * mask[None, ..., None]
    coef = StepCoef(7.5, 0.7, *s.step_scalars(t), float(s.signal_rates[t]), float(s.noise_rates[t]))
    got = engine.cfg_sched_step(eu, ec, x, coef, mask=mask, init_latent=init, init_noise=noise)
    assert np.abs(got - ref).max() <= 1e-3


def _pipeline(engine, **kw):
    from minsdtf_b200.stable_diffusion import StableDiffusion
    return StableDiffusion(engine=engine, synthetic=True, **kw)


def test_txt2img_loop_teacher_forced_and_free_running(engine_unet, engine_vae, unet_sd, vae_sd):
    """config 1 at reduced size: 32x32 latent, 4 DDIM steps, CFG 7.5, rescale 0.7.  Per-call eps parity is teacher-forced
    on the oracle's latents; the device loop (CUDA graph and eager) is compared end to end."""
    B, h, steps = 1, 32, 4
    noise, ctx, unc = synth.latents(B, h, h), synth.context(B), synth.uncond_context(B)
    trace = {}
    img_ref = O.generate_image({"unet": unet_sd, "vae": vae_sd}, ctx, unc, noise, num_steps=steps, guidance_scale=7.5,
                               guidance_rescale=0.7, trace=trace)
    for i, t in enumerate(trace["t"]):
        lat = trace["latent_in"][i]
        eu = engine_unet.unet(lat, temb(t, B), unc)
        ec = engine_unet.unet(lat, temb(t, B), ctx)
        assert rel(eu, trace["eps_u"][i]) <= EPS_BAR and rel(ec, trace["eps_c"][i]) <= EPS_BAR, (i, t)
    sd = _pipeline(engine_unet, img_height=8 * h, img_width=8 * h)
    sd.unconditional_context = unc[:1]
    img_g, lat_g = sd.generate_image(ctx, batch_size=B, num_steps=steps, diffusion_noise=noise, guidance_rescale=0.7,
                                     return_latent=True, use_cuda_graph=True)
    img_e, lat_e = sd.generate_image(ctx, batch_size=B, num_steps=steps, diffusion_noise=noise, guidance_rescale=0.7,
                                     return_latent=True, use_cuda_graph=False)
    assert np.array_equal(lat_g, lat_e) and np.array_equal(img_g, img_e)  # graph replay == eager launches, bitwise
    lat_ref = trace["latent_out"][-1]
    print(f"free-running final latent rel err {rel(lat_g, lat_ref):.4g}, image psnr {psnr_u8(img_g, img_ref):.2f} dB")
    assert rel(lat_g, lat_ref) <= 5e-2
    assert img_g.dtype == np.uint8 and img_g.shape == (B, 8 * h, 8 * h, 3)
    # decode parity from the identical final latent
    dec = engine_vae.to_uint8(engine_vae.vae_decode(np.asarray(lat_ref, np.float32)))
    assert psnr_u8(dec, img_ref) >= 30.0


def test_batched_cfg_equals_separate_calls(engine_unet):
    """the loop batches uncond+cond into one UNet pass; results must not depend on the batching"""
    B, h = 2, 16
    lat, ctx, unc = synth.latents(B, h, h, seed=3), synth.context(B, seed=4), synth.uncond_context(B, seed=5)
    te = temb(100, B)
    both = engine_unet.unet(np.concatenate([lat, lat]), np.concatenate([te, te]), np.concatenate([unc, ctx]))
    a, b = engine_unet.unet(lat, te, unc), engine_unet.unet(lat, te, ctx)
    assert np.array_equal(both[:B], a) and np.array_equal(both[B:], b)


def test_full_size_batching_and_determinism(engine_unet):
    """BASELINE-size latent (64x64): a UNet call is bit-identical when repeated and when its samples are evaluated in
    one batch or one by one (the one-kernel GroupNorm's partition and the split-K reduction order depend on the
    tensor shape only, never on the batch)"""
    B, h = 2, 64
    lat, ctx = synth.latents(B, h, h, seed=11), synth.context(B, seed=12)
    te = temb(300, B)
    both = engine_unet.unet(lat, te, ctx)
    again = engine_unet.unet(lat, te, ctx)
    one = engine_unet.unet(lat[1:], te[1:], ctx[1:])
    assert np.isfinite(both).all()
    assert np.array_equal(both, again)
    assert np.array_equal(both[1:], one)


def test_img2img_inpaint_controlnet_tcd_paths(engine_unet, engine_vae, engine_cnet, unet_sd, vae_sd, cnet_sd):
    """configs 3, 4 and 5b at reduced size against the oracle loop (free-running, so the bar is looser than per-call)"""
    B, h, H = 1, 16, 128
    ctx, unc = synth.context(B), synth.uncond_context(B)
    noise = synth.latents(B, h, h, seed=9)
    weights = {"unet": unet_sd, "vae": vae_sd, "controlnet": cnet_sd}
    sd = _pipeline(engine_unet, img_height=H, img_width=H)
    sd.unconditional_context = unc[:1]
    # --- inpaint (covers img2img): 10 steps, strength 0.8 -> 8 UNet steps starting one step below init_time ---
    src, msk = synth.smooth_image(H, H), synth.center_mask(H, H)
    in_arr, in_t = sd.preprocessed_image(src)
    m_arr, m_lat = sd.preprocessed_mask(msk, 5)
    init_ref = O.vae_encode(vae_sd, in_t)
    ref = O.generate_image(weights, ctx, unc, noise, num_steps=10, guidance_scale=7.5, guidance_rescale=0.7,
                           init_latent=init_ref, strength=0.8, latent_mask=m_lat, input_image_array=in_arr,
                           input_mask_array=m_arr, decode=False)
    got_img, got = sd.generate_image(ctx, batch_size=B, num_steps=10, diffusion_noise=noise, guidance_rescale=0.7,
                                     reference_image=src, reference_image_strength=0.8, inpaint_mask=msk, mask_blur_strength=5,
                                     return_latent=True)
    print(f"inpaint final latent rel err {rel(got, ref):.4g}")
    assert rel(got, ref) <= 5e-2
    assert got_img.shape == (B, H, H, 3)
    # --- ControlNet: 3 steps ---
    edge = synth.edge_map(H, H)
    hint_ref = O.hintnet_forward(cnet_sd, (edge.astype(np.float32) / 255.0)[None])
    ref = O.generate_image(weights, ctx, unc, noise, num_steps=3, guidance_scale=7.5, guidance_rescale=0.0, hint=hint_ref,
                           decode=False)
    _, got = sd.generate_image(ctx, batch_size=B, num_steps=3, diffusion_noise=noise, control_net_image=edge, return_latent=True)
    print(f"controlnet final latent rel err {rel(got, ref):.4g}")
    assert rel(got, ref) <= 5e-2
    # --- TCD 4 steps, guidance 0 (app.py defaults), noise from the global NumPy RNG ---
    sd_t = _pipeline(engine_unet, img_height=H, img_width=H, active_tcd=True)
    np.random.seed(123456)
    ref = O.generate_image(weights, ctx, None, noise, num_steps=4, guidance_scale=0.0, active_tcd=True, decode=False)
    np.random.seed(123456)
    _, got = sd_t.generate_image(ctx, batch_size=B, num_steps=4, diffusion_noise=noise, unconditional_guidance_scale=0.0,
                                 return_latent=True)
    print(f"tcd final latent rel err {rel(got, ref):.4g}")
    assert rel(got, ref) <= 5e-2


# ----------------------------------------------------------------------------------------------------------------
# BASELINE.json configurations at FULL size against the oracle (VERDICT r01 "what's weak" 2): every model at the
# tensor shapes the benchmark runs, so the 64-CTA GroupNorm partitions, the 512x512x128/256 decoder tensors and the
# 4096 / 9216-token attention meet the oracle too.  The oracle side costs tens of seconds of CPU per test.
# ----------------------------------------------------------------------------------------------------------------
def test_vae_decode_full_size_512(engine_vae, vae_sd):
    """config 1/2 decode: 64x64 latent -> 512x512 image, >= 30 dB against the oracle from the identical latent"""
    lat = synth.latents(1, 64, 64, seed=131) * 0.18215 * 3.0
    ref = O.vae_decode(vae_sd, lat)
    got = engine_vae.vae_decode(lat)
    assert got.shape == (1, 512, 512, 3) and np.isfinite(got).all()
    p = psnr_u8(engine_vae.to_uint8(got), O.to_uint8(ref))
    print(f"vae decode 512x512: rel {rel(got, ref):.4g} psnr {p:.2f} dB")
    assert p >= 30.0, p


def test_vae_encode_full_size_512(engine_vae, vae_sd):
    """config 3 encode: 512x512 image -> 64x64 latent"""
    img = synth.smooth_image(512, 512).astype(np.float32)[None] / 127.5 - 1.0
    ref = O.vae_encode(vae_sd, img)
    got = engine_vae.vae_encode(img)
    assert got.shape == (1, 64, 64, 4)
    # BASELINE.json states no bar for the encoder.  At this size the 24-convolution chain with bf16 activations sits above the
    # 2e-2 eps bar by construction: the fp32 oracle with every stored activation rounded to bf16 (O.set_round_dtype) is
    # itself 3.6e-2 away from the plain fp32 oracle on this input.  Bar: no worse than that storage-precision emulation
    # (and never above 5e-2); rms error <= 2e-2.
    O.set_round_dtype(torch.bfloat16)
    try:
        emu = O.vae_encode(vae_sd, img)
    finally:
        O.set_round_dtype(None)
    r, r_emu = rel(got, ref), rel(emu, ref)
    rms = float(np.sqrt(np.mean((got - ref) ** 2)) / np.sqrt(np.mean(ref ** 2)))
    print(f"vae encode 512x512: rel {r:.4g} (bf16-storage emulation of the oracle: {r_emu:.4g}), rms rel {rms:.4g}")
    assert r <= min(5e-2, max(EPS_BAR, 1.2 * r_emu)), (r, r_emu)
    assert rms <= 2e-2, rms


def test_controlnet_and_hintnet_full_size_64(engine_unet, engine_cnet, unet_sd, cnet_sd):
    """config 4: HintNet on a 512x512 edge map, ControlNet at the 64x64 latent, residual injection into the UNet"""
    B, h = 1, 64
    img = (synth.edge_map(8 * h, 8 * h).astype(np.float32) / 255.0)[None]
    hint_ref = O.hintnet_forward(cnet_sd, img)
    hint = engine_cnet.hintnet(img)
    assert rel(hint, hint_ref) <= EPS_BAR, rel(hint, hint_ref)
    lat, ctx, te = synth.latents(B, h, h, seed=121), synth.context(B, 77, seed=122), temb(440, B)
    ctr_ref = O.controlnet_forward(cnet_sd, lat, te, ctx, hint_ref)
    ctr = engine_cnet.controlnet(lat, te, ctx, hint_ref)
    worst = max(rel(a, b) for a, b in zip(ctr, ctr_ref))
    print(f"controlnet 64x64: hint rel {rel(hint, hint_ref):.4g}, worst residual rel {worst:.4g}")
    assert worst <= EPS_BAR, worst
    ref = O.unet_forward(unet_sd, lat, te, ctx, ctr_ref)
    got = engine_unet.unet(lat, te, ctx, ctr_ref)
    assert rel(got, ref) <= EPS_BAR, rel(got, ref)


def test_config1_full_size_25_steps_against_oracle(engine_unet, engine_vae, unet_sd, vae_sd):
    """BASELINE config 1 as written: 512x512, batch 1, 25 DDIM steps, CFG 7.5 (rescale 0.7, the API default), the same
    seeded start latent on both sides.  Free-running: the engine's own loop (one captured graph) against the oracle's
    50 UNet calls; plus teacher-forced eps parity at the first, a middle and the last step of the oracle trajectory."""
    B, h, steps = 1, 64, 25
    noise, ctx, unc = synth.latents(B, h, h), synth.context(B), synth.uncond_context(B)
    trace = {}
    img_ref = O.generate_image({"unet": unet_sd, "vae": vae_sd}, ctx, unc, noise, num_steps=steps, guidance_scale=7.5,
                               guidance_rescale=0.7, trace=trace)
    assert trace["t"][0] == 960 and trace["t"][-1] == 0 and len(trace["t"]) == steps
    for i in (0, 12, 24):
        lat, t = trace["latent_in"][i], trace["t"][i]
        ec = engine_unet.unet(lat, temb(t, B), ctx)
        assert rel(ec, trace["eps_c"][i]) <= EPS_BAR, (i, t, rel(ec, trace["eps_c"][i]))
    sd = _pipeline(engine_unet, img_height=8 * h, img_width=8 * h)
    sd.unconditional_context = unc[:1]
    img, lat = sd.generate_image(ctx, batch_size=B, num_steps=steps, diffusion_noise=noise, guidance_rescale=0.7,
                                 return_latent=True)
    lat_ref = trace["latent_out"][-1]
    r, p = rel(lat, lat_ref), psnr_u8(img, img_ref)
    print(f"config 1 (64x64 latent, 25 steps) free-running: final latent rel err {r:.4g}, image psnr {p:.2f} dB")
    assert img.shape == (B, 512, 512, 3) and img.dtype == np.uint8
    assert r <= 0.1, r
    assert p >= 25.0, p
    dec = engine_vae.to_uint8(engine_vae.vae_decode(np.asarray(lat_ref, np.float32)))
    p_tf = psnr_u8(dec, img_ref)
    print(f"config 1 decode from the oracle's final latent: {p_tf:.2f} dB")
    assert p_tf >= 30.0, p_tf


def test_configs_3_4_5_full_size_free_running_against_oracle(engine_unet, engine_vae, engine_cnet, unet_sd, vae_sd, cnet_sd):
    """BASELINE configs 3 (inpaint, covers img2img), 4 (ControlNet) and 5b (TCD) at the 512x512 size — step counts cut
    (10 x 0.8 = 8, 3 and 4 steps) to keep the CPU oracle to about a minute; the loops run free (no teacher forcing)."""
    B, h, H = 1, 64, 512
    ctx, unc = synth.context(B), synth.uncond_context(B)
    noise = synth.latents(B, h, h, seed=9)
    weights = {"unet": unet_sd, "vae": vae_sd, "controlnet": cnet_sd}
    sd = _pipeline(engine_unet, img_height=H, img_width=H)
    sd.unconditional_context = unc[:1]
    src, msk = synth.smooth_image(H, H), synth.center_mask(H, H)
    in_arr, in_t = sd.preprocessed_image(src)
    m_arr, m_lat = sd.preprocessed_mask(msk, 5)
    ref = O.generate_image(weights, ctx, unc, noise, num_steps=10, guidance_scale=7.5, guidance_rescale=0.7,
                           init_latent=O.vae_encode(vae_sd, in_t), strength=0.8, latent_mask=m_lat, input_image_array=in_arr,
                           input_mask_array=m_arr, decode=False)
    _, got = sd.generate_image(ctx, batch_size=B, num_steps=10, diffusion_noise=noise, guidance_rescale=0.7, reference_image=src,
                               reference_image_strength=0.8, inpaint_mask=msk, mask_blur_strength=5, return_latent=True)
    print(f"config 3 (inpaint 512x512, 8 steps) final latent rel err {rel(got, ref):.4g}")
    assert rel(got, ref) <= 5e-2
    edge = synth.edge_map(H, H)
    hint_ref = O.hintnet_forward(cnet_sd, (edge.astype(np.float32) / 255.0)[None])
    ref = O.generate_image(weights, ctx, unc, noise, num_steps=3, guidance_scale=7.5, guidance_rescale=0.0, hint=hint_ref, decode=False)
    _, got = sd.generate_image(ctx, batch_size=B, num_steps=3, diffusion_noise=noise, control_net_image=edge, return_latent=True)
    print(f"config 4 (ControlNet 512x512, 3 steps) final latent rel err {rel(got, ref):.4g}")
    assert rel(got, ref) <= 5e-2
    sd_t = _pipeline(engine_unet, img_height=H, img_width=H, active_tcd=True)
    np.random.seed(123456)
    ref = O.generate_image(weights, ctx, None, noise, num_steps=4, guidance_scale=0.0, active_tcd=True, decode=False)
    np.random.seed(123456)
    _, got = sd_t.generate_image(ctx, batch_size=B, num_steps=4, diffusion_noise=noise, unconditional_guidance_scale=0.0,
                                 return_latent=True)
    print(f"config 5b (TCD 512x512, 4 steps) final latent rel err {rel(got, ref):.4g}")
    assert rel(got, ref) <= 5e-2


# ----------------------------------------------------------------------------------------------------------------
# the three public entry points themselves (stable_diffusion.py:84-174), token-id prompts through the engine's text tower
# ----------------------------------------------------------------------------------------------------------------
def test_public_entry_points_text_to_image_image_to_image_inpaint(engine_unet, engine_vae, unet_sd, vae_sd):
    from oracle import text_oracle as TO
    B, h, H = 1, 16, 128
    text_sd = synth.make_state_dict("text_encoder")
    if "text_encoder" not in engine_unet.loaded:
        engine_unet.load_state_dict(text_sd, "text_encoder")
    sd = _pipeline(engine_unet, img_height=H, img_width=H)
    tokens, neg = synth.prompt_tokens(1, seed=5)[0], synth.prompt_tokens(1, seed=6)[0]
    ctx_ref, unc_ref = TO.text_encode(text_sd, tokens[None], -1), TO.text_encode(text_sd, neg[None], -1)
    noise = synth.latents(B, h, h, seed=41)
    weights = {"unet": unet_sd, "vae": vae_sd}
    seen = []
    # text_to_image: defaults guidance 7.5, rescale 0.7 (:84-96); diffusion_noise is not a kwarg there, so seed= it is
    sd._get_initial_diffusion_noise = lambda batch_size, seed: noise  # the reference draws TF Philox noise here
    img = sd.text_to_image(tokens, negative_prompt=neg, batch_size=B, num_steps=4, seed=7, callback=seen.append)
    ref = O.generate_image(weights, ctx_ref, unc_ref, noise, num_steps=4, guidance_scale=7.5, guidance_rescale=0.7)
    p = psnr_u8(img, ref)
    print(f"text_to_image psnr {p:.2f} dB, callback saw {seen}")
    assert img.shape == (B, H, H, 3) and p >= 28.0, p
    assert seen == [1, 2, 3, 4]  # once per iteration, as stable_diffusion.py:476-478
    # image_to_image: 10 steps x 0.8
    src, msk = synth.smooth_image(H, H), synth.center_mask(H, H)
    in_arr, in_t = sd.preprocessed_image(src)
    init_ref = O.vae_encode(vae_sd, in_t)
    img = sd.image_to_image(tokens, negative_prompt=neg, batch_size=B, num_steps=10, seed=7, reference_image=src,
                            reference_image_strength=0.8)
    ref = O.generate_image(weights, ctx_ref, unc_ref, noise, num_steps=10, guidance_scale=7.5, guidance_rescale=0.7,
                           init_latent=init_ref, strength=0.8)
    p = psnr_u8(img, ref)
    print(f"image_to_image psnr {p:.2f} dB")
    assert p >= 28.0, p
    # inpaint
    m_arr, m_lat = sd.preprocessed_mask(msk, 5)
    img = sd.inpaint(tokens, negative_prompt=neg, batch_size=B, num_steps=10, seed=7, reference_image=src,
                     reference_image_strength=0.8, inpaint_mask=msk, mask_blur_strength=5)
    ref = O.generate_image(weights, ctx_ref, unc_ref, noise, num_steps=10, guidance_scale=7.5, guidance_rescale=0.7,
                           init_latent=init_ref, strength=0.8, latent_mask=m_lat, input_image_array=in_arr,
                           input_mask_array=m_arr)
    p = psnr_u8(img, ref)
    print(f"inpaint psnr {p:.2f} dB")
    assert p >= 28.0, p
