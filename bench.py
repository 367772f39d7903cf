#!/usr/bin/env python
"""Headline benchmark: SD1.5 512x512 25-step txt2img images/s (BASELINE.json configs[1]: 8 prompts per GPU,
CFG 7.5, cond/uncond batched into one UNet pass of effective batch 16, random-init weights, synthetic contexts).

    python bench.py --gpus N --steps K --warmup W            # engine arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: CPU restatement of the Keras graphs

One "step" = one full generation of the per-GPU batch: 25 denoising steps (50 UNet evaluations per image) + VAE decode.
`value`  : images/s with every input already resident in HBM (torch CUDA tensors borrowed through DLPack), device time.
`e2e`    : images/s through the public API `StableDiffusion.generate_image` with HOST (NumPy) buffers — H2D of
           latents/contexts and D2H of the uint8 images are inside the timed region.
`roofline`: the dominant kernel AS THE STEP USES IT — the tcgen05 implicit-GEMM kernel (conv_gemm3_kernel: every conv and
           linear) over ALL its launches of one denoise step: summed algorithmic FLOP / summed CUDA-event time of those
           launches (sdtf_trace_begin/end: eager launches on the engine's stream, each bracketed by events), against the
           measured cuBLAS bf16 peak of MEASURED_PEAKS.json (burst; `frac_sustained` against the back-to-back figure).
           `roofline_best_shape` keeps the isolated 3x3 320->320 @64x64 number; `operator_classes` lists conv / attention /
           GroupNorm / LayerNorm totals per denoise step and for the VAE decode (GB/s of algorithmic bytes for the norms).
`strong_scaling` (N > 1): BASELINE configs[1] read literally — 8 prompts TOTAL, 8/N per GPU — measured in the same run;
           `--scaling strong` makes that the headline value instead of the weak one.
`cpu_baseline`: the oracle (torch-fp32 restatement of the reference graphs; Keras/TensorFlow are not installable
           offline) timed on this box's host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNET_GFLOP = 803.27      # per sample per UNet call @ latent 64x64, T=77
DECODER_GFLOP = 2514.52  # per image
CONV_GFLOP = 7.550       # 3x3 conv 64x64 320->320, per sample (Appendix A)
FALLBACK_PEAK_TFLOPS = 1590.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="prompts per GPU")
    ap.add_argument("--denoise-steps", type=int, default=25)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--profile-only", action="store_true", help="run the resident step --steps times and exit (for ncu)")
    ap.add_argument("--no-cfg-split", action="store_true", help="skip the extra CFG-split (GPU pairs + NCCL) measurement at N >= 2")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch prompts per GPU (default); strong: --batch prompts in total, batch/N per GPU")
    ap.add_argument("--no-strong", action="store_true", help="skip the extra strong-scaling measurement at N >= 2")
    ap.add_argument("--acct-graph", action="store_true",
                    help="operator accounting inside the captured step graph (external event-record nodes) instead of eager launches")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm, reasons, mx = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel_label):
    """DRAM bytes per launch of the roofline kernel from the committed ncu --set full capture (profiles/ncu_traffic.json)"""
    try:
        return int(json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel_label]["traffic"])
    except Exception:
        return None


def measured_hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6400.0, "fallback (B200_PROFILING.md)"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
        except Exception:
            pass
    return FALLBACK_PEAK_TFLOPS, "fallback (B200_PROFILING.md)"


def measured_sustained_peak():
    """cuBLAS bf16 throughput back to back for seconds (power-capped clocks): the right denominator for a whole step"""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(denoise_steps, size, reps=1):
    """Times UNet calls (B=1) and one VAE decode of the oracle on all host cores; returns images/s for the full
    25-step CFG workload extrapolated as 1 / (2*steps*t_unet + t_decode), plus a description of the sample."""
    import torch
    from minsdtf_b200 import synth
    from oracle import sd15_oracle as O
    from oracle.scheduler_oracle import timestep_embedding
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    h = size // 8
    unet, vae = synth.make_state_dict("unet"), synth.make_state_dict("decoder")
    lat, ctx, te = synth.latents(1, h, h), synth.context(1), timestep_embedding(500, 1)
    O.unet_forward(unet, lat, te, ctx)  # warm-up
    t0 = time.perf_counter()
    for _ in range(max(1, reps) * 2):
        O.unet_forward(unet, lat, te, ctx)
    t_unet = (time.perf_counter() - t0) / (max(1, reps) * 2)
    t0 = time.perf_counter()
    O.vae_decode(vae, lat * 0.18215)
    t_dec = time.perf_counter() - t0
    per_image = 2 * denoise_steps * t_unet + t_dec
    sample = (f"{2 * max(1, reps)} UNet calls (B=1, {size}x{size}) at {t_unet * 1e3:.0f} ms + 1 VAE decode at {t_dec * 1e3:.0f} ms on "
              f"{cores} threads, extrapolated to {2 * denoise_steps} UNet calls + 1 decode per image")
    return 1.0 / per_image, cores, sample, t_unet, t_dec


def run_reference(args, rank):
    """Reference arm: each step is a bounded sample of the workload — one UNet call (B=1) of the CPU restatement on all
    host cores; the VAE decode is timed once up front.  images/s = 1 / (2*steps*t_unet + t_decode)."""
    if rank != 0:
        return
    import torch
    from minsdtf_b200 import synth
    from oracle import sd15_oracle as O
    from oracle.scheduler_oracle import timestep_embedding
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    h = args.size // 8
    unet, vae = synth.make_state_dict("unet"), synth.make_state_dict("decoder")
    lat, ctx, te = synth.latents(1, h, h), synth.context(1), timestep_embedding(500, 1)
    t0 = time.perf_counter()
    O.vae_decode(vae, lat * 0.18215)
    t_dec = time.perf_counter() - t0
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.unet_forward(unet, lat, te, ctx)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    t_unet = float(np.mean(times))
    v = 1.0 / (2 * args.denoise_steps * t_unet + t_dec)
    meta = (cores, f"per step: 1 UNet call (B=1, {args.size}x{args.size}) = {t_unet * 1e3:.0f} ms; 1 VAE decode = {t_dec * 1e3:.0f} ms; "
                   f"{cores} threads; extrapolated to {2 * args.denoise_steps} UNet calls + 1 decode per image")
    line = {"impl": "reference", "metric": "images_per_second", "value": v, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / v if v else None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"SD1.5 txt2img {args.size}x{args.size}, {args.denoise_steps} DDIM steps, CFG 7.5, random-init weights; "
                                   "CPU restatement of the reference Keras graphs (keras/tensorflow not installable offline)"},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": meta[0], "kind": "port", "sample": meta[1]},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------------------------
# engine arm
# ----------------------------------------------------------------------------------------------------------------
def run_engine(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from minsdtf_b200 import synth
    from minsdtf_b200.scheduler import timestep_embedding
    from minsdtf_b200.stable_diffusion import StableDiffusion

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl engine needs a B200: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B, S, size = args.batch, args.denoise_steps, args.size
    if args.scaling == "strong":
        if args.batch % world:
            raise SystemExit(f"--scaling strong: {args.batch} prompts do not divide over {world} GPUs")
        B = args.batch // world
    h = size // 8
    sd = StableDiffusion(img_height=size, img_width=size, synthetic=True, device=local_rank)
    sd.load_all()
    eng = sd.engine
    # distinct prompts per rank (data parallel over prompts: no per-step communication)
    noise = synth.latents(B, h, h, seed=123456 + rank)
    ctx = synth.context(B, 77, seed=123457 + rank)
    unc = synth.uncond_context(B, 77)
    sd.unconditional_context = unc[:1]
    sd.scheduler.set_timesteps(S)
    exec_ts = [int(t) for t in sd.scheduler.timesteps]
    coefs = sd.scheduler.coefficients(exec_ts, 7.5, 0.7)
    t_emb = np.stack([timestep_embedding(t) for t in exec_ts])
    dev = torch.device("cuda", local_rank)
    d_noise, d_ctx, d_unc, d_temb = (torch.as_tensor(a).to(dev) for a in (noise, ctx, unc, t_emb))
    use_graph = not args.no_graph

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        imgs = eng.denoise(d_noise, d_ctx, d_unc, d_temb, coefs, decode=True, use_cuda_graph=use_graph)
        return imgs, eng.timings()

    def e2e_step():
        return sd.generate_image(ctx, batch_size=B, num_steps=S, unconditional_guidance_scale=7.5, diffusion_noise=noise,
                                 guidance_rescale=0.7, use_cuda_graph=use_graph)

    if args.profile_only:
        for _ in range(args.warmup + args.steps):
            resident_step()
        torch.cuda.synchronize()
        emit({"profile_only": True, "timings": eng.timings()})
        return

    # ---- device-resident arm ----
    for _ in range(args.warmup):
        resident_step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    dev_ms, loop_ms, dec_ms, launches = 0.0, 0.0, 0.0, 0
    for _ in range(args.steps):
        _, tm = resident_step()
        dev_ms += tm["total_ms"]; loop_ms += tm["loop_ms"]; dec_ms += tm["decode_ms"]; launches += tm["kernel_launches"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    # ---- end-to-end arm (host buffers through the public API) ----
    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = e2e_step()
    barrier()
    wall_e2e = time.perf_counter() - t0
    assert out.shape == (B, size, size, 3) and out.dtype == np.uint8

    # ---- CFG-split variant (N >= 2, even): GPU pairs, each rank one CFG branch, one ncclAllGather of eps per step ----
    split_info = None
    if world >= 2 and world % 2 == 0 and not args.no_cfg_split:
        from minsdtf_b200 import dist as D
        plan = D.setup_cfg_split(eng, rank, world, device=dev)
        Bp = 2 * B  # prompts per pair: every GPU still runs a UNet batch of 2B per step, as in the data-parallel arm
        p_noise = torch.as_tensor(synth.latents(Bp, h, h, seed=223456 + plan.pair)).to(dev)
        p_ctx = torch.as_tensor(synth.context(Bp, 77, seed=223457 + plan.pair)).to(dev)
        p_unc = torch.as_tensor(synth.uncond_context(Bp, 77)).to(dev)
        half = slice(plan.branch * B, (plan.branch + 1) * B)

        def split_step():
            lat = eng.denoise(p_noise, p_ctx, p_unc, d_temb, coefs, decode=False, use_cuda_graph=use_graph, cfg_split=True)
            return eng.to_uint8(eng.vae_decode(lat[half].contiguous()))  # each member decodes its half of the pair's images

        for _ in range(max(1, min(args.warmup, 2))):
            split_step()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            imgs = split_step()
        barrier()
        wall_split = time.perf_counter() - t0
        ts = torch.tensor([wall_split], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        split_info = {"value": Bp * (world // 2) * args.steps / ts.item(), "unit": "images/s", "ms_per_step": ts.item() / args.steps * 1e3,
                      "pairs": world // 2, "prompts_per_pair": Bp, "exchange": "ncclAllGather of eps (B,64,64,4) f32 per step, in-graph",
                      "timing": "host wall clock around K generations incl. barrier (max over ranks)"}
        assert tuple(imgs.shape) == (B, size, size, 3)
        # self-check: the split loop must reproduce the single-GPU batched loop on the same inputs (4 steps, eager and graph)
        worst = 0.0
        for graph in (False, True):
            a = eng.denoise(p_noise[:2], p_ctx[:2], p_unc[:2], d_temb[:4], coefs[:4], decode=False, use_cuda_graph=graph, cfg_split=True)
            b = eng.denoise(p_noise[:2], p_ctx[:2], p_unc[:2], d_temb[:4], coefs[:4], decode=False, use_cuda_graph=graph, cfg_split=False)
            worst = max(worst, float((a - b).abs().max()))
        wt = torch.tensor([worst], dtype=torch.float64, device=dev)
        dist.all_reduce(wt, op=dist.ReduceOp.MAX)
        split_info["max_abs_diff_vs_unsplit_loop"] = wt.item()
        split_info["equals_unsplit_loop"] = bool(wt.item() <= 1e-5)
        assert split_info["equals_unsplit_loop"], f"CFG split differs from the single-GPU loop by {wt.item()}"
        eng.comm_destroy()

    # ---- strong scaling (BASELINE configs[1] read literally: `--batch` prompts in TOTAL, batch/N per GPU) ----
    strong_info = None
    if world >= 2 and args.scaling == "weak" and not args.no_strong and args.batch % world == 0:
        Bs = args.batch // world
        s_noise, s_ctx, s_unc = d_noise[:Bs].contiguous(), d_ctx[:Bs].contiguous(), d_unc[:Bs].contiguous()
        for _ in range(max(1, min(args.warmup, 3))):
            eng.denoise(s_noise, s_ctx, s_unc, d_temb, coefs, decode=True, use_cuda_graph=use_graph)
        barrier()
        s_ms = 0.0
        for _ in range(args.steps):
            eng.denoise(s_noise, s_ctx, s_unc, d_temb, coefs, decode=True, use_cuda_graph=use_graph)
            s_ms += eng.timings()["total_ms"]
        barrier()
        st = torch.tensor([s_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
        strong_info = {"value": args.batch * args.steps / (st.item() / 1e3), "unit": "images/s", "scaling": "strong",
                       "prompts_total": args.batch, "prompts_per_gpu": Bs, "unet_batch_per_gpu": 2 * Bs,
                       "ms_per_step": st.item() / args.steps, "timing": "device time of the whole call (CUDA events), max over ranks"}

    # ---- max over ranks ----
    t = torch.tensor([dev_ms, wall, wall_e2e, loop_ms, dec_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall, wall_e2e, loop_ms, dec_ms = t.tolist()
    K = args.steps
    images = B * world * K
    value = images / (dev_ms / 1e3)
    e2e = images / wall_e2e
    unet_step_ms = loop_ms / (K * S)
    unet_tflops = (2 * B * UNET_GFLOP) / unet_step_ms  # GFLOP / ms == TFLOP/s
    peak, peak_src = measured_peak()

    if rank == 0:
        # ---- operator classes of ONE denoise step and of the VAE decode: eager launches, CUDA events around every operator ----
        # Default: two EAGER steps with an event pair around every operator, read back after the job (kernels of ~20 us are then
        # launch-bound on the host, so short operators read a few us long: a conservative number).  --acct-graph re-captures
        # the step graph with external event-record nodes instead; measured 868 and 940 TFLOP/s for the conv class in two
        # runs against 909 / 934 eager — the event nodes perturb the replay more than the eager gaps do.
        acct_graph = use_graph and args.acct_graph
        eng.trace_begin()
        eng.denoise(d_noise, d_ctx, d_unc, d_temb[:2], coefs[:2], decode=False, use_cuda_graph=acct_graph)
        tr_unet = eng.trace_end()
        n_rec = 1 if acct_graph else 2  # one captured step (the last replay is read) / two eager steps
        eng.trace_begin()
        eng.vae_decode(d_noise * 0.18215)
        tr_dec = eng.trace_end()
        hbm_peak, hbm_src = measured_hbm_peak()
        sustained = measured_sustained_peak()

        def classes(tr, div):
            out, tot = {}, sum(v["us"] for v in tr.values())
            for k, v in tr.items():
                if not v["launches"]:
                    continue
                e = {"launches": v["launches"] // div, "ms": v["us"] / div / 1e3, "share_of_operator_time": v["us"] / tot}
                if k in ("conv", "attn"):
                    e["tflops"] = v["flop"] / v["us"] / 1e6
                    e["frac_of_burst_peak"] = e["tflops"] / peak
                else:
                    e["gbs"] = v["bytes"] / v["us"] / 1e3
                    e["frac_of_hbm_peak"] = e["gbs"] / hbm_peak
                out[k] = e
            return out

        cls_unet, cls_dec = classes(tr_unet, n_rec), classes(tr_dec, 1)
        conv = tr_unet["conv"]
        conv_step_tflops = conv["flop"] / conv["us"] / 1e6
        # the isolated best shape (what round 1 reported as `roofline`), kept for continuity
        conv_ms = eng.bench_conv(2 * B, 64, 320, 320, 3, reps=40)
        conv_tflops = CONV_GFLOP * 2 * B / conv_ms
        conv_label = "conv_gemm3_kernel 3x3 320->320 @64x64, batch %d" % (2 * B)
        par = f"dp{world} over prompts, no per-step collective"
        if args.scaling == "strong":
            par += f"; strong scaling: {args.batch} prompts in total, {B} per GPU"
        line = {
            "metric": "images_per_second", "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"SD1.5 text_to_image {size}x{size}, {B} prompts per GPU, {S} DDIM steps, CFG 7.5 "
                                   f"(cond+uncond batched: UNet batch {2 * B}), guidance_rescale 0.7, random-init weights, "
                                   f"synthetic contexts (BASELINE.json configs[1])",
                       "parallelism": par,
                       "l2": "per-iteration working set (1.7 GB bf16 weights + GBs of activations) >> 126 MB L2; no flush needed",
                       "cuda_graph": use_graph},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": int(noise[:B].nbytes + ctx[:B].nbytes + unc[:B].nbytes + t_emb.nbytes),
                    "d2h_bytes_per_step": int(B * size * size * 3), "ms_per_step": wall_e2e / K * 1e3},
            "gpu_launches": int(launches),
            "unet_step_ms": unet_step_ms, "unet_step_tflops": unet_tflops, "unet_step_frac_of_peak": unet_tflops / peak,
            "unet_step_frac_of_sustained_peak": (unet_tflops / sustained) if sustained else None,
            "decode_ms_per_batch": dec_ms / K, "wall_ms_per_step": wall / K * 1e3,
            "roofline": {"bound": "tensor", "achieved": conv_step_tflops, "peak": peak, "unit": "TFLOP/s", "frac": conv_step_tflops / peak,
                         "frac_sustained": (conv_step_tflops / sustained) if sustained else None,
                         "traffic": ncu_traffic("conv_gemm3_kernel step average"),
                         "kernel": "conv_gemm3_kernel: all %d conv / linear launches of one denoise step (UNet batch %d), FLOP-weighted"
                                   % (conv["launches"] // n_rec, 2 * B),
                         "flop_per_launch": conv["flop"] / conv["launches"], "ms_per_launch": conv["us"] / conv["launches"] / 1e3,
                         "ms_per_step_in_kernel": conv["us"] / n_rec / 1e3,
                         "share_of_operator_time": conv["us"] / sum(v["us"] for v in tr_unet.values()),
                         "how": "sdtf_trace_begin/end: the step graph re-captured with an external CUDA event pair around every operator on the engine's stream, replayed, events read back after the job" if acct_graph else "sdtf_trace_begin/end: eager launches on the engine's stream, a CUDA event pair around every operator, read back after the job",
                         "peak_source": peak_src},
            "roofline_best_shape": {"bound": "tensor", "achieved": conv_tflops, "peak": peak, "unit": "TFLOP/s", "frac": conv_tflops / peak,
                                    "traffic": ncu_traffic(conv_label), "kernel": conv_label,
                                    "flop_per_launch": CONV_GFLOP * 2 * B * 1e9, "ms_per_launch": conv_ms},
            "operator_classes": {"denoise_step": cls_unet, "vae_decode": cls_dec, "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src},
            "clocks": clocks,
        }
        if split_info is not None:
            line["cfg_split"] = split_info
        if strong_info is not None:
            line["strong_scaling"] = strong_info
        if not args.skip_cpu_baseline and world == 1:
            v, cores, sample, _, _ = cpu_reference_rate(S, size)
            line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
        else:
            line["cpu_baseline"] = None
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """the ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version banner, torchrun
    notices) was redirected to stderr in main()"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # C-level writes to fd 1 (e.g. "NCCL version ...") must not pollute the JSON line
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_engine(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
