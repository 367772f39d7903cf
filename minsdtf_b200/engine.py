"""Python handle on one GPU's engine (libsdtf.so).  Mirrors the reference's model objects: every model is reached
through a `.predict_on_batch(list_of_arrays)`-shaped method that takes / returns NHWC float32 arrays, exactly the
seam the reference pipeline uses (stable_diffusion.py:415,439,447-457,482).

Inputs may be NumPy arrays (host: the engine does the H2D/D2H copies) or torch CUDA tensors (borrowed in place via
DLPack).  Outputs follow the input kind.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib, keys as K

_CTRL_LEVEL = [0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 3]
_CTRL_CH = [320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280, 1280]


class EngineError(RuntimeError):
    pass


def _alias_to_ldm():
    return {v: k for k, v in K.unet_alias_map().items()}


class Engine:
    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self._lib.sdtf_create(int(device), ctypes.byref(h))
        if rc != 0:
            msg = self._lib.sdtf_last_error(None)
            raise EngineError(f"sdtf_create failed ({rc}): {msg.decode() if msg else ''}")
        self._h = h
        self.device = int(device)
        self.loaded = set()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sdtf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.sdtf_last_error(self._h)
            raise EngineError(f"libsdtf error {rc}: {msg.decode() if msg else ''}")

    # ------------------------------------------------------------------------------------------------ weights
    def load_state_dict(self, sd: dict, component: str):
        """component: 'unet' | 'controlnet' | 'vae_decoder' | 'vae_encoder' | 'text_encoder'.  `sd` maps reference checkpoint keys
        (or, for the UNet, their diffusers aliases — ckpt_loader.py:2160-2166) to tensors in PyTorch layout."""
        gens = {"unet": [K.unet_keys], "controlnet": [K.controlnet_keys, K.hintnet_keys],
                "vae_decoder": [K.vae_decoder_keys], "vae_encoder": [K.vae_encoder_keys],
                "text_encoder": [K.text_encoder_keys]}[component]
        wanted = {}
        for g in gens:
            wanted.update(g())
        alias = K.unet_alias_map() if component == "unet" else {}
        missing = []
        for key, shape in wanted.items():
            t = sd.get(key)
            if t is None and key in alias:
                t = sd.get(alias[key])
            if t is None:
                missing.append(key)
                continue
            if tuple(t.shape) != tuple(shape):
                raise EngineError(f"{key}: shape {tuple(t.shape)} != expected {tuple(shape)}")
            dl = _lib.DL()
            if _lib.is_torch(t):
                import torch
                if t.dtype not in (torch.float32, torch.float16, torch.bfloat16):
                    t = t.float()
            else:
                t = np.asarray(t, dtype=np.float32)
            self._check(self._lib.sdtf_load_tensor(self._h, key.encode(), dl(t)))
        if missing:
            raise EngineError(f"{component}: {len(missing)} checkpoint tensors missing, e.g. {missing[:3]}")
        self._check(self._lib.sdtf_finalize_weights(self._h, component.encode()))
        self.loaded.add(component)

    @staticmethod
    def read_checkpoint(path: str) -> dict:
        """.safetensors or torch pickle (.ckpt / .pth, optionally wrapped in {"state_dict": ...}) -> {key: tensor}, as the
        reference's loader reads them (ckpt_loader.py:2139-2145)."""
        if path.endswith(".safetensors"):
            from safetensors import safe_open
            sd = {}
            with safe_open(path, framework="pt", device="cpu") as f:
                for k in f.keys():
                    sd[k] = f.get_tensor(k)
            return sd
        import torch
        sd = torch.load(path, map_location="cpu")
        if "state_dict" in sd:
            sd = sd["state_dict"]
        return sd

    def load_file(self, path: str, component: str):
        self.load_state_dict(self.read_checkpoint(path), component)

    # ------------------------------------------------------------------------------------------------ helpers
    @staticmethod
    def _like(x, shape, dtype=np.float32):
        if _lib.is_torch(x):
            import torch
            td = {np.float32: torch.float32, np.uint8: torch.uint8}[dtype]
            return torch.empty(shape, dtype=td, device=x.device)
        return np.empty(shape, dtype=dtype)

    @staticmethod
    def _f32(x):
        if x is None:
            return None
        if _lib.is_torch(x):
            import torch
            return x.to(torch.float32).contiguous()
        return np.ascontiguousarray(x, dtype=np.float32)

    # ------------------------------------------------------------------------------------------------ models
    def unet(self, latent, t_emb, context, controls=None):
        """DiffusionModel.predict_on_batch([latent, t_emb, context] + controls) (diffusion_model.py:163-283)."""
        latent, t_emb, context = self._f32(latent), self._f32(t_emb), self._f32(context)
        out = self._like(latent, tuple(latent.shape))
        dl = _lib.DL()
        carr = None
        if controls is not None:
            if len(controls) != 13:
                raise EngineError("controls must be the 13 ControlNet residuals")
            carr = dl.array([self._f32(c) for c in controls])
        self._check(self._lib.sdtf_unet_forward(self._h, dl(latent), dl(t_emb), dl(context), carr, dl(out)))
        return out

    def controlnet(self, latent, t_emb, context, hint):
        """ControlNet.predict_on_batch([latent, t_emb, context, hint]) -> 13 residuals (control_net.py:45-107)."""
        latent, t_emb, context, hint = self._f32(latent), self._f32(t_emb), self._f32(context), self._f32(hint)
        B, h, w, _ = latent.shape
        outs = [self._like(latent, (B, h >> lv, w >> lv, c)) for lv, c in zip(_CTRL_LEVEL, _CTRL_CH)]
        dl = _lib.DL()
        self._check(self._lib.sdtf_controlnet_forward(self._h, dl(latent), dl(t_emb), dl(context), dl(hint), dl.array(outs)))
        return outs

    def hintnet(self, image):
        """HintNet.predict_on_batch(image in [0,1]) (control_net.py:10-31)."""
        image = self._f32(image)
        B, H, W, _ = image.shape
        out = self._like(image, (B, H // 8, W // 8, 320))
        dl = _lib.DL()
        self._check(self._lib.sdtf_hintnet_forward(self._h, dl(image), dl(out)))
        return out

    def vae_decode(self, latent):
        """ImageDecoder.predict_on_batch(latent) (image_decoder.py:22-55)."""
        latent = self._f32(latent)
        B, h, w, _ = latent.shape
        out = self._like(latent, (B, 8 * h, 8 * w, 3))
        dl = _lib.DL()
        self._check(self._lib.sdtf_vae_decode(self._h, dl(latent), dl(out)))
        return out

    def vae_encode(self, image):
        """ImageEncoder.predict_on_batch(image in [-1,1]) (image_encoder.py:21-48)."""
        image = self._f32(image)
        B, H, W, _ = image.shape
        out = self._like(image, (B, H // 8, W // 8, 4))
        dl = _lib.DL()
        self._check(self._lib.sdtf_vae_encode(self._h, dl(image), dl(out)))
        return out

    def text_encode(self, tokens, clip_skip: int = -1):
        """TextEncoder(TextClipEmbedding([tokens, positions])) (text_encoder.py:106-135): int tokens (B,T<=77) ->
        context (B,T,768) float32 (NumPy).  A float (B,T,768) input is taken as the output of `text_embed` (possibly
        with textual-inversion vectors spliced in) and goes through TextEncoder only (text_encoder.py:125-135)."""
        arr = np.asarray(tokens)
        if np.issubdtype(arr.dtype, np.floating):
            emb = np.ascontiguousarray(arr, dtype=np.float32)
            if emb.ndim == 2:
                emb = emb[None]
            out = np.empty(emb.shape, np.float32)
            dl = _lib.DL()
            self._check(self._lib.sdtf_text_encode_embedded(self._h, dl(emb), int(clip_skip), dl(out)))
            return out
        tokens = np.ascontiguousarray(arr, dtype=np.int32)
        if tokens.ndim == 1:
            tokens = tokens[None]
        out = np.empty(tokens.shape + (768,), np.float32)
        dl = _lib.DL()
        self._check(self._lib.sdtf_text_encode(self._h, dl(tokens), int(clip_skip), dl(out)))
        return out

    def text_embed(self, tokens, positions=None):
        """TextClipEmbedding.predict_on_batch([tokens, positions]) (text_encoder.py:22-33,106-122) -> (B,T,768) float32."""
        tokens = np.ascontiguousarray(np.asarray(tokens, dtype=np.int32))
        if tokens.ndim == 1:
            tokens = tokens[None]
        if positions is not None:
            positions = np.ascontiguousarray(np.asarray(positions, dtype=np.int32))
            if positions.ndim == 1:
                positions = positions[None]
        out = np.empty(tokens.shape + (768,), np.float32)
        dl = _lib.DL()
        self._check(self._lib.sdtf_text_embed(self._h, dl(tokens), dl(positions), dl(out)))
        return out

    def cfg_sched_step(self, eps_u, eps_c, latent_prev, coef, noise=None, mask=None, init_latent=None, init_noise=None):
        """Fused CFG combine + rescale + Scheduler.step (+ inpaint blend) (stable_diffusion.py:458-475)."""
        eps_c, latent_prev = self._f32(eps_c), self._f32(latent_prev)
        out = self._like(eps_c, tuple(eps_c.shape))
        dl = _lib.DL()
        sc = coef if isinstance(coef, _lib.StepCoef) else _lib.StepCoef(*coef)
        self._check(self._lib.sdtf_cfg_sched_step(self._h, dl(self._f32(eps_u)), dl(eps_c), dl(latent_prev), ctypes.byref(sc),
                                                  dl(self._f32(noise)), dl(self._f32(mask)), dl(self._f32(init_latent)),
                                                  dl(self._f32(init_noise)), dl(out)))
        return out

    def to_uint8(self, decoded, blend_image=None, blend_mask=None):
        decoded = self._f32(decoded)
        out = self._like(decoded, tuple(decoded.shape), np.uint8)
        dl = _lib.DL()
        self._check(self._lib.sdtf_to_uint8(self._h, dl(decoded), dl(self._f32(blend_image)), dl(self._f32(blend_mask)), dl(out)))
        return out

    # ------------------------------------------------------------------------------------------------ 2-way CFG split
    def comm_unique_id(self, nccl_lib=None) -> bytes:
        buf = ctypes.create_string_buffer(128)
        rc = self._lib.sdtf_comm_unique_id(nccl_lib.encode() if nccl_lib else None, buf)
        if rc != 0:
            msg = self._lib.sdtf_last_error(None)
            raise EngineError(f"sdtf_comm_unique_id failed ({rc}): {msg.decode() if msg else ''}")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int = 2, nccl_lib=None):
        """Join the 2-rank NCCL communicator of a CFG pair (rank 0 = unconditional branch, 1 = conditional)."""
        if len(unique_id) != 128:
            raise EngineError("NCCL unique id must be 128 bytes")
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self._check(self._lib.sdtf_comm_init(self._h, nccl_lib.encode() if nccl_lib else None, buf, int(rank), int(world)))
        self.cfg_rank = int(rank)

    def comm_destroy(self):
        self._check(self._lib.sdtf_comm_destroy(self._h))
        self.cfg_rank = None

    def denoise(self, latent0, context, uncond_context, t_emb, coefs, step_noise=None, mask=None, init_latent=None,
                init_noise=None, hint_image=None, blend_image=None, blend_mask=None, decode=True, use_cuda_graph=True,
                return_latent=False, cfg_split=False, callback=None):
        """The whole loop of generate_image (stable_diffusion.py:442-486) in one call.  coefs: list of StepCoef in
        execution order; t_emb (n_steps, 320).  callback(iteration): called after every step (:476-478)."""
        latent0 = self._f32(latent0)
        B, h, w, _ = latent0.shape
        n = len(coefs)
        arr = (_lib.StepCoef * n)(*coefs)
        dl = _lib.DL()
        d = _lib.DenoiseDesc()
        d.n_steps, d.use_cuda_graph, d.decode = n, int(bool(use_cuda_graph)), int(bool(decode))
        d.cfg_split = int(bool(cfg_split))
        d.latent0 = dl(latent0)
        d.context = dl(self._f32(context))
        d.uncond_context = dl(self._f32(uncond_context))
        d.t_emb = dl(self._f32(t_emb))
        d.coefs = arr
        d.step_noise = dl(self._f32(step_noise))
        d.mask = dl(self._f32(mask))
        d.init_latent = dl(self._f32(init_latent))
        d.init_noise = dl(self._f32(init_noise))
        d.hint_image = dl(self._f32(hint_image))
        d.blend_image = dl(self._f32(blend_image))
        d.blend_mask = dl(self._f32(blend_mask))
        images = self._like(latent0, (B, 8 * h, 8 * w, 3), np.uint8) if decode else None
        lat = self._like(latent0, (B, h, w, 4)) if (return_latent or not decode) else None
        d.out_images = dl(images)
        d.out_latent = dl(lat)
        raised = []
        if callback is not None:
            def _on_step(iteration, _user):
                if raised:
                    return
                try:
                    callback(int(iteration))
                except BaseException as ex:  # an exception cannot unwind through the C frames: re-raised after the call
                    raised.append(ex)
            cb = _lib.ON_STEP(_on_step)
            d.on_step = ctypes.cast(cb, ctypes.c_void_p)
        self._check(self._lib.sdtf_denoise(self._h, ctypes.byref(d)))
        if raised:
            raise raised[0]
        if decode and return_latent:
            return images, lat
        return images if decode else lat

    def timings(self) -> dict:
        t = _lib.Timings()
        self._check(self._lib.sdtf_get_timings(self._h, ctypes.byref(t)))
        return {"loop_ms": t.loop_ms, "decode_ms": t.decode_ms, "total_ms": t.total_ms, "kernel_launches": t.kernel_launches}

    def trace_begin(self):
        """Start per-operator-class accounting (eager launches bracketed by CUDA events; see include/sdtf.h)."""
        self._check(self._lib.sdtf_trace_begin(self._h))

    def trace_end(self) -> dict:
        t = _lib.TraceSummary()
        self._check(self._lib.sdtf_trace_end(self._h, ctypes.byref(t)))
        return {k: {"launches": int(t.launches[i]), "us": float(t.us[i]), "flop": float(t.flop[i]), "bytes": float(t.bytes[i])}
                for i, k in enumerate(("conv", "attn", "gn", "ln"))}

    def bench_attention(self, batch, heads, nq, nk, d, reps=20, legacy=False) -> float:
        ms = ctypes.c_float()
        self._check(self._lib.sdtf_bench_attention(self._h, batch, heads, nq, nk, d, reps, int(legacy), ctypes.byref(ms)))
        return float(ms.value)

    def bench_conv(self, batch, hw, cin, cout, ksize=3, reps=20) -> float:
        ms = ctypes.c_float()
        self._check(self._lib.sdtf_bench_conv(self._h, batch, hw, cin, cout, ksize, reps, ctypes.byref(ms)))
        return float(ms.value)
