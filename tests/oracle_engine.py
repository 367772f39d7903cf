"""TEST INFRASTRUCTURE — an object with the Python `Engine` interface (minsdtf_b200/engine.py) whose models are the CPU
oracle.  It lets the host side of the product — `minsdtf_b200.StableDiffusion`: prompt handling, image / mask
preprocessing, timestep slicing, scheduler coefficients, the denoise descriptor — run in the CPU-only container and be
compared with the reference's own loop (tests/test_cpu_reference_harness.py).  The update it applies per step is the
documented contract of the fused kernel (include/sdtf.h, sdtf_step_coef)."""
import numpy as np

from oracle import sd15_oracle as O, text_oracle as TO
from oracle.scheduler_oracle import cfg_combine


class OracleEngine:
    def __init__(self):
        self.loaded = set()
        self.sd = {}
        self.calls = []

    def load_state_dict(self, sd, component):
        self.sd[component] = sd
        self.loaded.add(component)

    @staticmethod
    def read_checkpoint(path):
        from minsdtf_b200.engine import Engine
        return Engine.read_checkpoint(path)

    def unet(self, latent, t_emb, context, controls=None):
        return O.unet_forward(self.sd["unet"], latent, t_emb, context, controls)

    def controlnet(self, latent, t_emb, context, hint):
        return O.controlnet_forward(self.sd["controlnet"], latent, t_emb, context, hint)

    def hintnet(self, image):
        return O.hintnet_forward(self.sd["controlnet"], image)

    def vae_decode(self, latent):
        return O.vae_decode(self.sd["vae_decoder"], np.asarray(latent, np.float32))

    def vae_encode(self, image):
        return O.vae_encode(self.sd["vae_encoder"], image)

    def text_embed(self, tokens, positions=None):
        return TO.text_embed(self.sd["text_encoder"], tokens, positions)

    def text_encode(self, tokens, clip_skip=-1):
        arr = np.asarray(tokens)
        if np.issubdtype(arr.dtype, np.floating):
            return TO.encode_embedded(self.sd["text_encoder"], arr, clip_skip)
        return TO.text_encode(self.sd["text_encoder"], arr, clip_skip)

    def denoise(self, latent0, context, uncond_context, t_emb, coefs, step_noise=None, mask=None, init_latent=None,
                init_noise=None, hint_image=None, blend_image=None, blend_mask=None, decode=True, use_cuda_graph=True,
                return_latent=False, cfg_split=False, callback=None):
        self.calls.append(dict(n_steps=len(coefs), T=context.shape[1], Tu=None if uncond_context is None else uncond_context.shape[1]))
        x = np.asarray(latent0, np.float64)
        B = x.shape[0]
        hint = self.hintnet(hint_image) if hint_image is not None else None
        for i, c in enumerate(coefs):
            te = np.repeat(np.asarray(t_emb[i], np.float32)[None], B, axis=0)
            lat = np.asarray(x, np.float32)

            def call(ctx):
                ctrl = self.controlnet(lat, te, ctx, hint) if hint is not None else None
                return self.unet(lat, te, ctx, ctrl)

            if uncond_context is not None and c.guidance > 0:
                eps = cfg_combine(call(uncond_context), call(context), c.guidance, c.rescale)
            else:
                eps = call(context)
            x = c.ca * x + c.cb * eps
            if c.cn != 0.0:
                x = x + c.cn * step_noise[i]
            if mask is not None:
                m = np.asarray(mask)[None, ..., None]
                x = (c.sig_t * init_latent[None] + c.noi_t * init_noise) * (1.0 - m) + x * m
            if callback is not None:
                callback(i + 1)
        lat = np.asarray(x, np.float32)
        if not decode:
            return lat
        img = O.to_uint8(self.vae_decode(lat), None if blend_image is None else blend_image[None],
                         None if blend_mask is None else blend_mask[None, ..., None])
        return (img, lat) if return_latent else img
