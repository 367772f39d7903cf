"""`StableDiffusion` with the reference's Python surface (stable_diffusion/stable_diffusion.py:47-725) on top of the
B200 engine.

Kept verbatim from the reference: constructor kwargs (:620-631), `text_to_image / image_to_image / inpaint /
generate_image` signatures and defaults (:84-174, 317-334), the seven model properties reached only through
`.predict_on_batch` (:505-531), `scheduler`, the host-side image / mask preprocessing (:217-302) and the quirks of the
loop (img2img timestep slicing :406-416, inpaint re-noising at the *current* t :469-475, uint8 truncation :486).

Different by design: the loop body (2 UNet calls (+2 ControlNet calls) + CFG + scheduler step per iteration, :442-475)
and the decode (:482-486) run on the GPU inside one `sdtf_denoise` call; cond and uncond are batched into one UNet
pass; cross-attention K/V and HintNet features are computed once per image.

Out of scope this round: the CLIP text tower / tokenizer.  `encode_text` therefore needs a
user-supplied `text_encoder_fn`; `generate_image(encoded_text, ...)` — the reference's own lower-level entry — is
fully supported, as are precomputed embeddings passed as `prompt`.
"""
from __future__ import annotations

import os

import numpy as np

from .engine import Engine
from .scheduler import Scheduler, timestep_embedding

MAX_PROMPT_LENGTH = 77


class _Model:
    """Object with the reference models' `.predict_on_batch` seam, backed by an engine entry point."""

    def __init__(self, fn):
        self._fn = fn

    def predict_on_batch(self, inputs):
        return self._fn(inputs)

    __call__ = predict_on_batch


class StableDiffusionBase:
    def __init__(self, img_height=512, img_width=512, jit_compile=False, active_tcd=False):
        self.img_height = img_height
        self.img_width = img_width
        self.jit_compile = jit_compile  # accepted for compatibility; the engine is always compiled
        self.active_tcd = active_tcd
        self.scheduler = Scheduler(active_tcd=active_tcd)

    # ---------------------------------------------------------------------------------- public entry points
    def text_to_image(self, prompt, negative_prompt=None, batch_size=1, num_steps=50, unconditional_guidance_scale=7.5,
                      embedding=None, negative_embedding=None, seed=None, control_net_image=None, guidance_rescale=0.7,
                      callback=None):
        encoded_text = self.encode_text(prompt, embedding)
        return self.generate_image(encoded_text, negative_prompt=negative_prompt, batch_size=batch_size, num_steps=num_steps,
                                   unconditional_guidance_scale=unconditional_guidance_scale, seed=seed,
                                   negative_embedding=negative_embedding, control_net_image=control_net_image,
                                   guidance_rescale=guidance_rescale, callback=callback)

    def image_to_image(self, prompt, negative_prompt=None, batch_size=1, num_steps=50, unconditional_guidance_scale=7.5,
                       embedding=None, negative_embedding=None, seed=None, control_net_image=None, reference_image=None,
                       reference_image_strength=0.8, guidance_rescale=0.7, callback=None):
        encoded_text = self.encode_text(prompt, embedding)
        return self.generate_image(encoded_text, negative_prompt=negative_prompt, batch_size=batch_size, num_steps=num_steps,
                                   unconditional_guidance_scale=unconditional_guidance_scale, seed=seed,
                                   negative_embedding=negative_embedding, control_net_image=control_net_image,
                                   reference_image=reference_image, reference_image_strength=reference_image_strength,
                                   guidance_rescale=guidance_rescale, callback=callback)

    def inpaint(self, prompt, negative_prompt=None, batch_size=1, num_steps=50, unconditional_guidance_scale=7.5,
                embedding=None, negative_embedding=None, seed=None, control_net_image=None, reference_image=None,
                reference_image_strength=0.8, inpaint_mask=None, mask_blur_strength=None, guidance_rescale=0.7,
                callback=None):
        encoded_text = self.encode_text(prompt, embedding)
        return self.generate_image(encoded_text, negative_prompt=negative_prompt, batch_size=batch_size, num_steps=num_steps,
                                   unconditional_guidance_scale=unconditional_guidance_scale, seed=seed,
                                   negative_embedding=negative_embedding, control_net_image=control_net_image,
                                   reference_image=reference_image, reference_image_strength=reference_image_strength,
                                   inpaint_mask=inpaint_mask, mask_blur_strength=mask_blur_strength,
                                   guidance_rescale=guidance_rescale, callback=callback)

    def encode_text(self, prompt, embedding_data=None):
        """Reference: tokenizer -> TextClipEmbedding -> TextEncoder (:176-215).  Accepted here:
        * a float array (T,768) / (B,T,768): already-encoded text, passed through;
        * an integer array of CLIP token ids (T,) / (B,T<=77): encoded by the engine's text tower (`clip_skip` as given
          to the constructor);
        * a string: tokenised by `model.tokenizer` (any object with `.encode(str) -> ids`; the BPE vocabulary is a
          download in the reference, clip_tokenizer.py:79-82, and is not available offline), padded with 49407 to 77.
        Long-prompt weighting and textual-inversion embeddings (`embedding_data`) are host-side features outside the
        hot path."""
        if embedding_data is not None:
            raise NotImplementedError("textual-inversion embeddings are outside the hot path (SURVEY.md §2)")
        if isinstance(prompt, np.ndarray) and np.issubdtype(prompt.dtype, np.floating):
            return prompt.astype(np.float32)
        if isinstance(prompt, str):
            tok = getattr(self, "tokenizer", None)
            if tok is None:
                fn = getattr(self, "text_encoder_fn", None)
                if fn is not None:
                    return np.asarray(fn(prompt), dtype=np.float32)
                raise NotImplementedError(
                    "no tokenizer: the CLIP BPE vocabulary cannot be downloaded offline — set `model.tokenizer` (an object with "
                    ".encode(str) -> token ids), or pass token ids / an encoded (T,768) array as `prompt`")
            ids = list(tok.encode(prompt))
            if len(ids) > MAX_PROMPT_LENGTH:  # as the reference's tokenizer path: truncate, keep the end token
                ids = ids[:MAX_PROMPT_LENGTH - 1] + [49407]
            ids = ids + [49407] * (MAX_PROMPT_LENGTH - len(ids))
            prompt = np.asarray(ids, np.int32)
        tokens = np.asarray(prompt)
        if not np.issubdtype(tokens.dtype, np.integer):
            raise TypeError("encode_text: expected a string, integer token ids or an encoded float array")
        return self._encode_tokens(tokens)

    def _encode_tokens(self, tokens):
        raise NotImplementedError("token ids need the engine-backed StableDiffusion class")

    # ---------------------------------------------------------------------------------- host-side preprocessing
    def gaussian_blur(self, image, radius=3, h_axis=1, v_axis=2):
        """Separable binomial blur with reflected borders (:217-240)."""
        from scipy.ndimage import correlate1d
        if radius == 1:
            taps = np.array([1.0])
        else:
            taps = np.array([1.0, 1.0])
            for _ in range(radius - 2):
                taps = np.convolve(taps, [1.0, 1.0])
        taps = taps / taps.sum()
        out = correlate1d(image, taps, axis=h_axis, mode="reflect", cval=0.0, origin=0)
        return correlate1d(out, taps, axis=v_axis, mode="reflect", cval=0.0, origin=0)

    @staticmethod
    def resize(image_array, new_h=None, new_w=None):
        """Align-corners bilinear resize of an (h,w,c) array (:242-275)."""
        h, w, _c = image_array.shape
        if new_h == h and new_w == w:
            return image_array
        ys = np.linspace(0, h - 1, new_h)[:, None]
        xs = np.linspace(0, w - 1, new_w)[None, :]
        y0 = np.clip(np.floor(ys).astype(int), 0, h - 1)
        y1 = np.clip(np.ceil(ys).astype(int), 0, h - 1)
        x0 = np.clip(np.floor(xs).astype(int), 0, w - 1)
        x1 = np.clip(np.ceil(xs).astype(int), 0, w - 1)
        fy = (ys - y0)[..., None]
        fx = (xs - x0)[..., None]
        top = image_array[y0, x0, :] * (1.0 - fx) + image_array[y0, x1, :] * fx
        bot = image_array[y1, x0, :] * (1.0 - fx) + image_array[y1, x1, :] * fx
        return top * (1.0 - fy) + bot * fy

    def preprocessed_image(self, x):
        """-> ((1,H,W,3) in [0,1], (1,H,W,3) in [-1,1]) (:277-286)."""
        if type(x) is str:
            from PIL import Image
            x = np.array(Image.open(x).convert("RGB"))
        else:
            x = np.array(x)
        arr = np.array(self.resize(x, self.img_height, self.img_width), dtype=np.float32) / 255.0
        arr = arr[None, ..., :3]
        return arr, arr * 2.0 - 1.0

    def preprocessed_mask(self, x, blur_radius=5):
        """-> ((1,H,W,1) pixel mask, (1,h,w,1) latent mask) (:288-302; keeps the reference's (w//8, h//8) argument order)."""
        if type(x) is str:
            from PIL import Image
            x = np.array(Image.open(x).convert("L"))
        else:
            x = np.array(x)
        if x.ndim == 2:
            x = x[..., None]
        m = self.resize(x, self.img_height, self.img_width)
        if m.shape[-1] != 1:
            m = np.mean(m, axis=-1, keepdims=True)
        m = np.array(m, dtype=np.float32) / 255.0
        if blur_radius is not None:
            m = self.gaussian_blur(m, radius=blur_radius, h_axis=0, v_axis=1)
        lat = self.resize(m, self.img_width // 8, self.img_height // 8)
        return m[None], lat[None]

    # (rescale_noise_cfg, :304-315, has no host twin here: it is fused into the step kernel — Engine.cfg_sched_step)

    # ---------------------------------------------------------------------------------- the loop
    def generate_image(self, encoded_text, negative_prompt=None, batch_size=1, num_steps=50, unconditional_guidance_scale=7.5,
                       diffusion_noise=None, seed=None, negative_embedding=None, control_net_image=None, inpaint_mask=None,
                       mask_blur_strength=None, reference_image=None, reference_image_strength=0.8, guidance_rescale=0.0,
                       callback=None, return_latent=False, use_cuda_graph=True, cfg_split=False):
        if diffusion_noise is not None and seed is not None:
            raise ValueError("`diffusion_noise` and `seed` should not both be passed to `generate_image`. `seed` is only "
                             "used to generate diffusion noise when it's not already user-specified.")
        context = self._expand_tensor(encoded_text, batch_size)
        uncond = None
        if unconditional_guidance_scale > 0.0:
            if negative_prompt is None and negative_embedding is None:
                uncond = np.repeat(self._get_unconditional_context(), batch_size, axis=0)
            else:
                uncond = self._expand_tensor(self.encode_text("" if negative_prompt is None else negative_prompt,
                                                              negative_embedding), batch_size)
            # long prompts: both contexts must span the same number of 77-token windows; the shorter one is continued
            # with empty-prompt windows, as the reference pads it (long_prompt_weighting.py)
            tc, tu = context.shape[1], uncond.shape[1]
            if tc != tu:
                if tc % MAX_PROMPT_LENGTH or tu % MAX_PROMPT_LENGTH:
                    raise ValueError(f"context lengths {tc} and {tu} differ and are not multiples of 77")
                empty = np.repeat(self._get_unconditional_context(), batch_size, axis=0)
                if tu < tc:
                    uncond = np.concatenate([uncond] + [empty] * ((tc - tu) // MAX_PROMPT_LENGTH), axis=1)
                else:
                    context = np.concatenate([context] + [empty] * ((tu - tc) // MAX_PROMPT_LENGTH), axis=1)
        if diffusion_noise is not None:
            diffusion_noise = np.squeeze(diffusion_noise)
            if diffusion_noise.ndim == 3:
                diffusion_noise = np.repeat(diffusion_noise[None], batch_size, axis=0)
        self.scheduler.set_timesteps(num_steps)
        timesteps = self.scheduler.timesteps[::-1]
        init_time = init_latent = input_image_array = input_mask_array = latent_mask = None
        if inpaint_mask is not None:
            input_mask_array, latent_mask = self.preprocessed_mask(inpaint_mask, mask_blur_strength)
        if reference_image is not None and (0.0 < reference_image_strength < 1.0):
            input_image_array, input_image_tensor = self.preprocessed_image(reference_image)
            n = int(num_steps * reference_image_strength + 0.5)
            init_time = timesteps[n]
            init_latent = np.asarray(self.image_encoder.predict_on_batch(input_image_tensor))
            timesteps = timesteps[:n]
        noise = diffusion_noise
        if noise is None:
            noise = self._get_initial_diffusion_noise(batch_size, seed)
        noise = np.asarray(noise, dtype=np.float32)
        latent0 = self._get_initial_diffusion_latent(batch_size, init_latent, init_time, noise=noise)
        hint_image = None
        if control_net_image is not None:
            if type(control_net_image) is str:
                from PIL import Image
                img = np.array(Image.open(control_net_image).convert("RGB").resize((self.img_width, self.img_height)))
            elif type(control_net_image) is np.ndarray:
                img = self.resize(control_net_image, self.img_height, self.img_width)
            else:
                print("wrong controlnet image:{}".format(control_net_image))
                img = None
            if img is not None:
                hint_image = np.tile((np.array(img, dtype=np.float32) / 255.0)[None], (batch_size, 1, 1, 1))
        exec_ts = [int(t) for t in timesteps[::-1]]  # descending, as the reference iterates (:442)
        coefs = self.scheduler.coefficients(exec_ts, unconditional_guidance_scale, guidance_rescale)
        t_emb = np.stack([timestep_embedding(t) for t in exec_ts])
        step_noise = None
        if any(c.cn != 0.0 for c in coefs):  # TCD: one global-RNG draw per non-final step (scheduler.py:301)
            step_noise = np.zeros((len(exec_ts),) + latent0.shape, np.float32)
            for i, c in enumerate(coefs):
                if c.cn != 0.0:
                    step_noise[i] = np.random.randn(*latent0.shape).astype(np.float32)
        inpainting = latent_mask is not None and init_latent is not None
        blend = input_mask_array is not None and input_image_array is not None
        out = self.engine.denoise(
            latent0, context, uncond, t_emb, coefs, step_noise=step_noise,
            mask=np.asarray(latent_mask[0, ..., 0], np.float32) if inpainting else None,
            init_latent=np.asarray(init_latent[0], np.float32) if inpainting else None,
            init_noise=noise if inpainting else None, hint_image=hint_image,
            blend_image=input_image_array[0] if blend else None,
            blend_mask=input_mask_array[0, ..., 0] if blend else None,
            decode=True, use_cuda_graph=use_cuda_graph, return_latent=return_latent, cfg_split=cfg_split)
        if callback is not None:
            callback(len(exec_ts))
        return out

    # ---------------------------------------------------------------------------------- helpers (reference names)
    def _get_unconditional_context(self):
        ctx = getattr(self, "unconditional_context", None)
        if ctx is None:
            fn = getattr(self, "text_encoder_fn", None)
            if fn is not None:
                ctx = np.asarray(fn(""), np.float32)
            else:  # the empty prompt: <start> then <end> padding (stable_diffusion.py:533-541, clip_tokenizer.py)
                ctx = self._encode_tokens(np.asarray([[49406] + [49407] * (MAX_PROMPT_LENGTH - 1)], np.int32))
        ctx = np.asarray(ctx, np.float32)
        return ctx[None] if ctx.ndim == 2 else ctx

    def _expand_tensor(self, text_embedding, batch_size):
        text_embedding = np.squeeze(text_embedding)
        if text_embedding.ndim == 2:
            text_embedding = np.repeat(text_embedding[None], batch_size, axis=0)
        return np.asarray(text_embedding, np.float32)

    def _get_timestep_embedding(self, timestep, batch_size, dim=320, max_period=10000):
        return np.repeat(timestep_embedding(timestep, dim, max_period)[None], batch_size, axis=0)

    def _get_initial_diffusion_noise(self, batch_size, seed):
        # the reference draws from TF's stateless Philox via keras.random.normal (:555-557), which is not reproducible
        # without TensorFlow; a seeded NumPy Generator stands in (pass `diffusion_noise` for exact control)
        rng = np.random.default_rng(seed)
        return rng.standard_normal((batch_size, self.img_height // 8, self.img_width // 8, 4)).astype(np.float32)

    def _get_initial_diffusion_latent(self, batch_size, init_latent=None, init_time=None, seed=None, noise=None):
        if noise is None:
            noise = self._get_initial_diffusion_noise(batch_size, seed)
        if init_latent is None:
            return noise
        return (self.scheduler.signal_rates[init_time] * np.repeat(init_latent, batch_size, axis=0)
                + self.scheduler.noise_rates[init_time] * noise).astype(np.float32)

    @staticmethod
    def _get_pos_ids():
        return np.asarray([list(range(MAX_PROMPT_LENGTH))], dtype=np.int32)


class StableDiffusion(StableDiffusionBase):
    """Reference constructor (stable_diffusion.py:620-631).  `*_ckpt` may be a path (.safetensors / torch pickle in
    the reference's key convention) or an in-memory state dict.  Without real checkpoints offline, `synthetic=True`
    builds seeded random-init SD1.5 weights (minsdtf_b200.synth)."""

    def __init__(self, img_height=512, img_width=512, jit_compile=False, clip_skip=-1, unet_ckpt=None,
                 text_encoder_ckpt=None, vae_ckpt=None, lora_path=None, controlnet_path=None, active_tcd=False,
                 device=0, synthetic=False, engine=None):
        super().__init__(img_height, img_width, jit_compile, active_tcd)
        self.clip_skip = clip_skip
        self.unet_ckpt = unet_ckpt
        self.text_encoder_ckpt = text_encoder_ckpt
        self.vae_ckpt = vae_ckpt
        self.controlnet_path = controlnet_path
        self.lora_path = None
        if lora_path is not None:
            raise NotImplementedError("LoRA merging is load-time host math outside this round's scope (SURVEY.md §2)")
        self.synthetic = synthetic
        self.device = device
        self._engine = engine
        self._models = {}

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self.device)
        return self._engine

    def _ensure(self, component, src):
        eng = self.engine
        if component in eng.loaded:
            return
        if src is None:
            if not self.synthetic:
                raise FileNotFoundError(f"no checkpoint given for {component} and downloads are not possible offline; pass a "
                                        f"path / state dict or synthetic=True")
            from . import synth
            src = {"unet": lambda: synth.make_state_dict("unet"), "vae_decoder": lambda: synth.make_state_dict("decoder"),
                   "vae_encoder": lambda: synth.make_state_dict("encoder"),
                   "text_encoder": lambda: synth.make_state_dict("text_encoder"),
                   "controlnet": synth.make_controlnet_state_dict}[component]()
        if isinstance(src, (str, os.PathLike)):
            eng.load_file(str(src), component)
        else:
            eng.load_state_dict(src, component)

    def load_all(self, control=False, encoder=False):
        self._ensure("unet", self.unet_ckpt)
        self._ensure("vae_decoder", self.vae_ckpt)
        if encoder:
            self._ensure("vae_encoder", self.vae_ckpt)
        if control:
            self._ensure("controlnet", self.controlnet_path)

    # the seven overridable model properties of the reference (:505-531, 650-725)
    @property
    def diffusion_model(self):
        self._ensure("unet", self.unet_ckpt)
        return _Model(lambda x: self.engine.unet(x[0], x[1], x[2], list(x[3:]) if len(x) > 3 else None))

    @property
    def image_decoder(self):
        self._ensure("vae_decoder", self.vae_ckpt)
        return _Model(lambda x: self.engine.vae_decode(x))

    @property
    def image_encoder(self):
        self._ensure("vae_encoder", self.vae_ckpt)
        return _Model(lambda x: self.engine.vae_encode(x))

    @property
    def hint_net(self):
        self._ensure("controlnet", self.controlnet_path)
        return _Model(lambda x: self.engine.hintnet(x))

    @property
    def control_net(self):
        self._ensure("controlnet", self.controlnet_path)
        return _Model(lambda x: self.engine.controlnet(x[0], x[1], x[2], x[3]))

    # The reference splits the text tower into two Keras models, TextClipEmbedding([tokens, positions]) -> (B,77,768)
    # and TextEncoder(embedding) (:700-725); the engine runs embedding lookup and the 12 layers in one call, so the
    # first shim only carries the token ids to the second (the pair composes exactly like the reference's:
    # `text_encoder.predict_on_batch(text_clip_embedding.predict_on_batch([tokens, positions]))`).
    @property
    def text_clip_embedding(self):
        self._ensure("text_encoder", self.text_encoder_ckpt)
        return _Model(lambda x: np.asarray(x[0], np.int32))

    @property
    def text_encoder(self):
        self._ensure("text_encoder", self.text_encoder_ckpt)
        return _Model(lambda tokens: self.engine.text_encode(tokens, self.clip_skip))

    def _encode_tokens(self, tokens):
        """(T,) / (B,T) token ids -> (T,768) / (B,T,768).  T > 77 must be a multiple of 77: each 77-token window is
        encoded on its own and the contexts are concatenated, which is how the reference feeds long prompts to the
        UNet (long_prompt_weighting.py: chunks of <start> + 75 tokens + <end>), giving contexts of 154, 231, ... tokens."""
        self._ensure("text_encoder", self.text_encoder_ckpt)
        tokens = np.asarray(tokens, np.int32)
        one = tokens.ndim == 1
        tok2 = tokens[None] if one else tokens
        B, T = tok2.shape
        if T > MAX_PROMPT_LENGTH:
            if T % MAX_PROMPT_LENGTH:
                raise ValueError(f"token ids: length {T} is neither <= 77 nor a multiple of 77")
            m = T // MAX_PROMPT_LENGTH
            out = self.engine.text_encode(tok2.reshape(B * m, MAX_PROMPT_LENGTH), self.clip_skip).reshape(B, T, -1)
        else:
            out = self.engine.text_encode(tok2, self.clip_skip)
        return out[0] if one else out

    def generate_image(self, encoded_text, **kw):
        self._ensure("unet", self.unet_ckpt)
        self._ensure("vae_decoder", self.vae_ckpt)
        if kw.get("control_net_image") is not None:
            self._ensure("controlnet", self.controlnet_path)
        return super().generate_image(encoded_text, **kw)
