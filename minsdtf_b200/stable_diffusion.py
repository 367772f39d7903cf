"""`StableDiffusion` with the reference's Python surface (stable_diffusion/stable_diffusion.py:47-725) on top of the
B200 engine.

Kept verbatim from the reference: constructor kwargs (:620-631), `text_to_image / image_to_image / inpaint /
generate_image` signatures and defaults (:84-174, 317-334), the seven model properties reached only through
`.predict_on_batch` (:505-531), `scheduler`, the host-side image / mask preprocessing (:217-302) and the quirks of the
loop (img2img timestep slicing :406-416, inpaint re-noising at the *current* t :469-475, uint8 truncation :486).

Different by design: the loop body (2 UNet calls (+2 ControlNet calls) + CFG + scheduler step per iteration, :442-475)
and the decode (:482-486) run on the GPU inside one `sdtf_denoise` call; cond and uncond are batched into one UNet
pass; cross-attention K/V and HintNet features are computed once per image.

Text side: the CLIP text tower runs on the engine (`text_clip_embedding` / `text_encoder` models, as the
reference splits it); string prompts go through `prompt_weighting.weighted_text_embeddings` — `(word:1.3)` attention
syntax, up to four 75-token windows, mean-preserving re-weighting, textual-inversion vectors — with a CLIP BPE tokenizer
built from a local vocabulary file (the reference downloads it; `bpe_vocab=` / $SDTF_BPE_VOCAB / the Keras cache path).
LoRA files are merged into the checkpoint tensors at load time (`lora.py`).
"""
from __future__ import annotations

import os

import numpy as np

from . import lora as _lora, prompt_weighting as _pw
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

    def load_embedding(self, embedding_path):
        """Textual-inversion file -> (n_vectors, 768) array, or None (stable_diffusion.py:71-82): a torch pickle whose
        `string_to_param` dict holds the learned vectors (the last float entry wins, as in the reference)."""
        embedding = None
        if os.path.exists(str(embedding_path)):
            import torch
            state = torch.load(embedding_path, map_location="cpu")
            params = state.get("string_to_param") if hasattr(state, "get") else None
            if params is not None:
                for value in params.values():
                    if value.dtype in (torch.float32, torch.float16):
                        embedding = value.detach().float().numpy()
        return embedding

    def encode_text(self, prompt, embedding_data=None):
        """Prompt -> context (stable_diffusion.py:176-215).
        * a string (or list of strings): tokenised by `self.tokenizer`, attention syntax `(word:1.3)` / `[word]` parsed,
          up to 4 windows of 75 tokens each encoded on its own -> (B, 77 m, 768), weighted with the mean preserved
          (long_prompt_weighting.py:240-333);
        * `embedding_data`: path of a textual-inversion file (or, beyond the reference, an (n,768) array): n placeholder
          tokens are put in front of the prompt and their embedding rows replaced by the learned vectors (:202-209);
        * beyond the reference: a float array (T,768) / (B,T,768) is passed through as an already-encoded prompt, an
          integer array (T,) / (B,T<=77 or 77 m) is taken as CLIP token ids."""
        if isinstance(prompt, np.ndarray) and np.issubdtype(prompt.dtype, np.floating):
            return prompt.astype(np.float32)
        if isinstance(prompt, str) or (isinstance(prompt, (list, tuple)) and prompt and isinstance(prompt[0], str)):
            embedding, count = None, 0
            if embedding_data is not None:
                if isinstance(embedding_data, str):
                    embedding = self.load_embedding(embedding_data)
                    if embedding is None:
                        raise ValueError(f"failed to load embedding file: {embedding_data}.")
                elif isinstance(embedding_data, np.ndarray):
                    embedding = np.asarray(embedding_data, np.float32)
                if embedding is not None:
                    count = embedding.shape[0]
                    embedding = embedding[None]
            fn = getattr(self, "text_encoder_fn", None)
            if fn is not None and getattr(self, "_tokenizer", None) is None and embedding is None:
                return np.asarray(fn(prompt), dtype=np.float32)  # user-supplied text encoder
            return _pw.weighted_text_embeddings(self.tokenizer, self.text_clip_embedding.predict_on_batch,
                                                self.text_encoder.predict_on_batch, prompt, pad_token_id=49407,
                                                embedding=embedding, embedding_tokens_count=count)
        tokens = np.asarray(prompt)
        if not np.issubdtype(tokens.dtype, np.integer):
            raise TypeError("encode_text: expected a string, integer token ids or an encoded float array")
        return self._encode_tokens(tokens)

    def _encode_tokens(self, tokens):
        raise NotImplementedError("token ids need the engine-backed StableDiffusion class")

    @property
    def tokenizer(self):
        """CLIP BPE tokenizer (the reference's property, stable_diffusion.py:522-531, builds SimpleTokenizer() from a
        downloaded vocabulary).  Here the vocabulary file must be local: `bpe_vocab=` of the constructor,
        $SDTF_BPE_VOCAB, or where Keras would have cached the download."""
        if getattr(self, "_tokenizer", None) is None:
            from .bpe import ClipBPE
            cands = [getattr(self, "bpe_vocab", None), os.environ.get("SDTF_BPE_VOCAB"),
                     os.path.expanduser("~/.keras/datasets/bpe_simple_vocab_16e6.txt.gz")]
            path = next((c for c in cands if c and os.path.exists(str(c))), None)
            if path is None:
                raise FileNotFoundError(
                    "no CLIP BPE vocabulary: the reference downloads bpe_simple_vocab_16e6.txt.gz (clip_tokenizer.py:79-82), "
                    "which is impossible offline — pass bpe_vocab=<path>, set $SDTF_BPE_VOCAB, assign `model.tokenizer`, or "
                    "pass token ids / an encoded (T,768) array as `prompt`")
            self._tokenizer = ClipBPE(str(path))
        return self._tokenizer

    @tokenizer.setter
    def tokenizer(self, tok):
        self._tokenizer = tok

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
        if cfg_split and diffusion_noise is None and seed is None:
            raise ValueError("cfg_split: both GPUs of a pair must start from the same latent — pass `seed` or `diffusion_noise`")
        context = self._expand_tensor(encoded_text, batch_size)
        uncond = None
        if unconditional_guidance_scale > 0.0:
            if negative_prompt is None and negative_embedding is None:
                uncond = np.repeat(self._get_unconditional_context(), batch_size, axis=0)
            else:
                uncond = self._expand_tensor(self.encode_text("" if negative_prompt is None else negative_prompt,
                                                              negative_embedding), batch_size)
            # cond and uncond may span different numbers of 77-token windows (a long prompt against the empty negative
            # prompt): the reference evaluates the two branches as separate model calls (:454-457), and so does the
            # engine then (two B-sample passes per step instead of one 2B pass)
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
            # with the CFG split both members of a pair must draw the same noise: a generator seeded by `seed` replaces
            # the process-global one there
            draw = np.random.RandomState(0 if seed is None else seed).randn if cfg_split else np.random.randn
            for i, c in enumerate(coefs):
                if c.cn != 0.0:
                    step_noise[i] = draw(*latent0.shape).astype(np.float32)
        inpainting = latent_mask is not None and init_latent is not None
        blend = input_mask_array is not None and input_image_array is not None
        out = self.engine.denoise(
            latent0, context, uncond, t_emb, coefs, step_noise=step_noise,
            mask=np.asarray(latent_mask[0, ..., 0], np.float32) if inpainting else None,
            init_latent=np.asarray(init_latent[0], np.float32) if inpainting else None,
            init_noise=noise if inpainting else None, hint_image=hint_image,
            blend_image=input_image_array[0] if blend else None,
            blend_mask=input_mask_array[0, ..., 0] if blend else None,
            decode=True, use_cuda_graph=use_cuda_graph, return_latent=return_latent, cfg_split=cfg_split, callback=callback)
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
    builds seeded random-init SD1.5 weights (minsdtf_b200.synth).  `lora_path`: kohya-format LoRA file (or its state
    dict), merged into the UNet / text-encoder tensors before they are packed (ckpt_loader.py:2169-2180, 2196-2276)."""

    def __init__(self, img_height=512, img_width=512, jit_compile=False, clip_skip=-1, unet_ckpt=None,
                 text_encoder_ckpt=None, vae_ckpt=None, lora_path=None, controlnet_path=None, active_tcd=False,
                 device=0, synthetic=False, engine=None, bpe_vocab=None):
        super().__init__(img_height, img_width, jit_compile, active_tcd)
        self.clip_skip = clip_skip
        self.unet_ckpt = unet_ckpt
        self.text_encoder_ckpt = text_encoder_ckpt
        self.vae_ckpt = vae_ckpt
        self.controlnet_path = controlnet_path
        self.lora_path = None
        self.text_encoder_lora_dict = None
        self.unet_lora_dict = None
        if isinstance(lora_path, dict) or os.path.exists(str(lora_path)):  # a missing file is ignored, as the reference does (:642)
            self.text_encoder_lora_dict, self.unet_lora_dict = _lora.load_lora(lora_path)
            self.lora_path = lora_path
        self.synthetic = synthetic
        self.device = device
        self.bpe_vocab = bpe_vocab
        self._engine = engine
        self._models = {}
        self._loaded_here = set()
        self.scheduler.device = device

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self.device)
        self.scheduler.engine = self._engine  # Scheduler.step (the reference's per-call API) runs on the same engine
        return self._engine

    def _ensure(self, component, src):
        eng = self.engine
        deltas = {"unet": self.unet_lora_dict, "text_encoder": self.text_encoder_lora_dict}.get(component)
        if component in eng.loaded and (not deltas or component in self._loaded_here):
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
            src = eng.read_checkpoint(str(src))
        if deltas:
            from . import keys as K
            src, n, unused = _lora.merge(src, deltas, K.unet_alias_map() if component == "unet" else None)
            print(f"Apply {n}/{len(deltas)} lora weights" if unused else f"Apply {n} lora weights")
            if unused:
                print(f"Failed to apply lora list: {unused}")
        eng.load_state_dict(src, component)
        self._loaded_here.add(component)

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

    # The reference's two text models (:674-690, 700-725): TextClipEmbedding([tokens, positions]) -> (B,77,768) embeddings
    # and TextEncoder(embeddings) -> context.  Both are engine entry points (sdtf_text_embed, sdtf_text_encode_embedded),
    # so textual-inversion vectors can be spliced in between exactly as long_prompt_weighting.py:202-209 does.
    @property
    def text_clip_embedding(self):
        self._ensure("text_encoder", self.text_encoder_ckpt)
        return _Model(lambda x: self.engine.text_embed(x[0], x[1] if len(x) > 1 else None))

    @property
    def text_encoder(self):
        self._ensure("text_encoder", self.text_encoder_ckpt)
        return _Model(lambda emb: self.engine.text_encode(emb, self.clip_skip))

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

    def generate_image(self, encoded_text, negative_prompt=None, batch_size=1, num_steps=50, unconditional_guidance_scale=7.5,
                       diffusion_noise=None, seed=None, negative_embedding=None, control_net_image=None, inpaint_mask=None,
                       mask_blur_strength=None, reference_image=None, reference_image_strength=0.8, guidance_rescale=0.0,
                       callback=None, **engine_options):
        """Same positional signature as the reference (stable_diffusion.py:317-334); `engine_options`: return_latent,
        use_cuda_graph, cfg_split."""
        self._ensure("unet", self.unet_ckpt)
        self._ensure("vae_decoder", self.vae_ckpt)
        if control_net_image is not None:
            self._ensure("controlnet", self.controlnet_path)
        return super().generate_image(
            encoded_text, negative_prompt=negative_prompt, batch_size=batch_size, num_steps=num_steps,
            unconditional_guidance_scale=unconditional_guidance_scale, diffusion_noise=diffusion_noise, seed=seed,
            negative_embedding=negative_embedding, control_net_image=control_net_image, inpaint_mask=inpaint_mask,
            mask_blur_strength=mask_blur_strength, reference_image=reference_image,
            reference_image_strength=reference_image_strength, guidance_rescale=guidance_rescale, callback=callback,
            **engine_options)
