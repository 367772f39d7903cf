"""TEST INFRASTRUCTURE — CPU oracle, never imported by the product path.

torch-fp32 CPU restatement of the reference's Keras graphs for the SD1.5 denoising hot path.  Each function cites
the reference lines it restates.  Tensors cross this module's API as NumPy NHWC fp32, exactly like the
reference's `predict_on_batch`; weights are a dict {checkpoint key: torch tensor in PyTorch layout} keyed by the
reference's names (ckpt_loader.py CKPT_MAPPING), i.e. the same file the reference would load.

PARITY UNPINNED for the graphs: the arithmetic of the reference lives in `keras` (+ TensorFlow backend), which
is neither vendored in /root/reference nor installable offline, and the reference ships no tests or golden
vectors.  What IS pinned: the scheduler (oracle/scheduler_oracle.py vs the imported
scheduler.py), the weight-name tables, and the parameter/tensor counts of every graph built here.
Keras-3 semantics relied upon: GroupNormalization(groups=32, axis=-1) biased variance, eps inside rsqrt;
LayerNormalization(axis=-1); Dense kernel (in,out) == torch (out,in)^T; Conv2D HWIO 'valid' after explicit
ZeroPadding2D; UpSampling2D(2) nearest; swish = x*sigmoid(x); softmax over the last axis.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .scheduler_oracle import OracleScheduler, cfg_combine, timestep_embedding

# Optional emulation of the engine's storage precision: when set to torch.bfloat16 every activation that the
# engine materialises in HBM is rounded to bf16 (used only to budget the bf16-vs-fp32 error).
_ROUND = None


def set_round_dtype(dt):
    global _ROUND
    _ROUND = dt


_ROUND_STREAM = True  # False: keep the residual stream (block outputs) unrounded (fp32-stream engine variant)


def set_round_stream(flag):
    global _ROUND_STREAM
    _ROUND_STREAM = flag


def _r(x):
    return x if _ROUND is None else x.to(_ROUND).to(torch.float32)


def _rs(x):
    """rounding applied to residual-stream tensors (outputs of `branch + x` sums)."""
    return _r(x) if _ROUND_STREAM else x


def _w(sd, key):
    w = sd[key]
    if not isinstance(w, torch.Tensor):
        w = torch.as_tensor(np.asarray(w))
    w = w.to(torch.float32)
    return _r(w) if (_ROUND is not None and w.ndim >= 2) else w


def _nchw(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).permute(0, 3, 1, 2).contiguous()


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().numpy()


def conv(sd, key, x, stride=1, pad=(1, 1, 1, 1)):
    """layers.py:17-25 PaddedConv2D: ZeroPadding2D then VALID Conv2D.  pad = (left, right, top, bottom)."""
    w = _w(sd, key + ".weight")
    if w.shape[-1] == 1:
        pad = (0, 0, 0, 0)
    x = F.pad(x, pad)
    return _r(F.conv2d(x, w, _w(sd, key + ".bias"), stride=stride))


def group_norm(sd, key, x, silu):
    """keras GroupNormalization(groups=32 default, epsilon=1e-5) (+ Activation('swish'))."""
    y = F.group_norm(x, 32, _w(sd, key + ".weight"), _w(sd, key + ".bias"), eps=1e-5)
    if silu:
        y = F.silu(y)
    return _r(y)


def linear(sd, key, x, bias=True):
    return F.linear(x, _w(sd, key + ".weight"), _w(sd, key + ".bias") if bias else None)


def res_block(sd, p, x, temb):
    """diffusion_model.py:22-51 ResBlock (LDM key names)."""
    h = group_norm(sd, f"{p}.in_layers.0", x, True)
    h = F.conv2d(F.pad(h, (1, 1, 1, 1)), _w(sd, f"{p}.in_layers.2.weight"), _w(sd, f"{p}.in_layers.2.bias"))
    e = linear(sd, f"{p}.emb_layers.1", temb)                        # :30,47
    h = _r(h + e[:, :, None, None])                                  # :48
    h = group_norm(sd, f"{p}.out_layers.0", h, True)
    h = F.conv2d(F.pad(h, (1, 1, 1, 1)), _w(sd, f"{p}.out_layers.3.weight"), _w(sd, f"{p}.out_layers.3.bias"))
    if f"{p}.skip_connection.weight" in sd:
        x = F.conv2d(x, _w(sd, f"{p}.skip_connection.weight"), _w(sd, f"{p}.skip_connection.bias"))  # :36-38
    return _rs(h + x)                                                # :51


def cross_attention(sd, p, x, context, heads):
    """diffusion_model.py:99-129; x (B,N,C), context (B,T,Cc) or None for self-attention."""
    ctx = x if context is None else context
    q = _r(linear(sd, f"{p}.to_q", x, bias=False))
    k = _r(linear(sd, f"{p}.to_k", ctx, bias=False))
    v = _r(linear(sd, f"{p}.to_v", ctx, bias=False))
    B, N, C = q.shape
    d = C // heads
    q = q.view(B, N, heads, d).permute(0, 2, 1, 3)
    k = k.view(B, -1, heads, d).permute(0, 2, 3, 1)
    v = v.view(B, -1, heads, d).permute(0, 2, 1, 3)
    score = torch.matmul(q, k) * (d ** -0.5)                         # :123
    w = torch.softmax(score, dim=-1)                                 # :124
    o = torch.matmul(w, v).permute(0, 2, 1, 3).reshape(B, N, C)     # :126-128
    o = _r(o)
    return linear(sd, f"{p}.to_out.0", o)


def geglu(sd, p, x):
    """diffusion_model.py:142-153 (tanh-approximate GELU on the gate half)."""
    h = linear(sd, f"{p}.proj", x)
    half = h.shape[-1] // 2
    val, gate = h[..., :half], h[..., half:]
    t = torch.tanh(gate * 0.7978845608 * (1 + 0.044715 * gate ** 2))
    return _r(val * 0.5 * gate * (1 + t))


def transformer_block(sd, p, x, context, heads):
    """diffusion_model.py:81-96."""
    C = x.shape[-1]

    def ln(name, t):
        return _r(F.layer_norm(t, (C,), _w(sd, f"{p}.{name}.weight"), _w(sd, f"{p}.{name}.bias"), eps=1e-5))

    x = _rs(cross_attention(sd, f"{p}.attn1", ln("norm1", x), None, heads) + x)
    x = _rs(cross_attention(sd, f"{p}.attn2", ln("norm2", x), context, heads) + x)
    x = _rs(linear(sd, f"{p}.ff.net.2", geglu(sd, f"{p}.ff.net.0", ln("norm3", x))) + x)
    return x


def attentions(sd, p, x, context, heads):
    """diffusion_model.py:54-78 Attentions (spatial transformer with conv 1x1 projections)."""
    B, C, H, W = x.shape
    h = group_norm(sd, f"{p}.norm", x, False)
    h = _r(F.conv2d(h, _w(sd, f"{p}.proj_in.weight"), _w(sd, f"{p}.proj_in.bias")))
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)                   # :75 row-major (y, x) tokens
    h = transformer_block(sd, f"{p}.transformer_blocks.0", h, context, heads)
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return _rs(F.conv2d(h, _w(sd, f"{p}.proj_out.weight"), _w(sd, f"{p}.proj_out.bias")) + x)


def time_mlp(sd, p, t_emb):
    """diffusion_model.py:184-188: Dense -> swish -> Dense(activation=swish)."""
    t = F.silu(linear(sd, f"{p}.time_embed.0", t_emb))
    return F.silu(linear(sd, f"{p}.time_embed.2", t))


_ENC = [(1, True), (2, True), (4, True), (5, True), (7, True), (8, True), (10, False), (11, False)]


def _encoder_half(sd, p, x, temb, ctx):
    """input_blocks.1..11 + middle_block shared by the UNet (diffusion_model.py:193-229) and ControlNet
    (control_net.py:58-89).  x = conv_in output (and hint, for ControlNet).  Returns (x_mid, 12 skips)."""
    outs = [x]
    has_attn = dict(_ENC)
    for i in range(1, 12):
        if i in (3, 6, 9):
            x = _r(F.conv2d(F.pad(x, (1, 1, 1, 1)), _w(sd, f"{p}.input_blocks.{i}.0.op.weight"),
                            _w(sd, f"{p}.input_blocks.{i}.0.op.bias"), stride=2))       # :200
        else:
            x = res_block(sd, f"{p}.input_blocks.{i}.0", x, temb)
            if has_attn[i]:
                x = attentions(sd, f"{p}.input_blocks.{i}.1", x, ctx, 8)
        outs.append(x)
    x = res_block(sd, f"{p}.middle_block.0", x, temb)
    x = attentions(sd, f"{p}.middle_block.1", x, ctx, 8)
    x = res_block(sd, f"{p}.middle_block.2", x, temb)
    return x, outs


_DEC = [(0, False, False), (1, False, False), (2, False, True), (3, True, False), (4, True, False), (5, True, True),
        (6, True, False), (7, True, False), (8, True, True), (9, True, False), (10, True, False), (11, True, False)]


def unet_forward(sd, latent, t_emb, context, controls=None):
    """DiffusionModel forward, diffusion_model.py:163-283.  latent (B,h,w,4), t_emb (B,320), context (B,T,768),
    controls: optional list of 13 NHWC arrays.  Returns eps (B,h,w,4) NHWC fp32."""
    p = "model.diffusion_model"
    with torch.no_grad():
        x = _r(_nchw(latent))
        ctx = _r(torch.as_tensor(np.asarray(context), dtype=torch.float32))
        temb = time_mlp(sd, p, torch.as_tensor(np.asarray(t_emb), dtype=torch.float32))
        x = _r(F.conv2d(F.pad(x, (1, 1, 1, 1)), _w(sd, f"{p}.input_blocks.0.0.weight"),
                        _w(sd, f"{p}.input_blocks.0.0.bias")))                               # :191
        x, outs = _encoder_half(sd, p, x, temb, ctx)
        if controls is not None:                                                            # :230-234
            x = _r(x + _nchw(controls[12]))
            outs = [_r(o + _nchw(c)) for o, c in zip(outs, controls[:12])]
        for i, at, up in _DEC:
            x = torch.cat([x, outs.pop()], dim=1)                                           # :237 [x, skip]
            x = res_block(sd, f"{p}.output_blocks.{i}.0", x, temb)
            j = 1
            if at:
                x = attentions(sd, f"{p}.output_blocks.{i}.1", x, ctx, 8)
                j = 2
            if up:                                                                          # :132-139
                x = F.interpolate(x, scale_factor=2, mode="nearest")
                x = _r(F.conv2d(F.pad(x, (1, 1, 1, 1)), _w(sd, f"{p}.output_blocks.{i}.{j}.conv.weight"),
                                _w(sd, f"{p}.output_blocks.{i}.{j}.conv.bias")))
        x = group_norm(sd, f"{p}.out.0", x, True)                                           # :277-278
        x = F.conv2d(F.pad(x, (1, 1, 1, 1)), _w(sd, f"{p}.out.2.weight"), _w(sd, f"{p}.out.2.bias"))
        return _nhwc(x)


def hintnet_forward(sd, image):
    """control_net.py:10-31; image (B,H,W,3) in [0,1] -> (B,H/8,W/8,320)."""
    strides = [1, 1, 2, 1, 2, 1, 2, 1]
    with torch.no_grad():
        x = _r(_nchw(image))
        for i, s in enumerate(strides):
            k = f"control_model.input_hint_block.{2 * i}"
            x = F.conv2d(F.pad(x, (1, 1, 1, 1)), _w(sd, k + ".weight"), _w(sd, k + ".bias"), stride=s)
            if i < 7:
                x = F.silu(x)
            x = _r(x)
        return _nhwc(x)


def controlnet_forward(sd, latent, t_emb, context, hint):
    """control_net.py:45-107 -> list of 13 NHWC residuals."""
    p = "control_model"
    with torch.no_grad():
        x = _r(_nchw(latent))
        ctx = _r(torch.as_tensor(np.asarray(context), dtype=torch.float32))
        temb = time_mlp(sd, p, torch.as_tensor(np.asarray(t_emb), dtype=torch.float32))
        x = F.conv2d(F.pad(x, (1, 1, 1, 1)), _w(sd, f"{p}.input_blocks.0.0.weight"), _w(sd, f"{p}.input_blocks.0.0.bias"))
        x = _r(x + _nchw(hint))                                                             # :56
        xm, outs = _encoder_half(sd, p, x, temb, ctx)
        outs.append(xm)
        res = []
        for i, o in enumerate(outs):
            k = f"{p}.zero_convs.{i}.0" if i < 12 else f"{p}.middle_block_out.0"
            res.append(_nhwc(F.conv2d(o, _w(sd, k + ".weight"), _w(sd, k + ".bias"))))      # :92-106
        return res


def _vae_res(sd, p, x):
    """layers.py:62-80 ResnetBlock."""
    h = conv(sd, f"{p}.conv1", group_norm(sd, f"{p}.norm1", x, True))
    h = conv(sd, f"{p}.conv2", group_norm(sd, f"{p}.norm2", h, True))
    if f"{p}.conv_shortcut.weight" in sd:
        x = F.conv2d(x, _w(sd, f"{p}.conv_shortcut.weight"), _w(sd, f"{p}.conv_shortcut.bias"))
    return _r(h + x)


def _vae_attn(sd, p, x):
    """layers.py:28-59 AttentionBlock: single head, scale 1/sqrt(C)."""
    B, C, H, W = x.shape
    h = group_norm(sd, f"{p}.group_norm", x, False).permute(0, 2, 3, 1).reshape(B, H * W, C)
    q = _r(linear(sd, f"{p}.query", h))
    k = _r(linear(sd, f"{p}.key", h))
    v = _r(linear(sd, f"{p}.value", h))
    s = torch.matmul(q, k.transpose(1, 2)) * (1.0 / math.sqrt(C))                           # :48-49
    w = torch.softmax(s, dim=-1)
    o = _r(torch.matmul(w, v))
    o = linear(sd, f"{p}.proj_attn", o).reshape(B, H, W, C).permute(0, 3, 1, 2)
    return _r(o + x)


def vae_decode(sd, latent):
    """image_decoder.py:22-55; latent (B,h,w,4) -> (B,8h,8w,3) in ~[-1,1]."""
    with torch.no_grad():
        x = _nchw(latent) * (1.0 / 0.18215)                                                 # :27
        x = _r(x)
        x = conv(sd, "post_quant_conv", x)
        x = conv(sd, "decoder.conv_in", x)
        x = _vae_res(sd, "decoder.mid_block.resnets.0", x)
        x = _vae_attn(sd, "decoder.mid_block.attentions.0", x)
        x = _vae_res(sd, "decoder.mid_block.resnets.1", x)
        for b in range(4):
            for r in range(3):
                x = _vae_res(sd, f"decoder.up_blocks.{b}.resnets.{r}", x)
            if b < 3:
                x = F.interpolate(x, scale_factor=2, mode="nearest")
                x = conv(sd, f"decoder.up_blocks.{b}.upsamplers.0.conv", x)
        x = group_norm(sd, "decoder.conv_norm_out", x, True)
        x = F.conv2d(F.pad(x, (1, 1, 1, 1)), _w(sd, "decoder.conv_out.weight"), _w(sd, "decoder.conv_out.bias"))
        return _nhwc(x)


def vae_encode(sd, image):
    """image_encoder.py:21-48; image (B,H,W,3) in [-1,1] -> latent mean * 0.18215, (B,H/8,W/8,4)."""
    with torch.no_grad():
        x = _r(_nchw(image))
        x = conv(sd, "encoder.conv_in", x)
        for b in range(4):
            for r in range(2):
                x = _vae_res(sd, f"encoder.down_blocks.{b}.resnets.{r}", x)
            if b < 3:
                x = conv(sd, f"encoder.down_blocks.{b}.downsamplers.0.conv", x, stride=2, pad=(0, 1, 0, 1))  # :31
        x = _vae_res(sd, "encoder.mid_block.resnets.0", x)
        x = _vae_attn(sd, "encoder.mid_block.attentions.0", x)
        x = _vae_res(sd, "encoder.mid_block.resnets.1", x)
        x = group_norm(sd, "encoder.conv_norm_out", x, True)
        x = F.conv2d(F.pad(x, (1, 1, 1, 1)), _w(sd, "encoder.conv_out.weight"), _w(sd, "encoder.conv_out.bias"))
        x = F.conv2d(x, _w(sd, "quant_conv.weight"), _w(sd, "quant_conv.bias"))
        return _nhwc(x[:, :4] * 0.18215)                                                    # :47


def to_uint8(decoded, input_image_array=None, input_mask_array=None):
    """stable_diffusion.py:483-486 (note: truncation, not rounding)."""
    d = np.array((decoded + 1.0) * 0.5, dtype=np.float32)
    if input_mask_array is not None and input_image_array is not None:
        d = input_image_array * (1.0 - input_mask_array) + d * input_mask_array
    return np.clip(d * 255.0, 0, 255).astype("uint8")


def generate_image(weights, context, unconditional_context, diffusion_noise, num_steps=25, guidance_scale=7.5,
                   guidance_rescale=0.0, active_tcd=False, init_latent=None, strength=0.8, latent_mask=None,
                   hint=None, input_image_array=None, input_mask_array=None, trace=None, unet_fn=None,
                   decode=True):
    """The loop of stable_diffusion.py:384-486 with every model call going to this oracle.

    weights: {"unet": sd, "vae": sd, "controlnet": sd}.  context/unconditional_context (B,T,768).
    diffusion_noise (B,h,w,4).  init_latent (1,h,w,4): img2img / inpaint start (already VAE-encoded).
    trace: optional dict collecting per-step tensors (teacher-forcing material for the parity tests).
    """
    sched = OracleScheduler(active_tcd=active_tcd)
    sched.set_timesteps(num_steps)
    timesteps = sched.timesteps[::-1]
    B = diffusion_noise.shape[0]
    noise = diffusion_noise

    def noised(t):
        return sched.signal_rates[t] * np.repeat(init_latent, B, axis=0) + sched.noise_rates[t] * noise  # :566-567

    if init_latent is not None:
        n = int(num_steps * strength + 0.5)
        init_time = timesteps[n]
        timesteps = timesteps[:n]
        latent = noised(init_time)
    else:
        latent = noise
    unet = unet_fn or (lambda lat, te, ctx, ctrl: unet_forward(weights["unet"], lat, te, ctx, ctrl))
    for index, t in list(enumerate(timesteps))[::-1]:
        latent_prev = latent
        t_emb = timestep_embedding(t, B)

        def call(ctx):
            ctrl = None
            if hint is not None:
                ctrl = controlnet_forward(weights["controlnet"], np.asarray(latent, np.float32), t_emb, ctx, hint)
            return unet(np.asarray(latent, np.float32), t_emb, ctx, ctrl)

        if guidance_scale > 0.0:
            eps_u = call(unconditional_context)
            eps_c = call(context)
            eps = cfg_combine(eps_u, eps_c, guidance_scale, guidance_rescale)
        else:
            eps_u = None
            eps_c = call(context)
            eps = eps_c
        latent = sched.step(eps, t, latent_prev)
        if latent_mask is not None and init_latent is not None:
            latent = noised(t) * (1.0 - latent_mask) + latent * latent_mask                 # :469-475
        if trace is not None:
            trace.setdefault("t", []).append(int(t))
            trace.setdefault("latent_in", []).append(np.asarray(latent_prev, np.float32))
            trace.setdefault("eps_u", []).append(eps_u)
            trace.setdefault("eps_c", []).append(eps_c)
            trace.setdefault("latent_out", []).append(np.asarray(latent))
    if not decode:
        return latent
    decoded = vae_decode(weights["vae"], np.asarray(latent, np.float32))
    return to_uint8(decoded, input_image_array, input_mask_array)
