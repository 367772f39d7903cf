"""GPU unit tests of the individual kernels, each against a plain torch fp32 reference of the same op (fed the same
bf16-rounded inputs, so the tolerance only has to cover fp32-accumulate + one bf16 output rounding)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from minsdtf_b200 import _lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bf(x):
    return torch.as_tensor(x).to(torch.bfloat16).to(torch.float32)


def test_conv_gemm_standalone_cases():
    """tests/cuda/test_gemm.cu: 18 conv / linear / GEGLU / concat / stride-2 cases vs a naive CUDA reference."""
    from minsdtf_b200 import build
    exe = build.build_test_gemm()
    r = subprocess.run(["timeout", "300", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "ALL CASES PASSED" in r.stdout


@pytest.mark.parametrize("B,Nq,Nk,heads,d", [
    (2, 256, 256, 8, 40),      # one q tile per 128, 2 kv tiles
    (1, 4096, 4096, 8, 40),    # 64x64 latent self-attention
    (3, 4096, 512, 8, 40),     # 384 work items on 148 persistent CTAs: two or three items per CTA
    (3, 4000, 500, 8, 40),     # ... ragged queries and keys
    (5, 1024, 512, 8, 80),     # d = 80: 160 work items on 148 persistent CTAs
    (5, 1000, 500, 8, 80),     # ... ragged
    (2, 1024, 1024, 8, 80),
    (2, 256, 256, 8, 160),
    (2, 64, 64, 8, 160),       # 8x8 mid block: fewer queries than a tile
    (2, 1024, 77, 8, 40),      # cross-attention, ragged keys
    (2, 256, 77, 8, 160),
    (1, 1024, 154, 8, 80),     # long-prompt context (2 x 77)
    (1, 300, 200, 8, 40),      # ragged queries and keys
    (1, 300, 200, 8, 80),      # ragged, two-tile d=80 kernel
    (2, 576, 77, 8, 80),       # 768x768 level-1 cross-attention
    (1, 2304, 2304, 4, 80),    # 768x768 level-1 self-attention
    (16, 4096, 77, 8, 40),     # resident-K/V cross-attention kernel at the benchmark shape
    (3, 300, 77, 8, 40),       # ... ragged queries
    (1, 1024, 128, 8, 80),     # ... full key tile
    (2, 512, 16, 8, 40),       # ... tiny contexts
    (1, 256, 1, 8, 40),
    (2, 9216, 77, 8, 40),      # 768x768 level-0 cross-attention
])
def test_attention_matches_torch(engine, B, Nq, Nk, heads, d):
    _check_attention(engine, B, Nq, Nk, heads, d, legacy=False)


@pytest.mark.parametrize("B,Nq,Nk,heads,d", [(1, 512, 384, 8, 40), (2, 1024, 77, 8, 40), (1, 9216, 9216, 2, 40)])
def test_attention_d40_kernels_agree(engine, B, Nq, Nk, heads, d):
    """the two-query-tile kernel (default for d=40) and the one-tile kernel, each vs torch; 9216 = 96x96 latent (768^2)"""
    _check_attention(engine, B, Nq, Nk, heads, d, legacy=True)
    _check_attention(engine, B, Nq, Nk, heads, d, legacy=False)


@pytest.mark.parametrize("B,N", [(2, 256), (1, 320), (1, 64), (1, 4096), (1, 9216)])
def test_vae_attention_flash_kernel_matches_torch(engine, B, N):
    """single head of 512 (layers.py:28-59): 16x16 / ragged 20x16 / 8x8 latents, 64x64 (512x512 images), 96x96 (768x768)"""
    _check_attention(engine, B, N, N, 1, 512, legacy=False)


def _check_attention(engine, B, Nq, Nk, heads, d, legacy):
    g = torch.Generator().manual_seed(B * 1000 + Nq + Nk + d)
    C = heads * d
    q = _bf(torch.randn(B, Nq, C, generator=g))
    k = _bf(torch.randn(B, Nk, C, generator=g))
    v = _bf(torch.randn(B, Nk, C, generator=g))
    out = np.empty((B, Nq, C), np.float32)
    dl = _lib.DL()
    engine._check(engine._lib.sdtf_test_attention(engine._h, dl(q.numpy()), dl(k.numpy()), dl(v.numpy()),
                                                   -heads if legacy else heads, dl(out)))
    qh = q.view(B, Nq, heads, d).permute(0, 2, 1, 3).cuda()
    kh = k.view(B, Nk, heads, d).permute(0, 2, 1, 3).cuda()
    vh = v.view(B, Nk, heads, d).permute(0, 2, 1, 3).cuda()
    s = torch.matmul(qh, kh.transpose(-1, -2)) * d ** -0.5
    ref = torch.matmul(torch.softmax(s, -1), vh).permute(0, 2, 1, 3).reshape(B, Nq, C).cpu().numpy()
    assert np.isfinite(out).all()
    # relative bars (the output of a long softmax average is small: absmax 0.13 at 4096 keys, 0.08 at 9216):
    # max|d| / max|ref| <= 2e-2 (the eps bar of BASELINE.json) and rms(d) / rms(ref) <= 1e-2
    d = out.astype(np.float64) - ref
    rel_max = np.abs(d).max() / np.abs(ref).max()
    rel_rms = np.sqrt(np.mean(d * d)) / np.sqrt(np.mean(ref.astype(np.float64) ** 2))
    assert rel_max <= 2e-2, f"max|d|/max|ref| = {rel_max:.4g} (max|ref| {np.abs(ref).max():.4g})"
    assert rel_rms <= 1e-2, f"rms(d)/rms(ref) = {rel_rms:.4g}"


@pytest.mark.parametrize("B,H,W,C,mode", [
    (2, 64, 64, 320, 1), (2, 32, 32, 640, 0), (2, 16, 16, 1280, 1), (2, 8, 8, 2560, 1), (1, 32, 32, 960, 1),
    (1, 16, 16, 1920, 1), (1, 128, 128, 128, 1), (1, 64, 64, 512, 0), (3, 24, 24, 256, 1),
    (2, 64, 64, 320, 2), (2, 32, 32, 640, 2), (2, 16, 16, 1280, 2), (1, 5, 7, 320, 2), (1, 5, 7, 640, 2), (1, 3, 3, 1280, 2),
    (1, 8, 8, 768, 2), (16, 8, 8, 1280, 1), (18, 16, 16, 640, 0), (20, 8, 8, 320, 1),
])
def test_norms_match_torch(engine, B, H, W, C, mode):
    g = torch.Generator().manual_seed(C + mode)
    x = _bf(torch.randn(B, H, W, C, generator=g) * 1.5 + 0.3)
    gamma = 1 + 0.1 * torch.randn(C, generator=g)
    beta = 0.05 * torch.randn(C, generator=g)
    out = np.empty((B, H, W, C), np.float32)
    dl = _lib.DL()
    engine._check(engine._lib.sdtf_test_norm(engine._h, dl(x.numpy()), dl(gamma.numpy()), dl(beta.numpy()), mode, dl(out)))
    if mode == 2:
        ref = torch.nn.functional.layer_norm(x, (C,), gamma, beta, eps=1e-5)
    else:
        ref = torch.nn.functional.group_norm(x.permute(0, 3, 1, 2), 32, gamma, beta, eps=1e-5)
        if mode == 1:
            ref = torch.nn.functional.silu(ref)
        ref = ref.permute(0, 2, 3, 1)
    err = (torch.as_tensor(out) - ref).abs().max().item()
    assert err < 3e-2, err


def test_engine_rejects_bad_arguments(engine):
    """error behaviour at the boundary: bad shapes come back as error codes with a message, not crashes"""
    from minsdtf_b200.engine import EngineError
    with pytest.raises(EngineError):
        engine.unet(np.zeros((1, 16, 16, 3), np.float32), np.zeros((1, 320), np.float32), np.zeros((1, 77, 768), np.float32))
    with pytest.raises(EngineError):
        engine._check(engine._lib.sdtf_finalize_weights(engine._h, b"no_such_component"))
