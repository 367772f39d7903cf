// gemm.cuh — the one tensor-core kernel every contraction in the SD1.5 path goes through:
//   D[M, N] = sum_{tap, k} A_tap[M, k] * Wt[tap][N, k]   (+ fused epilogue)
// covering  3x3 / 1x1 convolutions as implicit GEMM (reference: layers.py:17-25 PaddedConv2D and all of
// its call sites in diffusion_model.py / control_net.py / image_decoder.py / image_encoder.py) and every
// Dense layer (diffusion_model.py:30,62,67,90,102-108,146).
//
// B200 mapping
//   * A (activations, NHWC bf16) is fetched by TMA as a 4-D box (64 ch, bw, bh, bn) with bw*bh*bn = 128
//     output pixels; the filter tap is a coordinate shift of the box and the zero padding of the
//     convolution is TMA's out-of-bounds zero fill, so there is no im2col buffer and no halo logic.
//     Stride-2 convolutions use the tensor map's element strides.  A channel concat (diffusion_model.py:
//     237-273) is a second tensor map walked after the first in the K loop (or one strided view).
//   * Wt (weights) is pre-packed at load time as [tap][N][K] bf16, fetched as a (64, BN) box.
//   * Both land in 128B-swizzled shared memory; one elected thread issues tcgen05.mma (M=128, N=BN,
//     K=16) accumulating fp32 in TMEM; tcgen05.commit releases smem stages / publishes the accumulator.
//   * Epilogue: TMEM -> registers (tcgen05.ld 32x32b), + bias, + per-sample time-embedding column,
//     + residual, SiLU / GEGLU gate, -> bf16 (or fp32) NHWC stores, optionally into a channel slice of a
//     wider (concat) buffer.
#pragma once
#include "tc05.cuh"

namespace sdtf {

enum : int { ACT_NONE = 0, ACT_SILU = 1, ACT_GEGLU = 2, ACT_QGELU = 3 };  // QGELU: x * sigmoid(1.702 x) (CLIP text tower)

struct GemmParams {
  // output pixel space [B][H][W] and the 128-pixel tile shape
  int W, H, B;
  int bw, bh, bn;
  int tiles_x, tiles_y;  // tiles along W and H (tiles along B = gridDim.x / (tiles_x * tiles_y))
  // K loop
  int taps, tap_w;   // number of filter taps and filter width (tap -> (r, s) = (tap / tap_w, tap % tap_w))
  int pad_x, pad_y;  // left / top zero padding
  int stride;        // input coordinate = output coordinate * stride + tap offset - pad
  int kc0, kc1;      // 64-channel chunks taken from source 0 then source 1
  // N
  int N, BN;
  int tmem_cols;
  int stages;
  // epilogue
  const float* bias;          // [N] (already permuted for GEGLU) or null
  const float* temb;          // [B][temb_ld] per-sample column add (time embedding projection) or null
  int temb_ld;
  const __nv_bfloat16* res;   // residual, NHWC with pixel stride res_ld, or null
  long long res_ld;
  void* out;                  // NHWC, pixel stride out_ld (elements)
  long long out_ld;
  int out_fp32;
  int act;
  float out_scale;            // applied to the accumulator before bias (1.0 = off)
  // ---- LayerNorm folded into the GEMMs around it (gemm3.cuh only; see DESIGN.md "LayerNorm") ----
  // producer side: this GEMM's OUTPUT rows are the input of a LayerNorm: per row and 32-column pass, sum and sum of squares
  // of the (bf16-rounded) outputs -> ln_out[(column / 32) * ln_rows + row]
  float2* ln_out;
  // consumer side: this GEMM's A operand is the UN-normalised LayerNorm input; the weights carry gamma, and the epilogue
  // applies y = rstd (acc - mean c1[n]) + c0[n] with the row statistics summed from ln_slots partials
  const float2* ln_in;
  const float* ln_c1;         // [N]: sum_k of the gamma-scaled, bf16-rounded weight row; c0 sits in `bias`
  int ln_slots;
  long long ln_rows;          // rows (pixels) = distance between partial slots
  float ln_inv_c;             // 1 / (channels normalised)
};

static constexpr int kBM = 128;
static constexpr int kBK = 64;
static constexpr int kATileBytes = kBM * kBK * 2;  // 16 KB

// (__fdividef: the IEEE division's slow-path call + check per element was 32 CALL / FCHK pairs in every epilogue pass)
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }
__device__ __forceinline__ float qgelu_f(float x) { return __fdividef(x, 1.f + __expf(-1.702f * x)); }
// tanh-approximation GELU exactly as the reference spells it (diffusion_model.py:150-153)
__device__ __forceinline__ float gelu_tanh_f(float g) {
  float u = g * 0.7978845608f * (1.f + 0.044715f * g * g);
  return 0.5f * g * (1.f + tanhf(u));
}

__global__ void __launch_bounds__(128)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  using namespace tc05;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = kATileBytes + (uint32_t)p.BN * 128u;
  const uint32_t bar_base = smem_base + (uint32_t)p.stages * stage_bytes;
  // barriers: full[stages], empty[stages], accum_full, then the TMEM base address slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * p.stages);
  const uint32_t tmem_slot = accum_bar + 8u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (size_t)p.stages * stage_bytes + 8u * (2 * p.stages) + 8u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  const int n_tile = blockIdx.y;
  int mt = blockIdx.x;
  const int tx = mt % p.tiles_x; mt /= p.tiles_x;
  const int ty = mt % p.tiles_y; mt /= p.tiles_y;
  const int tb = mt;
  const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb * p.bn;
  const int kchunks = p.kc0 + p.kc1;
  const int iters = p.taps * kchunks;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmB);
    if (p.kc1) prefetch_tmap(&tmA1);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot_ptr;

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer =====
      int it = 0;
      for (int tap = 0; tap < p.taps; ++tap) {
        const int r = tap / p.tap_w, s = tap - r * p.tap_w;
        const int cx = x0 * p.stride + s - p.pad_x;
        const int cy = y0 * p.stride + r - p.pad_y;
        for (int kc = 0; kc < kchunks; ++kc, ++it) {
          const int st = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          mbar_wait(empty_bar(st), ph ^ 1u);
          const uint32_t sa = smem_base + (uint32_t)st * stage_bytes;
          const uint32_t sb = sa + kATileBytes;
          mbar_expect_tx(full_bar(st), stage_bytes);
          if (kc < p.kc0) tma_load_4d(sa, &tmA0, full_bar(st), kc * kBK, cx, cy, b0);
          else            tma_load_4d(sa, &tmA1, full_bar(st), (kc - p.kc0) * kBK, cx, cy, b0);
          tma_load_3d(sb, &tmB, full_bar(st), kc * kBK, n_tile * p.BN, tap);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer =====
      const uint32_t idesc = make_idesc_bf16(kBM, p.BN, 0, 0);
      for (int it = 0; it < iters; ++it) {
        const int st = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(full_bar(st), ph);
        fence_after_sync();
        const uint32_t sa = smem_base + (uint32_t)st * stage_bytes;
        const uint32_t sb = sa + kATileBytes;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          const uint64_t da = make_smem_desc_sw128(sa + k * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(sb + k * 32, 16, 1024);
          mma_f16_ss(tmem_acc, da, db, idesc, (it | k) != 0);
        }
        mma_commit(empty_bar(st));  // frees this smem stage once the MMAs above have read it
      }
      mma_commit(accum_bar);  // accumulator complete
    }
  }
  __syncwarp();

  // ===== epilogue: all 4 warps; warp w owns TMEM lanes [32w, 32w+32) = tile rows =====
  mbar_wait(accum_bar, 0);
  fence_after_sync();

  const int row = warp * 32 + lane;
  int rr = row;
  const int lx = rr % p.bw; rr /= p.bw;
  const int ly = rr % p.bh; rr /= p.bh;
  const int ox = x0 + lx, oy = y0 + ly, ob = b0 + rr;
  const bool row_ok = (ox < p.W) && (oy < p.H) && (ob < p.B);
  const long long pix = ((long long)ob * p.H + oy) * p.W + ox;
  const uint32_t t_row = tmem_acc + ((uint32_t)(warp * 32) << 16);

  const bool geglu = (p.act == ACT_GEGLU);
  const int ncols = geglu ? p.BN / 2 : p.BN;          // output columns produced by this tile
  const int col_base = n_tile * ncols;                // first output column
  const int Nout = geglu ? p.N / 2 : p.N;

  for (int c = 0; c < ncols; c += 16) {
    uint32_t v[16];
    float f[16];
    __syncwarp();  // tcgen05.ld is warp-collective: reconverge after the masked stores below
    tmem_ld16(t_row + (uint32_t)c, v);
    if (geglu) {
      uint32_t g[16];
      tmem_ld16(t_row + (uint32_t)(ncols + c), g);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float val = __uint_as_float(v[i]), gate = __uint_as_float(g[i]);
        if (p.bias) {
          val += __ldg(p.bias + n_tile * p.BN + c + i);
          gate += __ldg(p.bias + n_tile * p.BN + ncols + c + i);
        }
        f[i] = val * gelu_tanh_f(gate);
      }
    } else {
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * p.out_scale;
      const int col = col_base + c;
      if (col + 16 <= Nout) {
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + col + i));
            f[i] += bv.x; f[i + 1] += bv.y; f[i + 2] += bv.z; f[i + 3] += bv.w;
          }
        }
        if (p.temb && row_ok) {
          const float* tp = p.temb + (long long)ob * p.temb_ld + col;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 tv = __ldg(reinterpret_cast<const float4*>(tp + i));
            f[i] += tv.x; f[i + 1] += tv.y; f[i + 2] += tv.z; f[i + 3] += tv.w;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (col + i < Nout) {
            if (p.bias) f[i] += __ldg(p.bias + col + i);
            if (p.temb && row_ok) f[i] += __ldg(p.temb + (long long)ob * p.temb_ld + col + i);
          }
        }
      }
    }
    if (row_ok) {
    const int col = col_base + c;
    const bool full = (col + 16 <= Nout);
    if (p.res) {
      const __nv_bfloat16* rp = p.res + pix * p.res_ld + col;
      if (full && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
        const uint4 r0 = *reinterpret_cast<const uint4*>(rp);
        const uint4 r1 = *reinterpret_cast<const uint4*>(rp + 8);
        const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&rw[i]);
          f[2 * i] += __bfloat162float(h.x);
          f[2 * i + 1] += __bfloat162float(h.y);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (col + i < Nout) f[i] += __bfloat162float(rp[i]);
      }
    }
    if (p.act == ACT_SILU) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = silu_f(f[i]);
    }
    if (p.act == ACT_QGELU) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = qgelu_f(f[i]);
    }
    if (p.out_fp32) {
      float* op = reinterpret_cast<float*>(p.out) + pix * p.out_ld + col;
      if (full && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(op + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (col + i < Nout) op[i] = f[i];
      }
    } else {
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_ld + col;
      if (full && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
        uint4 o0, o1;
        o0.x = pack_bf16(f[0], f[1]);   o0.y = pack_bf16(f[2], f[3]);
        o0.z = pack_bf16(f[4], f[5]);   o0.w = pack_bf16(f[6], f[7]);
        o1.x = pack_bf16(f[8], f[9]);   o1.y = pack_bf16(f[10], f[11]);
        o1.z = pack_bf16(f[12], f[13]); o1.w = pack_bf16(f[14], f[15]);
        *reinterpret_cast<uint4*>(op) = o0;
        *reinterpret_cast<uint4*>(op + 8) = o1;
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (col + i < Nout) op[i] = __float2bfloat16(f[i]);
      }
    }
    }  // row_ok
  }

  // teardown: everyone done reading TMEM before the allocating warp frees it
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_acc, (uint32_t)p.tmem_cols);
}

}  // namespace sdtf
