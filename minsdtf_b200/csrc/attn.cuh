// attn.cuh — flash-style multi-head attention on tcgen05 for the SpatialTransformer blocks
// (reference: diffusion_model.py:99-129 CrossAttention — softmax(Q K^T * d^-1/2) V, 8 heads, never
// materialising the (B, 8, N, N) score tensor the reference builds at :123-126).
//
// One CTA = 128 query rows of one (batch, head).  Per 128-key tile:
//   S = Q K^T        tcgen05.mma  M=128, N=keys(<=128), K=d     -> TMEM columns [0,128)
//   online softmax   one thread per query row: tcgen05.ld S, running max / sum in fp32 registers,
//                    P = exp2(S*scale - m) written as bf16 into 128B-swizzled smem (the A operand of PV)
//   O (+)= P V       tcgen05.mma  M=128, N=d, K=keys; V is read straight from its [key][d] layout as an
//                    MN-major B operand, O accumulates in TMEM columns [128, 128+d) and is rescaled in
//                    place (tcgen05.ld / st) when the running max moves.
// Q/K/V tiles arrive by TMA through 4-D maps (column within the head, head, token, batch): a head's 64-column box is
// clipped at the head size d by the hardware (d = 40: columns 40..63 of the shared-memory row are zero-filled, nothing
// is padded in HBM), and ragged key counts (77-token context) are zero-filled the same way and masked in the softmax.
#pragma once
#include "common.cuh"
#include "tc05.cuh"

namespace sdtf {

struct AttnParams {
  int Nq, Nk;        // queries / keys per batch element
  int d;             // true head size (scale = d^-1/2, output columns per head)
  int dstride;       // column distance between heads in the Q/K/V matrices (= d: heads are stored densely)
  float scale_log2;  // d^-1/2 * log2(e)
  bf16* out;         // [B*Nq][ldo], head h at column h*d
  long long ldo;
  long long* prof;   // optional [16] global cycle counters summed over CTAs (SDTF_ATTN_PROFILE=1), null: off
  int debug;         // timing experiments only (SDTF_ATTN_DEBUG; results are garbage): 1 no exp, 2 no max, 4 no S load, 8 no P store, 16 no PV MMAs, 32 no QK MMAs, 64 no K/V TMA
  int heads, n_qt, n_items;  // persistent kernels (attn2h): work items = (batch, head, 256-query tile), item = (b * heads + head) * n_qt + qt
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// DCH: 64-column chunks per head (1: d<=64, 2: d<=128, 3: d<=192); KS: QK^T k-steps = ceil(d/16);
// DV: PV N extent (multiple of 16 >= d); KST/VST: K and V smem stages.
template <int DCH, int KS, int DV, int KST, int VST>
__global__ void __launch_bounds__(128)
attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
            const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using namespace tc05;
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t kChunk = 128 * 128;  // one [128 rows][64 bf16] swizzled chunk = 16 KB
  constexpr uint32_t kTmemCols = (128 + DV <= 256) ? 256 : 512;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sK = sQ + DCH * kChunk;
  const uint32_t sV = sK + KST * DCH * kChunk;
  const uint32_t sP = sV + VST * DCH * kChunk;
  const uint32_t bars = sP + 2 * kChunk;
  const uint32_t q_full = bars, s_full = bars + 8, o_full = bars + 16;
  auto k_full = [&](int s) { return bars + 24u + 8u * s; };
  auto v_full = [&](int s) { return bars + 24u + 8u * (KST + s); };
  const uint32_t tmem_slot = bars + 24u + 8u * (KST + VST);
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* genP = gen + (sP - base);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, head = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.Nk + 127) / 128;
  const bool leader = threadIdx.x == 0;

  pdl_trigger();
  if (leader) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1); mbar_init(s_full, 1); mbar_init(o_full, 1);
    for (int s = 0; s < KST; ++s) mbar_init(k_full(s), 1);
    for (int s = 0; s < VST; ++s) mbar_init(v_full(s), 1);
    fence_mbar_init();
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(tmem_slot, kTmemCols); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_wait();  // Q / K / V come from the projection GEMMs
  const uint32_t tS = tmem, tO = tmem + 128;

  auto load_k = [&](int j) {
    const int st = j % KST;
    mbar_expect_tx(k_full(st), DCH * kChunk);
    for (int c = 0; c < DCH; ++c) tma_load_4d(sK + (st * DCH + c) * kChunk, &tmK, k_full(st), 64 * c, head, j * 128, b);
  };
  auto load_v = [&](int j) {
    const int st = j % VST;
    mbar_expect_tx(v_full(st), DCH * kChunk);
    for (int c = 0; c < DCH; ++c) tma_load_4d(sV + (st * DCH + c) * kChunk, &tmV, v_full(st), 64 * c, head, j * 128, b);
  };
  auto keys_in_tile = [&](int j) {  // valid keys of tile j rounded up to the MMA granularity
    int n = p.Nk - j * 128;
    n = n > 128 ? 128 : n;
    return (n + 15) & ~15;
  };
  auto issue_s = [&](int j) {
    const int st = j % KST;
    mbar_wait(k_full(st), (uint32_t)(j / KST) & 1u);
    fence_after_sync();
    const uint32_t idesc = make_idesc_bf16(128, keys_in_tile(j), 0, 0);
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const uint32_t off = (uint32_t)(k >> 2) * kChunk + (uint32_t)(k & 3) * 32u;
      mma_f16_ss(tS, make_smem_desc_sw128(sQ + off, 16, 1024),
                 make_smem_desc_sw128(sK + st * DCH * kChunk + off, 16, 1024), idesc, k != 0);
    }
    mma_commit(s_full);
  };

  if (leader) {
    mbar_expect_tx(q_full, DCH * kChunk);
    for (int c = 0; c < DCH; ++c) tma_load_4d(sQ + c * kChunk, &tmQ, q_full, 64 * c, head, q0, b);
    for (int j = 0; j < KST && j < nkv; ++j) load_k(j);
    for (int j = 0; j < VST && j < nkv; ++j) load_v(j);
    mbar_wait(q_full, 0);
    issue_s(0);
  }
  __syncwarp();

  const int row = warp * 32 + lane;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  float m_run = -INFINITY, l_run = 0.f;

  for (int j = 0; j < nkv; ++j) {
    const int nk_valid = min(128, p.Nk - j * 128);
    const int ncols = (nk_valid + 15) & ~15;
    // ---- S_j ready ----
    mbar_wait(s_full, (uint32_t)j & 1u);
    fence_after_sync();
    if (leader && j + KST < nkv) load_k(j + KST);  // K stage of tile j is free once S_j has completed
    __syncwarp();
    // pass 1: row max
    float mx = -INFINITY;
    for (int c = 0; c < ncols; c += 16) {
      uint32_t v[16];
      tmem_ld16(tS + lane_off + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c + i < nk_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
    }
    const float m_new = fmaxf(m_run, mx * p.scale_log2);
    const float alpha = ex2f(m_run - m_new);
    // ---- PV_{j-1} done: O may be rescaled, P / V stage may be overwritten ----
    if (j > 0) {
      mbar_wait(o_full, (uint32_t)(j - 1) & 1u);
      fence_after_sync();
      if (leader && j - 1 + VST < nkv) load_v(j - 1 + VST);
      __syncwarp();
      for (int c = 0; c < DV; c += 16) {
        uint32_t v[16];
        tmem_ld16(tO + lane_off + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
        tmem_st16(tO + lane_off + c, v);
      }
      tmem_st_wait();
    }
    // pass 2: P = exp2(S*scale - m), row sum, bf16 P -> swizzled smem
    float lsum = 0.f;
    for (int c = 0; c < ncols; c += 16) {
      uint32_t v[16];
      tmem_ld16(tS + lane_off + c, v);
      tmem_ld_wait();
      float pf[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float e = ex2f(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new));
        pf[i] = (c + i < nk_valid) ? e : 0.f;
        lsum += pf[i];
      }
      const int chunk = c >> 6;
      const int u0 = (c & 63) >> 3;  // first 16-byte unit inside the 128-byte row
      uint8_t* rowp = genP + chunk * kChunk + row * 128;
      uint4 w0, w1;
      w0.x = pack_bf16(pf[0], pf[1]);   w0.y = pack_bf16(pf[2], pf[3]);
      w0.z = pack_bf16(pf[4], pf[5]);   w0.w = pack_bf16(pf[6], pf[7]);
      w1.x = pack_bf16(pf[8], pf[9]);   w1.y = pack_bf16(pf[10], pf[11]);
      w1.z = pack_bf16(pf[12], pf[13]); w1.w = pack_bf16(pf[14], pf[15]);
      *reinterpret_cast<uint4*>(rowp + (((u0) ^ (row & 7)) << 4)) = w0;
      *reinterpret_cast<uint4*>(rowp + (((u0 + 1) ^ (row & 7)) << 4)) = w1;
    }
    l_run = l_run * alpha + lsum;
    m_run = m_new;
    fence_proxy_async_smem();  // P (generic proxy) -> visible to the tensor core's async proxy
    fence_before_sync();
    __syncthreads();
    if (leader) {
      fence_after_sync();
      const int st = j % VST;
      mbar_wait(v_full(st), (uint32_t)(j / VST) & 1u);
      fence_after_sync();
      const uint32_t idesc = make_idesc_bf16(128, DV, 0, 1);  // B = V is MN-major
      const int ksteps = ncols >> 4;
      for (int k = 0; k < ksteps; ++k) {
        const uint64_t da = make_smem_desc_sw128(sP + (uint32_t)(k >> 2) * kChunk + (uint32_t)(k & 3) * 32u, 16, 1024);
        const uint64_t db = make_smem_desc_sw128(sV + st * DCH * kChunk + (uint32_t)k * 2048u, kChunk, 1024);
        mma_f16_ss(tO, da, db, idesc, (j | k) != 0);
      }
      mma_commit(o_full);
      if (j + 1 < nkv) issue_s(j + 1);  // S_{j+1} queues behind PV_j on the tensor pipe
    }
    __syncwarp();
  }

  // ---- epilogue: O / l -> bf16 ----
  mbar_wait(o_full, (uint32_t)(nkv - 1) & 1u);
  fence_after_sync();
  const float inv_l = 1.f / l_run;
  const int q = q0 + row;
  const bool ok = q < p.Nq;
  bf16* orow = p.out + ((long long)b * p.Nq + q) * p.ldo + head * p.d;
  for (int c = 0; c < DV; c += 16) {
    uint32_t v[16];
    __syncwarp();
    tmem_ld16(tO + lane_off + c, v);
    tmem_ld_wait();
    if (ok) {
      uint4 w0, w1;
      w0.x = pack_bf16(__uint_as_float(v[0]) * inv_l, __uint_as_float(v[1]) * inv_l);
      w0.y = pack_bf16(__uint_as_float(v[2]) * inv_l, __uint_as_float(v[3]) * inv_l);
      w0.z = pack_bf16(__uint_as_float(v[4]) * inv_l, __uint_as_float(v[5]) * inv_l);
      w0.w = pack_bf16(__uint_as_float(v[6]) * inv_l, __uint_as_float(v[7]) * inv_l);
      w1.x = pack_bf16(__uint_as_float(v[8]) * inv_l, __uint_as_float(v[9]) * inv_l);
      w1.y = pack_bf16(__uint_as_float(v[10]) * inv_l, __uint_as_float(v[11]) * inv_l);
      w1.z = pack_bf16(__uint_as_float(v[12]) * inv_l, __uint_as_float(v[13]) * inv_l);
      w1.w = pack_bf16(__uint_as_float(v[14]) * inv_l, __uint_as_float(v[15]) * inv_l);
      if (c + 8 <= p.d) *reinterpret_cast<uint4*>(orow + c) = w0;
      if (c + 16 <= p.d) *reinterpret_cast<uint4*>(orow + c + 8) = w1;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}


// ------------------------------------------------------------------------------------------------------
// d <= 64 heads (SD1.5: d = 40 at the 64x64 / 96x96 level, N = 4096 / 9216 tokens): the softmax exponentials,
// not the tensor core, bound this shape (128 ex2 per row and tile on a 16-lane MUFU vs 384 MMA cycles), so the
// kernel is organised around keeping the MUFU pipe busy:
//   * one CTA owns TWO 128-row query tiles; warps 0-3 / 4-7 are the softmax groups of tile 0 / 1, warps 8 / 9 issue
//     the tcgen05.mma of group 0 / 1 (each group is its own pipeline, so the two drift out of phase and one group's
//     exponentials fill the MUFU while the other loads / reduces), warp 10 issues every TMA load.  K/V tiles are
//     fetched once for both query tiles.
//   * S_g lives in TMEM columns [128g, 128g+128).  A softmax thread pulls its whole 128-column row into
//     registers and immediately hands the TMEM buffer back (s_empty), so Q K^T of the NEXT key tile runs while
//     the exponentials of this one are computed.
//   * a quarter of the exponentials is evaluated on the FMA pipe (Cody-Waite split + cubic, rel. error 6e-4, below
//     the bf16 rounding of P) instead of MUFU.EX2 (ncu: xu pipe 54 % busy, the binding unit of this kernel).
//   * O_g accumulates in TMEM columns [256+64g, +DV) across all key tiles.  The running max is only moved (and
//     O rescaled through tcgen05.ld/st) when it grows by more than 2^8; otherwise P is computed against the stale
//     max (values <= 256, exact after the final division by the row sum which uses the same max).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// 2^x on the FMA / ALU pipes: x = n + f with n = round(x) (magic-number add), 2^f by a cubic on [-0.5, 0.5],
// 2^n by an integer add into the exponent field.  Valid for x in [-126, 126].
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;  // 1.5 * 2^23: the low mantissa bits of t hold n
  const float f = x - (t - 12582912.f);
  const float p = fmaf(fmaf(fmaf(0.0555041f, f, 0.2402265f), f, 0.6931472f), f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

using tc05::f32x2; using tc05::pack2; using tc05::unpack2; using tc05::fma2; using tc05::add2;  // packed fp32 pairs (tc05.cuh)
// 2^x for a pair on the FMA / ALU pipes (same Cody-Waite split + cubic as ex2_poly); inputs are clamped to >= -126
__device__ __forceinline__ void ex2_poly2(f32x2 x2, float& e0, float& e1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  x2 = pack2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const f32x2 magic = pack2(12582912.f, 12582912.f), nmagic = pack2(-12582912.f, -12582912.f);
  const f32x2 t = add2(x2, magic);                       // low mantissa bits of t hold n = round(x)
  const f32x2 f = fma2(add2(t, nmagic), pack2(-1.f, -1.f), x2);  // x - n in [-0.5, 0.5]
  f32x2 p = fma2(pack2(0.0555041f, 0.0555041f), f, pack2(0.2402265f, 0.2402265f));
  p = fma2(p, f, pack2(0.6931472f, 0.6931472f));
  p = fma2(p, f, pack2(1.0f, 1.0f));
  float p0, p1, t0, t1;
  unpack2(p, p0, p1);
  unpack2(t, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

static constexpr int kA2Threads = 352;
#define A2_TIMED(acc, stmt)                    \
  do {                                         \
    if (prof_on) {                             \
      const long long _t0 = clock64();         \
      stmt;                                    \
      acc += clock64() - _t0;                  \
    } else {                                   \
      stmt;                                    \
    }                                          \
  } while (0)

// DCH = 64-column chunks per head row (1: d <= 64, 2: d = 80), KS = k-steps of Q K^T, DV = accumulator columns of P V,
// KST / VST = K / V ring depth.  d = 80 (the 32x32 level, N = 1024) runs <2, 5, 80, 2, 1>: 14 chunks of shared memory;
// V is single-buffered because P V_j only needs V_j one softmax pass after Q K^T_j needed K_j.
template <int DCH, int KS, int DV, int KST, int VST>
__global__ void __launch_bounds__(kA2Threads, 1)
attn2q_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using namespace tc05;
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t kChunk = 128 * 128;  // [128 rows][64 bf16] swizzled = 16 KB
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kOStride = DV <= 64 ? 64u : 128u;  // accumulator g at TMEM column 256 + kOStride * g
  static_assert(DV <= 128, "S (2 x 128 columns) and O (2 x kOStride) must fit the 512 TMEM columns");
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;                          // 2 tiles x DCH chunks
  const uint32_t sK = sQ + 2 * DCH * kChunk;         // KST x DCH chunks
  const uint32_t sV = sK + KST * DCH * kChunk;       // VST x DCH chunks
  const uint32_t sP = sV + VST * DCH * kChunk;       // 2 groups x 2 chunks
  const uint32_t bars = sP + 4 * kChunk;
  const uint32_t q_full = bars;
  auto k_full = [&](int s) { return bars + 8u + 8u * s; };
  auto k_empty = [&](int s) { return bars + 8u + 8u * (KST + s); };
  auto v_full = [&](int s) { return bars + 8u + 8u * (2 * KST + s); };
  auto v_empty = [&](int s) { return bars + 8u + 8u * (2 * KST + VST + s); };
  const uint32_t gbars = bars + 8u + 8u * (2 * KST + 2 * VST);
  auto s_full = [&](int g) { return gbars + 8u * g; };
  auto s_empty = [&](int g) { return gbars + 16u + 8u * g; };
  auto p_full = [&](int g) { return gbars + 32u + 8u * g; };
  auto o_done = [&](int g) { return gbars + 48u + 8u * g; };
  const uint32_t tmem_slot = gbars + 64u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, head = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.Nk + 127) / 128;
  const bool prof_on = p.prof != nullptr;
  const long long t_start = prof_on ? clock64() : 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < KST; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 2); }  // released by both groups' MMA warps
    for (int s = 0; s < VST; ++s) { mbar_init(v_full(s), 1); mbar_init(v_empty(s), 2); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1); mbar_init(s_empty(g), 4);  // one arrive per softmax warp
      mbar_init(p_full(g), 4); mbar_init(o_done(g), 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot_ptr;

  auto keys_in_tile = [&](int j) {  // valid keys of tile j rounded up to the MMA granularity
    int n = p.Nk - j * 128;
    n = n > 128 ? 128 : n;
    return (n + 15) & ~15;
  };

  if (warp == 10) {
    // ===== TMA producer =====
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * DCH * kChunk);
#pragma unroll
      for (int c = 0; c < DCH; ++c) {
        tma_load_4d(sQ + c * kChunk, &tmQ, q_full, 64 * c, head, q0, b);
        tma_load_4d(sQ + (DCH + c) * kChunk, &tmQ, q_full, 64 * c, head, q0 + 128, b);
      }
      long long w_ke = 0, w_ve = 0;
      for (int j = 0; j < nkv; ++j) {
        const int ks = j % KST, vs = j % VST;
        A2_TIMED(w_ke, mbar_wait(k_empty(ks), ((uint32_t)(j / KST) & 1u) ^ 1u));
        if (p.debug & 64) mbar_arrive(k_full(ks));
        else {
          mbar_expect_tx(k_full(ks), DCH * kChunk);
#pragma unroll
          for (int c = 0; c < DCH; ++c) tma_load_4d(sK + (ks * DCH + c) * kChunk, &tmK, k_full(ks), 64 * c, head, j * 128, b);
        }
        A2_TIMED(w_ve, mbar_wait(v_empty(vs), ((uint32_t)(j / VST) & 1u) ^ 1u));
        if (p.debug & 64) mbar_arrive(v_full(vs));
        else {
          mbar_expect_tx(v_full(vs), DCH * kChunk);
#pragma unroll
          for (int c = 0; c < DCH; ++c) tma_load_4d(sV + (vs * DCH + c) * kChunk, &tmV, v_full(vs), 64 * c, head, j * 128, b);
        }
      }
      if (prof_on) {
        atomicAdd((unsigned long long*)p.prof + 0, (unsigned long long)w_ke);
        atomicAdd((unsigned long long*)p.prof + 1, (unsigned long long)w_ve);
        atomicAdd((unsigned long long*)p.prof + 2, (unsigned long long)(clock64() - t_start));
      }
    }
    __syncwarp();
  } else if (warp == 8 || warp == 9) {
    // ===== MMA issuer of group g = warp - 8 =====
    if (elect_one()) {
      const int g = warp - 8;
      const uint32_t tS = tmem + 128u * g, tO = tmem + 256u + kOStride * g;
      const uint64_t dq = make_smem_desc_sw128(sQ + g * DCH * kChunk, 16, 1024);
      const uint64_t dp = make_smem_desc_sw128(sP + (uint32_t)(2 * g) * kChunk, 16, 1024);
      auto issue_qk = [&](int j) {
        const int st = j % KST;
        const uint32_t idesc = make_idesc_bf16(128, keys_in_tile(j), 0, 0);
        const uint64_t dk = make_smem_desc_sw128(sK + st * DCH * kChunk, 16, 1024);
        if (!(p.debug & 32)) {
#pragma unroll
          for (int k = 0; k < KS; ++k) {  // k-step k: chunk k / 4, 32 bytes per step inside the 128-byte swizzled row
            const uint64_t off = (uint64_t)((k >> 2) * (kChunk >> 4) + (k & 3) * 2);
            mma_f16_ss(tS, dq + off, dk + off, idesc, k != 0);
          }
        }
        mma_commit(s_full(g));
        mma_commit(k_empty(st));
      };
      long long w_kf = 0, w_se = 0, w_vf = 0, w_pf = 0;
      A2_TIMED(w_kf, mbar_wait(q_full, 0));
      A2_TIMED(w_kf, mbar_wait(k_full(0), 0));
      fence_after_sync();
      issue_qk(0);
      const uint32_t idesc_pv = make_idesc_bf16(128, DV, 0, 1);  // B = V is MN-major
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) {
          A2_TIMED(w_kf, mbar_wait(k_full((j + 1) % KST), (uint32_t)((j + 1) / KST) & 1u));
          A2_TIMED(w_se, mbar_wait(s_empty(g), (uint32_t)j & 1u));  // S_j has been pulled into registers
          fence_after_sync();
          issue_qk(j + 1);
        }
        const int vs = j % VST;
        A2_TIMED(w_vf, mbar_wait(v_full(vs), (uint32_t)(j / VST) & 1u));
        A2_TIMED(w_pf, mbar_wait(p_full(g), (uint32_t)j & 1u));  // P_j is in shared memory
        fence_after_sync();
        const int ksteps = keys_in_tile(j) >> 4;
        for (int k = 0; k < ksteps && !(p.debug & 16); ++k) {
          const uint64_t da = dp + (uint64_t)((k >> 2) * (kChunk >> 4) + (k & 3) * 2);
          const uint64_t db = make_smem_desc_sw128(sV + vs * DCH * kChunk + (uint32_t)k * 2048u, kChunk, 1024);
          mma_f16_ss(tO, da, db, idesc_pv, (j | k) != 0);
        }
        mma_commit(o_done(g));
        mma_commit(v_empty(vs));
      }
      if (prof_on && g == 0) {
        atomicAdd((unsigned long long*)p.prof + 3, (unsigned long long)w_kf);
        atomicAdd((unsigned long long*)p.prof + 4, (unsigned long long)w_se);
        atomicAdd((unsigned long long*)p.prof + 5, (unsigned long long)w_vf);
        atomicAdd((unsigned long long*)p.prof + 6, (unsigned long long)w_pf);
        atomicAdd((unsigned long long*)p.prof + 7, (unsigned long long)(clock64() - t_start));
      }
    }
    __syncwarp();
  } else {
    // ===== softmax groups =====
    const int g = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + 128u * g + lane_off;
    const uint32_t tO = tmem + 256u + kOStride * g + lane_off;
    uint8_t* rowp = gen + (sP - base) + (uint32_t)(2 * g) * kChunk + row * 128;
    const int sw = row & 7;
    float m_run = -INFINITY, l_run = 0.f;
    long long w_sf = 0, w_od = 0, t_loop = 0;
    for (int j = 0; j < nkv; ++j) {
      const int nk_valid = min(128, p.Nk - j * 128);
      A2_TIMED(w_sf, mbar_wait(s_full(g), (uint32_t)j & 1u));
      fence_after_sync();
      uint32_t sv[128];
      if (!(p.debug & 4)) {
        tmem_ld32_at<0>(tS, sv);
        tmem_ld32_at<32>(tS + 32, sv);
        tmem_ld32_at<64>(tS + 64, sv);
        tmem_ld32_at<96>(tS + 96, sv);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 128; ++i) sv[i] = __float_as_uint(0.001f * (float)(i + lane));
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty(g));
      if (nk_valid < 128) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= nk_valid) sv[i] = 0xff800000u;  // -inf
      }
      float mx = fmax3(__uint_as_float(sv[0]), __uint_as_float(sv[1]), __uint_as_float(sv[2]));
      float mx2 = fmax3(__uint_as_float(sv[3]), __uint_as_float(sv[4]), __uint_as_float(sv[5]));
#pragma unroll
      for (int i = 6; i + 3 < 128; i += 4) {
        mx = fmax3(mx, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
        mx2 = fmax3(mx2, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
      }
      mx = fmax3(mx, mx2, fmaxf(__uint_as_float(sv[126]), __uint_as_float(sv[127])));
      if (p.debug & 2) mx = __uint_as_float(sv[5]);
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const bool grow = (m_new - m_run) > 8.f;
      if (j > 0) {  // PV_{j-1}(g) must be complete before O is touched or P overwritten
        A2_TIMED(w_od, mbar_wait(o_done(g), (uint32_t)(j - 1) & 1u));
        fence_after_sync();
      }
      const long long t_l0 = prof_on ? clock64() : 0;
      if (__any_sync(0xffffffffu, grow)) {
        const float alpha = ex2f(m_run - m_new);
        m_run = m_new;
        l_run *= alpha;
        if (j > 0) {
#pragma unroll
          for (int c = 0; c < DV; c += 16) {
            uint32_t v[16];
            tmem_ld16(tO + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st16(tO + c, v);
          }
          tmem_st_wait();
        }
      }
      const float neg_m = -m_run;
      float lsum0 = 0.f, lsum1 = 0.f;
#pragma unroll
      for (int c = 0; c < 128; c += 8) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xarg = fmaf(__uint_as_float(sv[c + i]), p.scale_log2, neg_m);
          e[i] = (i >= 6) ? ex2_poly(xarg) : ex2f(xarg);  // 2 of 8 on the FMA pipe
          if (p.debug & 1) e[i] = xarg;
        }
        lsum0 += (e[0] + e[1]) + (e[2] + e[3]);
        lsum1 += (e[4] + e[5]) + (e[6] + e[7]);
        uint4 w;
        w.x = pack_bf16(e[0], e[1]); w.y = pack_bf16(e[2], e[3]);
        w.z = pack_bf16(e[4], e[5]); w.w = pack_bf16(e[6], e[7]);
        const int chunk = c >> 6, u = (c & 63) >> 3;
        if (!(p.debug & 8)) *reinterpret_cast<uint4*>(rowp + chunk * kChunk + ((u ^ sw) << 4)) = w;
      }
      l_run += lsum0 + lsum1;
      fence_proxy_async_smem();  // P (generic proxy) -> visible to the tensor core's async proxy
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(g));
      if (prof_on) t_loop += clock64() - t_l0;
    }
    if (prof_on && threadIdx.x == 0) {
      atomicAdd((unsigned long long*)p.prof + 8, (unsigned long long)w_sf);
      atomicAdd((unsigned long long*)p.prof + 9, (unsigned long long)w_od);
      atomicAdd((unsigned long long*)p.prof + 10, (unsigned long long)t_loop);
      atomicAdd((unsigned long long*)p.prof + 11, (unsigned long long)(clock64() - t_start));
    }
    // ---- epilogue: O / l -> bf16 ----
    mbar_wait(o_done(g), (uint32_t)(nkv - 1) & 1u);
    fence_after_sync();
    const float inv_l = 1.f / l_run;
    const int q = q0 + g * 128 + row;
    const bool ok = q < p.Nq;
    bf16* orow = p.out + ((long long)b * p.Nq + q) * p.ldo + head * p.d;
#pragma unroll
    for (int c = 0; c < DV; c += 16) {
      uint32_t v[16];
      tmem_ld16(tO + c, v);
      tmem_ld_wait();
      if (ok) {
        uint4 w0, w1;
        w0.x = pack_bf16(__uint_as_float(v[0]) * inv_l, __uint_as_float(v[1]) * inv_l);
        w0.y = pack_bf16(__uint_as_float(v[2]) * inv_l, __uint_as_float(v[3]) * inv_l);
        w0.z = pack_bf16(__uint_as_float(v[4]) * inv_l, __uint_as_float(v[5]) * inv_l);
        w0.w = pack_bf16(__uint_as_float(v[6]) * inv_l, __uint_as_float(v[7]) * inv_l);
        w1.x = pack_bf16(__uint_as_float(v[8]) * inv_l, __uint_as_float(v[9]) * inv_l);
        w1.y = pack_bf16(__uint_as_float(v[10]) * inv_l, __uint_as_float(v[11]) * inv_l);
        w1.z = pack_bf16(__uint_as_float(v[12]) * inv_l, __uint_as_float(v[13]) * inv_l);
        w1.w = pack_bf16(__uint_as_float(v[14]) * inv_l, __uint_as_float(v[15]) * inv_l);
        if (c + 8 <= p.d) *reinterpret_cast<uint4*>(orow + c) = w0;
        if (c + 16 <= p.d) *reinterpret_cast<uint4*>(orow + c + 8) = w1;
      }
      __syncwarp();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, kTmemCols);
  if (prof_on && threadIdx.x == 0) atomicAdd((unsigned long long*)p.prof + 12, (unsigned long long)(clock64() - t_start));
}

// ------------------------------------------------------------------------------------------------------
// attn2h: the d <= 64 kernel with HALF-ROW softmax threads.  Profiling attn2q (one thread per 128-wide S row) showed
// the softmax warps, not MUFU / tensor / barriers, bound it: ~850 instructions per warp and key tile at IPC 0.38 —
// two warps per scheduler holding 128 live S registers each cannot hide their own latencies.  Here a query row is
// served by TWO threads (keys [0,64) and [64,128) of every key tile), 16 softmax warps in all: four warps per
// scheduler, half the live registers, and the halves never have to agree on a maximum because each keeps its own
// online-softmax state (m, l) and its own accumulator O_{g,h} in TMEM (4 x 48 columns); P V for a half starts as soon
// as that half's 64 columns of P are in shared memory.  The two partial results are merged once, at the end:
// O = (O_a 2^(m_a-m) + O_b 2^(m_b-m)) / (l_a 2^(m_a-m) + l_b 2^(m_b-m)).
// Warps 0-15 softmax (quad = w&3, query tile g = (w>>2)&1, half h = w>>3), 16/17 MMA issue of tile 0/1, 18 TMA.
// ------------------------------------------------------------------------------------------------------
// A value that must stay in its register: routed through a shuffle with the thread's own lane, which ptxas will not
// re-derive (an empty asm with a "+r" operand is transparent to it: it went on rebuilding the addresses from %tid).
__device__ __forceinline__ void pin_reg(uint32_t& x, int lane) { x = __shfl_sync(0xffffffffu, x, lane); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
}
#ifndef SDTF_ATTN_SMX_REGS
#define SDTF_ATTN_SMX_REGS 112
#define SDTF_ATTN_CTL_REGS 32
#endif
static constexpr int kCtlRegs = SDTF_ATTN_CTL_REGS, kSmxRegs = SDTF_ATTN_SMX_REGS;  // setmaxnreg split of 640 x 96 registers (A/B: SDTF_ATTN_SETREG)
static constexpr int kAHThreads = 640;  // 16 softmax warps + one control warpgroup (MMA x2, TMA, idle)
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void softmax_bar_sync() { asm volatile("bar.sync 2, 512;" ::: "memory"); }

// DEFER (A/B switch SDTF_ATTN_DEFER = 0 | 16 | 32): exponentials of the first DEFER keys of a tile are computed into
// registers BEFORE waiting for P V_{j-1} (the wait only guards the P buffer and the rare O rescale).
// PP16: score pairs out of every 8 (16 keys) whose exponentials run on the FMA pipe instead of MUFU.EX2 (2 = a quarter,
// 3 = three eighths, 4 = half); 0 selects the round-1 scalar code (2 of 8 scalar polynomials) for A/B runs.
template <int KS, int DV, int DEFER, int PP16, bool SETREG = false>
__global__ void __launch_bounds__(kAHThreads, 1)
attn2h_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using namespace tc05;
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t kChunk = 128 * 128;  // [128 rows][64 bf16] swizzled = 16 KB
  constexpr int KST = 3, VST = 3;
  constexpr uint32_t kTmemCols = 512;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;                       // 2 chunks
  const uint32_t sK = sQ + 2 * kChunk;            // KST chunks
  const uint32_t sV = sK + KST * kChunk;          // VST chunks
  const uint32_t sP = sV + VST * kChunk;          // chunk 2g+h = P of query tile g, key half h
  const uint32_t xch_off = (sP - base) + 4 * kChunk;  // float2 [2][256]: (m, l) of the h = 1 threads for the final merge (by item parity)
  const uint32_t bars = base + xch_off + 4096u;
  const uint32_t q_full = bars;
  auto k_full = [&](int s) { return bars + 8u + 8u * s; };
  auto k_empty = [&](int s) { return bars + 8u + 8u * (KST + s); };
  auto v_full = [&](int s) { return bars + 8u + 8u * (2 * KST + s); };
  auto v_empty = [&](int s) { return bars + 8u + 8u * (2 * KST + VST + s); };
  const uint32_t gbars = bars + 8u + 8u * (2 * KST + 2 * VST);
  auto s_full = [&](int g) { return gbars + 8u * g; };
  auto s_empty = [&](int g) { return gbars + 16u + 8u * g; };
  auto p_full = [&](int g, int h) { return gbars + 32u + 8u * (2 * g + h); };
  auto o_done = [&](int g, int h) { return gbars + 64u + 8u * (2 * g + h); };
  auto v_ones = [&](int s) { return gbars + 96u + 8u * s; };
  const uint32_t q_empty = gbars + 96u + 8u * VST;   // every Q K^T of the item has completed: Q may be overwritten
  auto o_free = [&](int g) { return gbars + 104u + 8u * VST + 8u * g; };  // the merge has read tile g's accumulators
  const uint32_t tmem_slot = gbars + 120u + 8u * VST;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkv = (p.Nk + 127) / 128;
  // Persistent: CTA c works on items c, c + gridDim.x, ... (consecutive items share a head's K / V).  Barrier phases run on
  // the CTA-wide key-tile counter t, Q / accumulator hand-overs on the item counter, so the TMA and MMA warps run into the
  // next item while the softmax warps merge and store the current one: the ~5 us of launch, TMEM allocation, first loads
  // and epilogue that each 44-us CTA of the one-item-per-CTA grid paid are hidden (SDTF_ATTN_PERSIST=0: one item per CTA).
  auto item_coords = [&](int item, int& q0, int& head, int& b) {
    const int qt = item % p.n_qt, hb = item / p.n_qt;
    q0 = qt * 256; head = hb % p.heads; b = hb / p.heads;
  };

  pdl_trigger();
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 2);  // both tiles' MMA warps
    for (int g = 0; g < 2; ++g) mbar_init(o_free(g), 4);  // the four h = 0 warps of the tile
    for (int s = 0; s < KST; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 2); }  // released by both tiles' MMA warps
    for (int s = 0; s < VST; ++s) { mbar_init(v_full(s), 1); mbar_init(v_empty(s), 2); mbar_init(v_ones(s), 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(s_empty(g), 8);  // one arrive per softmax warp of the tile (both halves)
      for (int h = 0; h < 2; ++h) { mbar_init(p_full(g, h), 4); mbar_init(o_done(g, h), 1); }
    }
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_wait();  // Q / K / V come from the projection GEMMs

  auto keys_in_tile = [&](int j) {  // valid keys of tile j rounded up to the MMA granularity
    int n = p.Nk - j * 128;
    n = n > 128 ? 128 : n;
    return (n + 15) & ~15;
  };

  // register budget: the kernel is launched with 96 registers per thread (640 x 96 = 60 K).  With SETREG every role branch
  // starts with setmaxnreg: the control warpgroup keeps 32 registers, the four softmax warpgroups take 112 — enough for a
  // 64-wide S row, 32 packed results in flight and the loop-invariant addresses without spills (0.698 -> 0.617 ms at
  // 64x64 together with the pinned addresses and the direct P stores).  The instruction must sit INSIDE each branch: placed
  // before the role dispatch ptxas allocates the whole kernel under the smaller of the two budgets.

  if (warp == 18) {
    // ===== TMA producer =====
    if (SETREG) setmaxnreg_dec<kCtlRegs>();
    if (elect_one()) {
      int t = 0, it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        int q0, head, b;
        item_coords(item, q0, head, b);
        if (it > 0) mbar_wait(q_empty, (uint32_t)(it - 1) & 1u);
        mbar_expect_tx(q_full, 2 * kChunk);
        tma_load_4d(sQ, &tmQ, q_full, 0, head, q0, b);
        tma_load_4d(sQ + kChunk, &tmQ, q_full, 0, head, q0 + 128, b);
        for (int j = 0; j < nkv; ++j, ++t) {
          const int ks = t % KST, vs = t % VST;
          mbar_wait(k_empty(ks), ((uint32_t)(t / KST) & 1u) ^ 1u);
          mbar_expect_tx(k_full(ks), kChunk);
          tma_load_4d(sK + ks * kChunk, &tmK, k_full(ks), 0, head, j * 128, b);
          mbar_wait(v_empty(vs), ((uint32_t)(t / VST) & 1u) ^ 1u);
          mbar_expect_tx(v_full(vs), kChunk);
          tma_load_4d(sV + vs * kChunk, &tmV, v_full(vs), 0, head, j * 128, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 16 || warp == 17) {
    // ===== MMA issuer of query tile g = warp - 16 =====
    if (SETREG) setmaxnreg_dec<kCtlRegs>();
    if (elect_one()) {
      const int g = warp - 16;
      const uint32_t tS = tmem + 128u * g;
      const uint64_t dq = make_smem_desc_sw128(sQ + g * kChunk, 16, 1024);
      // (this thread lives on 32 registers: ring slot / phase / tile parity are carried incrementally instead of being
      //  derived from a tile counter, and the two instruction descriptors of Q K^T are built once)
      static_assert(KST == VST, "the K and V rings advance together");
      const uint32_t idesc_pv = make_idesc_bf16(128, DV, 0, 1);  // B = V is MN-major
      int ks = 0;                  // ring slot of the current tile
      uint32_t kph = 0, tpar = 0;  // its ring phase; parity of the CTA-wide tile counter
      bool any = false;            // a tile has been issued before (the S buffer / P chunks have a previous tenant)
      // Q K^T of the tile in ring slot `st` / phase `ph`: waits for K and for the softmax warps to have pulled the previous S
      auto issue_qk = [&](int st, uint32_t ph, uint32_t prev_par, bool last) {
        mbar_wait_lean(k_full(st), ph);
        if (any) mbar_wait_lean(s_empty(g), prev_par);
        fence_after_sync();
        const uint64_t dk = make_smem_desc_sw128(sK + st * kChunk, 16, 1024);
        const uint32_t idesc = make_idesc_bf16(128, last ? keys_in_tile(nkv - 1) : 128, 0, 0);
#pragma unroll
        for (int k = 0; k < KS; ++k) mma_f16_ss(tS, dq + 2 * k, dk + 2 * k, idesc, k != 0);
        mma_commit(s_full(g));
        mma_commit(k_empty(st));
        any = true;
      };
      const int my_items = (int)blockIdx.x < p.n_items ? (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
      for (int it = 0; it < my_items; ++it) {
        mbar_wait_lean(q_full, (uint32_t)it & 1u);
        issue_qk(ks, kph, tpar ^ 1u, nkv == 1);
        for (int j = 0; j < nkv; ++j) {
          const int ks1 = ks + 1 == KST ? 0 : ks + 1;
          const uint32_t kph1 = ks + 1 == KST ? kph ^ 1u : kph;
          if (j + 1 < nkv) issue_qk(ks1, kph1, tpar, j + 2 == nkv);
          else mma_commit(q_empty);  // arrives when every Q K^T of this item has completed
          mbar_wait_lean(v_ones(ks), kph);  // V tile landed and its ones column is in place
          if (j == 0 && it > 0) mbar_wait_lean(o_free(g), (uint32_t)(it - 1) & 1u);  // the previous item's merge has read O
          fence_after_sync();
          const int ksteps = j + 1 == nkv ? (keys_in_tile(nkv - 1) >> 4) : 8;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait_lean(p_full(g, h), tpar);  // this half of P_j is in shared memory
            fence_after_sync();
            const uint32_t tO = tmem + 256u + (uint32_t)(DV * (2 * g + h));
            const uint64_t dp = make_smem_desc_sw128(sP + (uint32_t)(2 * g + h) * kChunk, 16, 1024);
            for (int k = 4 * h; k < 4 * h + 4 && k < ksteps; ++k) {
              const uint64_t db = make_smem_desc_sw128(sV + ks * kChunk + (uint32_t)k * 2048u, kChunk, 1024);
              mma_f16_ss(tO, dp + 2 * (k & 3), db, idesc_pv, (j > 0) || (k > 4 * h));
            }
            mma_commit(o_done(g, h));
          }
          mma_commit(v_empty(ks));
          ks = ks1; kph = kph1; tpar ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp == 19) {
    // ===== ones column: V[:, d] = 1 in every landed V tile, so that P V also produces the softmax denominator
    // (column d of O = sum_k P[q,k]) on the tensor core instead of one FADD per exponential in the softmax warps
    if (SETREG) setmaxnreg_dec<kCtlRegs>();
    const uint32_t chunk = (uint32_t)(p.d * 2) >> 4, within = (uint32_t)(p.d * 2) & 15u;
    const int my_items = (int)blockIdx.x < p.n_items ? (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    for (int j = 0; j < my_items * nkv; ++j) {
      const int vs = j % VST;
      mbar_wait(v_full(vs), (uint32_t)(j / VST) & 1u);
      uint8_t* vt = gen + (sV - base) + (uint32_t)vs * kChunk;
#pragma unroll
      for (int r = lane; r < 128; r += 32)
        *reinterpret_cast<uint16_t*>(vt + r * 128 + ((chunk ^ (uint32_t)(r & 7)) << 4) + within) = 0x3F80;  // bf16 1.0
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(v_ones(vs));
    }
  } else if (warp < 16) {
    // ===== softmax: thread = (query row, key half) =====
    if (SETREG) setmaxnreg_inc<kSmxRegs>();
    const int g = (warp >> 2) & 1, h = warp >> 3;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + 128u * g + 64u * h + lane_off;
    const uint32_t tO = tmem + 256u + (uint32_t)(DV * (2 * g + h)) + lane_off;
    uint8_t* rowp = gen + (sP - base) + (uint32_t)(2 * g + h) * kChunk + row * 128;
    const int sw = row & 7;
    // 16-byte chunk c of this thread's P row sits at prow ^ (c << 4) (128-byte swizzle; bits 4-6 of the row start are 0)
    uint32_t prow = (sP + (uint32_t)(2 * g + h) * kChunk + (uint32_t)row * 128u) | ((uint32_t)sw << 4);
    // loop-invariant addresses, pinned: ptxas otherwise rebuilds each of them from %tid and the shared window every tile.
    // s_empty(g) = s_full(g) + 16 and o_done(g, h) = p_full(g, h) + 32 by the barrier layout above.
    uint32_t b_sfull = s_full(g), b_pfull = p_full(g, h), tSp = tS;
    int last_valid = p.Nk - (nkv - 1) * 128 - 64 * h;  // valid keys of this half in the last tile (all others are full)
    last_valid = last_valid < 0 ? 0 : (last_valid > 64 ? 64 : last_valid);
    if (SETREG) {
      pin_reg(b_sfull, lane); pin_reg(b_pfull, lane); pin_reg(tSp, lane);
      pin_reg(prow, lane);
    }
    const uint32_t b_staken = b_sfull + 16u, b_odone = b_pfull + 32u;
    const int my_items = (int)blockIdx.x < p.n_items ? (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    int t = 0;
    for (int it = 0; it < my_items; ++it) {
    float m_run = -INFINITY;  // (the running denominator lives in column d of the accumulator)
    for (const int t_end = t + nkv; t < t_end; ++t) {
      const int nvalid = (t == t_end - 1) ? last_valid : 64;
      mbar_wait(b_sfull, (uint32_t)t & 1u);
      fence_after_sync();
      uint32_t sv[64];
      tmem_ld32_at<0>(tSp, sv);
      tmem_ld32_at<32>(tSp + 32, sv);
      tmem_ld_wait();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_staken);
      if (nvalid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= nvalid) sv[i] = 0xff800000u;  // -inf
      }
      float mx = fmax3(__uint_as_float(sv[0]), __uint_as_float(sv[1]), __uint_as_float(sv[2]));
      float mx2 = fmax3(__uint_as_float(sv[3]), __uint_as_float(sv[4]), __uint_as_float(sv[5]));
#pragma unroll
      for (int i = 6; i + 3 < 64; i += 4) {
        mx = fmax3(mx, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
        mx2 = fmax3(mx2, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
      }
      mx = fmax3(mx, mx2, fmaxf(__uint_as_float(sv[62]), __uint_as_float(sv[63])));
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const bool grow = (m_new - m_run) > 8.f;  // (-inf) - (-inf) = NaN -> false: nothing to move
      // P V of the CTA's previous tile (of this half) must be complete before O is touched or P overwritten
      bool waited = (t == 0);
      if (DEFER == 0 && !waited) {
        mbar_wait(b_odone, (uint32_t)(t - 1) & 1u);
        fence_after_sync();
        waited = true;
      }
      if (__any_sync(0xffffffffu, grow)) {
        if (!waited) {
          mbar_wait(b_odone, (uint32_t)(t - 1) & 1u);
          fence_after_sync();
          waited = true;
        }
        const float alpha = grow ? ex2f(m_run - m_new) : 1.f;
        if (grow) m_run = m_new;
        if (t + nkv != t_end) {  // not the item's first tile: O holds partial sums
#pragma unroll
          for (int c = 0; c < DV; c += 16) {
            uint32_t v[16];
            tmem_ld16(tO + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st16(tO + c, v);
          }
          tmem_st_wait();
        }
      }
      const float neg_m = (m_run == -INFINITY) ? 0.f : -m_run;  // no valid key yet: every term below becomes 2^-inf = 0
      // (masked keys were set to -inf above: 2^-inf = 0 on the MUFU path, 2^-126 on the FMA path — below anything bf16 keeps)
      const f32x2 scale2 = pack2(p.scale_log2, p.scale_log2), negm2 = pack2(neg_m, neg_m);
      auto exp8 = [&](int c) {
        float e[8];
        if (PP16 == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float xarg = fmaf(__uint_as_float(sv[c + i]), p.scale_log2, neg_m);
            e[i] = (i >= 6) ? ex2_poly(xarg) : ex2f(xarg);  // 2 of 8 on the FMA pipe
          }
        } else {
          // pairs on the FMA pipe in this block of 8 keys: PP16 / 2, the odd one going to the odd blocks
          const int npoly = (PP16 >> 1) + ((PP16 & 1) & (c >> 3));
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const f32x2 x2 = fma2(pack2(__uint_as_float(sv[c + 2 * j]), __uint_as_float(sv[c + 2 * j + 1])), scale2, negm2);
            if (j >= 4 - npoly) {
              ex2_poly2(x2, e[2 * j], e[2 * j + 1]);
            } else {
              float x0, x1;
              unpack2(x2, x0, x1);
              e[2 * j] = ex2f(x0);
              e[2 * j + 1] = ex2f(x1);
            }
          }
        }
        uint4 w;
        w.x = pack_bf16(e[0], e[1]); w.y = pack_bf16(e[2], e[3]);
        w.z = pack_bf16(e[4], e[5]); w.w = pack_bf16(e[6], e[7]);
        return w;
      };
      uint4 early[DEFER / 8 + 1];
#pragma unroll
      for (int c = 0; c < DEFER; c += 8) early[c >> 3] = exp8(c);
      if (!waited) {
        mbar_wait(b_odone, (uint32_t)(t - 1) & 1u);
        fence_after_sync();
      }
      if (SETREG) {
#pragma unroll
        for (int c = 0; c < DEFER; c += 8) st_shared_v4(prow ^ (uint32_t)(c << 1), early[c >> 3]);
#pragma unroll
        for (int c = DEFER; c < 64; c += 8) st_shared_v4(prow ^ (uint32_t)(c << 1), exp8(c));
      } else {
#pragma unroll
        for (int c = 0; c < DEFER; c += 8) *reinterpret_cast<uint4*>(rowp + (((c >> 3) ^ sw) << 4)) = early[c >> 3];
#pragma unroll
        for (int c = DEFER; c < 64; c += 8) *reinterpret_cast<uint4*>(rowp + (((c >> 3) ^ sw) << 4)) = exp8(c);
      }
      fence_proxy_async_smem();  // P (generic proxy) -> visible to the tensor core's async proxy
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_pfull);
    }
    // ---- merge the two halves of each row and write O / l as bf16 ----
    float2* xch = reinterpret_cast<float2*>(gen + xch_off) + (it & 1) * 256;
    if (h == 1) xch[g * 128 + row] = make_float2(m_run, 0.f);
    softmax_bar_sync();
    if (h == 0) {
      int q0, head, b;
      item_coords((int)blockIdx.x + it * (int)gridDim.x, q0, head, b);
      mbar_wait(o_done(g, 0), (uint32_t)(t - 1) & 1u);
      mbar_wait(o_done(g, 1), (uint32_t)(t - 1) & 1u);
      fence_after_sync();
      const float2 other = xch[g * 128 + row];
      const float m_all = fmaxf(m_run, other.x);
      const float fa = (m_run == -INFINITY) ? 0.f : ex2f(m_run - m_all);
      const float fb = (other.x == -INFINITY) ? 0.f : ex2f(other.x - m_all);
      const uint32_t tOb = tO + (uint32_t)DV;  // accumulator of the other half (same TMEM lanes)
      float la, lb;  // denominators: column d of each accumulator
      {
        uint32_t va[16], vb[16];
        const int cl = p.d & ~15;
        tmem_ld16(tO + cl, va);
        tmem_ld16(tOb + cl, vb);
        tmem_ld_wait();
        la = 0.f; lb = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i == (p.d & 15)) { la = __uint_as_float(va[i]); lb = __uint_as_float(vb[i]); }
      }
      const float inv_l = 1.f / ((fa != 0.f ? la * fa : 0.f) + (fb != 0.f ? lb * fb : 0.f));
      const float ca = fa * inv_l, cb = fb * inv_l;
      const int q = q0 + g * 128 + row;
      const bool ok = q < p.Nq;
      bf16* orow = p.out + ((long long)b * p.Nq + q) * p.ldo + head * p.d;
#pragma unroll
      for (int c = 0; c < DV; c += 16) {
        uint32_t va[16], vb[16];
        tmem_ld16(tO + c, va);
        tmem_ld16(tOb + c, vb);
        tmem_ld_wait();
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float xa = fa != 0.f ? __uint_as_float(va[i]) : 0.f;  // an accumulator that never saw a key is uninitialised
          const float xb = fb != 0.f ? __uint_as_float(vb[i]) : 0.f;
          o[i] = xa * ca + xb * cb;
        }
        if (ok) {
          uint4 w0, w1;
          w0.x = pack_bf16(o[0], o[1]);   w0.y = pack_bf16(o[2], o[3]);
          w0.z = pack_bf16(o[4], o[5]);   w0.w = pack_bf16(o[6], o[7]);
          w1.x = pack_bf16(o[8], o[9]);   w1.y = pack_bf16(o[10], o[11]);
          w1.z = pack_bf16(o[12], o[13]); w1.w = pack_bf16(o[14], o[15]);
          if (c + 8 <= p.d) *reinterpret_cast<uint4*>(orow + c) = w0;
          if (c + 16 <= p.d) *reinterpret_cast<uint4*>(orow + c + 8) = w1;
        }
        __syncwarp();
      }
      // both accumulators of this row are in registers / on their way to HBM: the next item's first P V may overwrite them
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free(g));
    }
    }  // items
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, kTmemCols);
}

constexpr size_t attn2h_smem_bytes() { return 1024 + (size_t)(2 + 3 + 3 + 4) * 128 * 128 + 4096 + 8 + 8 * 12 + 96 + 8 * 3 + 24 + 16 + 16; }

// ------------------------------------------------------------------------------------------------------
// attn2x: the half-row design of attn2h for d = 80 (the 32x32 level).  Four accumulators of 96 columns (80 + the ones
// column, rounded to the MMA granularity) leave 128 TMEM columns for S, so the two query tiles SHARE one S buffer: a
// thread pulls its 64 scores into registers as soon as they exist, tile g's Q K^T is issued when the other tile's
// threads have pulled theirs, and the two tiles settle half a key tile apart.  K / V rows are two 64-column chunks;
// K is double-buffered, V single (P V_j needs V_j a softmax pass after Q K^T_j needed K_j), each ring with its own
// producer so that K_{j+1} is not held back behind the wait for V_j's slot.
// Warps 0-15 softmax (quad = w&3, query tile g = (w>>2)&1, half h = w>>3), 16/17 MMA issue of tile 0/1, 18 TMA of Q and K,
// 19 TMA of V + the ones column.
// ------------------------------------------------------------------------------------------------------
template <int KS, int DV, int PP16, int KST, int VST>
__global__ void __launch_bounds__(kAHThreads, 1)
attn2x_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using namespace tc05;
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t kChunk = 128 * 128;  // [128 rows][64 bf16] swizzled = 16 KB
  constexpr int DCH = 2;
  constexpr uint32_t kTmemCols = 512;
  static_assert(128 + 4 * DV <= 512 && DV % 16 == 0 && DV <= 128, "S (128 columns) and four accumulators must fit TMEM");
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;                          // 2 tiles x DCH chunks
  const uint32_t sK = sQ + 2 * DCH * kChunk;         // KST x DCH chunks
  const uint32_t sV = sK + KST * DCH * kChunk;       // VST x DCH chunks
  const uint32_t sP = sV + VST * DCH * kChunk;       // chunk 2g+h = P of query tile g, key half h
  const uint32_t xch_off = (sP - base) + 4 * kChunk;  // float [2][128]: running maxima of the h = 1 threads for the final merge
  const uint32_t bars = base + xch_off + 1024u;
  const uint32_t q_full = bars;
  auto k_full = [&](int s) { return bars + 8u + 8u * s; };
  auto k_empty = [&](int s) { return bars + 24u + 8u * s; };
  auto v_full = [&](int s) { return bars + 40u + 8u * s; };
  auto v_empty = [&](int s) { return bars + 56u + 8u * s; };
  auto v_ones = [&](int s) { return bars + 72u + 8u * s; };
  const uint32_t gbars = bars + 88u;
  static_assert(KST <= 2 && VST <= 2, "two barrier slots per ring");
  auto s_full = [&](int g) { return gbars + 8u * g; };
  auto s_taken = [&](int g) { return gbars + 16u + 8u * g; };
  auto p_full = [&](int g, int h) { return gbars + 32u + 8u * (2 * g + h); };
  auto o_done = [&](int g, int h) { return gbars + 64u + 8u * (2 * g + h); };
  const uint32_t q_empty = gbars + 96u;  // every Q K^T of the item has completed: Q may be overwritten
  auto o_free = [&](int g) { return gbars + 104u + 8u * g; };  // the merge has read tile g's accumulators
  const uint32_t tmem_slot = gbars + 120u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkv = (p.Nk + 127) / 128;
  // persistent CTAs as in attn2h: items = (batch, head, 256-query tile), barrier phases on the CTA-wide tile counter
  auto item_coords = [&](int item, int& q0, int& head, int& b) {
    const int qt = item % p.n_qt, hb = item / p.n_qt;
    q0 = qt * 256; head = hb % p.heads; b = hb / p.heads;
  };
  const int my_items = (int)blockIdx.x < p.n_items ? (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  pdl_trigger();
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 2);  // both tiles' MMA warps
    for (int g = 0; g < 2; ++g) mbar_init(o_free(g), 4);  // the four h = 0 warps of the tile
    for (int s = 0; s < KST; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 2); }  // released by both tiles' MMA warps
    for (int s = 0; s < VST; ++s) { mbar_init(v_full(s), 1); mbar_init(v_empty(s), 2); mbar_init(v_ones(s), 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(s_taken(g), 8);  // one arrive per softmax warp of the tile (both halves)
      for (int h = 0; h < 2; ++h) { mbar_init(p_full(g, h), 4); mbar_init(o_done(g, h), 1); }
    }
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_wait();  // Q / K / V come from the projection GEMMs

  auto keys_in_tile = [&](int j) {  // valid keys of tile j rounded up to the MMA granularity
    int n = p.Nk - j * 128;
    n = n > 128 ? 128 : n;
    return (n + 15) & ~15;
  };

  if (warp == 18) {
    // ===== TMA producer of Q and K =====
    setmaxnreg_dec<kCtlRegs>();
    if (elect_one()) {
      int t = 0;
      for (int it = 0; it < my_items; ++it) {
        int q0, head, b;
        item_coords((int)blockIdx.x + it * (int)gridDim.x, q0, head, b);
        if (it > 0) mbar_wait(q_empty, (uint32_t)(it - 1) & 1u);
        mbar_expect_tx(q_full, 2 * DCH * kChunk);
#pragma unroll
        for (int c = 0; c < DCH; ++c) {
          tma_load_4d(sQ + c * kChunk, &tmQ, q_full, 64 * c, head, q0, b);
          tma_load_4d(sQ + (DCH + c) * kChunk, &tmQ, q_full, 64 * c, head, q0 + 128, b);
        }
        for (int j = 0; j < nkv; ++j, ++t) {
          const int ks = t % KST;
          mbar_wait(k_empty(ks), ((uint32_t)(t / KST) & 1u) ^ 1u);
          mbar_expect_tx(k_full(ks), DCH * kChunk);
#pragma unroll
          for (int c = 0; c < DCH; ++c) tma_load_4d(sK + (ks * DCH + c) * kChunk, &tmK, k_full(ks), 64 * c, head, j * 128, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 16 || warp == 17) {
    // ===== MMA issuer of query tile g = warp - 16 =====
    setmaxnreg_dec<kCtlRegs>();
    if (elect_one()) {
      const int g = warp - 16;
      const uint32_t tS = tmem;  // shared by the two tiles
      const uint64_t dq = make_smem_desc_sw128(sQ + g * DCH * kChunk, 16, 1024);
      // tile t of the CTA = key tile j of the current item (this thread lives on 32 registers: unbounded waits, see attn2h)
      auto issue_qk = [&](int t, int j) {
        const int st = t % KST;
        mbar_wait_lean(k_full(st), (uint32_t)(t / KST) & 1u);
        // the S buffer is free once the OTHER tile's threads have pulled their scores: tile 1's previous S before tile 0's
        // next one, tile 0's S_t before tile 1's S_t
        if (g == 0) { if (t > 0) mbar_wait_lean(s_taken(1), (uint32_t)(t - 1) & 1u); }
        else mbar_wait_lean(s_taken(0), (uint32_t)t & 1u);
        fence_after_sync();
        const uint32_t idesc = make_idesc_bf16(128, keys_in_tile(j), 0, 0);
        const uint64_t dk = make_smem_desc_sw128(sK + st * DCH * kChunk, 16, 1024);
#pragma unroll
        for (int k = 0; k < KS; ++k) {  // k-step k: chunk k / 4, 32 bytes per step inside the 128-byte swizzled row
          const uint64_t off = (uint64_t)((k >> 2) * (kChunk >> 4) + (k & 3) * 2);
          mma_f16_ss(tS, dq + off, dk + off, idesc, k != 0);
        }
        mma_commit(s_full(g));
        mma_commit(k_empty(st));
      };
      const uint32_t idesc_pv = make_idesc_bf16(128, DV, 0, 1);  // B = V is MN-major
      int t = 0;
      for (int it = 0; it < my_items; ++it) {
        mbar_wait_lean(q_full, (uint32_t)it & 1u);
        issue_qk(t, 0);
        for (int j = 0; j < nkv; ++j, ++t) {
          if (j + 1 < nkv) issue_qk(t + 1, j + 1);
          else mma_commit(q_empty);  // arrives when every Q K^T of this item has completed
          const int vs = t % VST;
          mbar_wait_lean(v_ones(vs), (uint32_t)(t / VST) & 1u);  // V tile landed and its ones column is in place
          if (j == 0 && it > 0) mbar_wait_lean(o_free(g), (uint32_t)(it - 1) & 1u);  // the previous item's merge has read O
          fence_after_sync();
          const int ksteps = keys_in_tile(j) >> 4;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait_lean(p_full(g, h), (uint32_t)t & 1u);  // this half of P_j is in shared memory
            fence_after_sync();
            const uint32_t tO = tmem + 128u + (uint32_t)(DV * (2 * g + h));
            const uint64_t dp = make_smem_desc_sw128(sP + (uint32_t)(2 * g + h) * kChunk, 16, 1024);
            for (int k = 4 * h; k < 4 * h + 4 && k < ksteps; ++k) {
              const uint64_t db = make_smem_desc_sw128(sV + (uint32_t)(vs * DCH) * kChunk + (uint32_t)k * 2048u, kChunk, 1024);
              mma_f16_ss(tO, dp + 2 * (k & 3), db, idesc_pv, (j > 0) || (k > 4 * h));
            }
            mma_commit(o_done(g, h));
          }
          mma_commit(v_empty(vs));
        }
      }
    }
    __syncwarp();
  } else if (warp == 19) {
    // ===== V producer + ones column: V[:, d] = 1 in every landed V tile, so that P V also produces the softmax
    // denominator (column d of O = sum_k P[q,k]) on the tensor core =====
    setmaxnreg_dec<kCtlRegs>();
    const uint32_t c64 = (uint32_t)p.d >> 6, chunk = ((uint32_t)(p.d & 63) * 2) >> 4, within = (uint32_t)(p.d * 2) & 15u;
    const int n_tiles = my_items * nkv;
    auto load_v = [&](int t) {  // lane 0: V tile t of the CTA into its ring slot, once P V of the slot's previous tenant has completed
      const int slot = t % VST;
      int q0, head, b;
      item_coords((int)blockIdx.x + (t / nkv) * (int)gridDim.x, q0, head, b);
      mbar_wait(v_empty(slot), ((uint32_t)(t / VST) & 1u) ^ 1u);
      mbar_expect_tx(v_full(slot), DCH * kChunk);
#pragma unroll
      for (int c = 0; c < DCH; ++c) tma_load_4d(sV + (slot * DCH + c) * kChunk, &tmV, v_full(slot), 64 * c, head, (t % nkv) * 128, b);
    };
    if (lane == 0)
      for (int t = 0; t < VST - 1 && t < n_tiles; ++t) load_v(t);
    for (int j = 0; j < n_tiles; ++j) {
      if (VST == 1 && lane == 0) load_v(j);
      __syncwarp();
      const int vs = j % VST;
      mbar_wait(v_full(vs), (uint32_t)(j / VST) & 1u);
      uint8_t* vt = gen + (sV - base) + (uint32_t)(vs * DCH + c64) * kChunk;
#pragma unroll
      for (int r = lane; r < 128; r += 32)
        *reinterpret_cast<uint16_t*>(vt + r * 128 + ((chunk ^ (uint32_t)(r & 7)) << 4) + within) = 0x3F80;  // bf16 1.0
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(v_ones(vs));
        if (VST > 1 && j + VST - 1 < n_tiles) load_v(j + VST - 1);  // prefetch after the patch: the patch never waits for a P V
      }
      __syncwarp();
    }
  } else if (warp < 16) {
    // ===== softmax: thread = (query row, key half) =====
    setmaxnreg_inc<kSmxRegs>();
    const int g = (warp >> 2) & 1, h = warp >> 3;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + 64u * h + lane_off;
    const uint32_t tO = tmem + 128u + (uint32_t)(DV * (2 * g + h)) + lane_off;
    const int sw = row & 7;
    // 16-byte chunk c of this thread's P row sits at prow ^ (c << 4) (128-byte swizzle; bits 4-6 of the row start are 0)
    uint32_t prow = (sP + (uint32_t)(2 * g + h) * kChunk + (uint32_t)row * 128u) | ((uint32_t)sw << 4);
    // loop-invariant addresses, pinned: ptxas otherwise rebuilds each of them from %tid and the shared window every tile
    uint32_t b_sfull = s_full(g), b_pfull = p_full(g, h), tSp = tS;  // s_taken(g) = s_full(g) + 16, o_done(g, h) = p_full(g, h) + 32
    int last_valid = p.Nk - (nkv - 1) * 128 - 64 * h;  // valid keys of this half in the last tile (all others are full)
    last_valid = last_valid < 0 ? 0 : (last_valid > 64 ? 64 : last_valid);
    pin_reg(b_sfull, lane); pin_reg(b_pfull, lane); pin_reg(tSp, lane);
    pin_reg(prow, lane);
    const uint32_t b_staken = b_sfull + 16u, b_odone = b_pfull + 32u;
    int t = 0;
    for (int it = 0; it < my_items; ++it) {
    float m_run = -INFINITY;  // (the running denominator lives in column d of the accumulator)
    for (const int t_end = t + nkv; t < t_end; ++t) {
      const int nvalid = (t == t_end - 1) ? last_valid : 64;
      mbar_wait(b_sfull, (uint32_t)t & 1u);
      fence_after_sync();
      uint32_t sv[64];
      tmem_ld32_at<0>(tSp, sv);
      tmem_ld32_at<32>(tSp + 32, sv);
      tmem_ld_wait();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_staken);
      if (nvalid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= nvalid) sv[i] = 0xff800000u;  // -inf
      }
      float mx = fmax3(__uint_as_float(sv[0]), __uint_as_float(sv[1]), __uint_as_float(sv[2]));
      float mx2 = fmax3(__uint_as_float(sv[3]), __uint_as_float(sv[4]), __uint_as_float(sv[5]));
#pragma unroll
      for (int i = 6; i + 3 < 64; i += 4) {
        mx = fmax3(mx, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
        mx2 = fmax3(mx2, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
      }
      mx = fmax3(mx, mx2, fmaxf(__uint_as_float(sv[62]), __uint_as_float(sv[63])));
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const bool grow = (m_new - m_run) > 8.f;  // (-inf) - (-inf) = NaN -> false: nothing to move
      bool waited = (t == 0);  // P V of the CTA's previous tile (this half) must be complete before O is touched or P overwritten
      if (__any_sync(0xffffffffu, grow)) {
        if (!waited) {
          mbar_wait(b_odone, (uint32_t)(t - 1) & 1u);
          fence_after_sync();
          waited = true;
        }
        const float alpha = grow ? ex2f(m_run - m_new) : 1.f;
        if (grow) m_run = m_new;
        if (t + nkv != t_end) {  // not the item's first tile: O holds partial sums
#pragma unroll
          for (int c = 0; c < DV; c += 16) {
            uint32_t v[16];
            tmem_ld16(tO + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st16(tO + c, v);
          }
          tmem_st_wait();
        }
      }
      const float neg_m = (m_run == -INFINITY) ? 0.f : -m_run;  // no valid key yet: every term below becomes 2^-inf = 0
      const f32x2 scale2 = pack2(p.scale_log2, p.scale_log2), negm2 = pack2(neg_m, neg_m);
      auto exp8 = [&](int c) {
        float e[8];
        // pairs on the FMA pipe in this block of 8 keys: PP16 / 2, the odd one going to the odd blocks
        const int npoly = (PP16 >> 1) + ((PP16 & 1) & (c >> 3));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const f32x2 x2 = fma2(pack2(__uint_as_float(sv[c + 2 * j]), __uint_as_float(sv[c + 2 * j + 1])), scale2, negm2);
          if (j >= 4 - npoly) {
            ex2_poly2(x2, e[2 * j], e[2 * j + 1]);
          } else {
            float x0, x1;
            unpack2(x2, x0, x1);
            e[2 * j] = ex2f(x0);
            e[2 * j + 1] = ex2f(x1);
          }
        }
        uint4 w;
        w.x = pack_bf16(e[0], e[1]); w.y = pack_bf16(e[2], e[3]);
        w.z = pack_bf16(e[4], e[5]); w.w = pack_bf16(e[6], e[7]);
        return w;
      };
      uint4 early[2];
      early[0] = exp8(0);
      early[1] = exp8(8);
      if (!waited) {
        mbar_wait(b_odone, (uint32_t)(t - 1) & 1u);
        fence_after_sync();
      }
      st_shared_v4(prow, early[0]);
      st_shared_v4(prow ^ 16u, early[1]);
#pragma unroll
      for (int c = 16; c < 64; c += 8) st_shared_v4(prow ^ (uint32_t)(c << 1), exp8(c));
      fence_proxy_async_smem();  // P (generic proxy) -> visible to the tensor core's async proxy
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_pfull);
    }
    // ---- merge the two halves of each row and write O / l as bf16 ----
    // (one exchange buffer is enough: an h = 1 thread can finish the NEXT item only after its h = 0 partner has pulled that
    //  item's first scores — the shared S buffer orders them — i.e. after the partner has read this item's entry; nkv >= 2)
    float* xch = reinterpret_cast<float*>(gen + xch_off) + g * 128;
    if (h == 1) xch[row] = m_run;
    softmax_bar_sync();
    if (h == 0) {
      int q0, head, b;
      item_coords((int)blockIdx.x + it * (int)gridDim.x, q0, head, b);
      mbar_wait(o_done(g, 0), (uint32_t)(t - 1) & 1u);
      mbar_wait(o_done(g, 1), (uint32_t)(t - 1) & 1u);
      fence_after_sync();
      const float m_other = xch[row];
      const float m_all = fmaxf(m_run, m_other);
      const float fa = (m_run == -INFINITY) ? 0.f : ex2f(m_run - m_all);
      const float fb = (m_other == -INFINITY) ? 0.f : ex2f(m_other - m_all);
      const uint32_t tOb = tO + (uint32_t)DV;  // accumulator of the other half (same TMEM lanes)
      float la, lb;  // denominators: column d of each accumulator
      {
        uint32_t va[16], vb[16];
        const int cl = p.d & ~15;
        tmem_ld16(tO + cl, va);
        tmem_ld16(tOb + cl, vb);
        tmem_ld_wait();
        la = 0.f; lb = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i == (p.d & 15)) { la = __uint_as_float(va[i]); lb = __uint_as_float(vb[i]); }
      }
      const float inv_l = 1.f / ((fa != 0.f ? la * fa : 0.f) + (fb != 0.f ? lb * fb : 0.f));
      const float ca = fa * inv_l, cb = fb * inv_l;
      const int q = q0 + g * 128 + row;
      const bool ok = q < p.Nq;
      bf16* orow = p.out + ((long long)b * p.Nq + q) * p.ldo + head * p.d;
#pragma unroll
      for (int c = 0; c < DV; c += 16) {
        uint32_t va[16], vb[16];
        tmem_ld16(tO + c, va);
        tmem_ld16(tOb + c, vb);
        tmem_ld_wait();
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float xa = fa != 0.f ? __uint_as_float(va[i]) : 0.f;  // an accumulator that never saw a key is uninitialised
          const float xb = fb != 0.f ? __uint_as_float(vb[i]) : 0.f;
          o[i] = xa * ca + xb * cb;
        }
        if (ok) {
          uint4 w0, w1;
          w0.x = pack_bf16(o[0], o[1]);   w0.y = pack_bf16(o[2], o[3]);
          w0.z = pack_bf16(o[4], o[5]);   w0.w = pack_bf16(o[6], o[7]);
          w1.x = pack_bf16(o[8], o[9]);   w1.y = pack_bf16(o[10], o[11]);
          w1.z = pack_bf16(o[12], o[13]); w1.w = pack_bf16(o[14], o[15]);
          if (c + 8 <= p.d) *reinterpret_cast<uint4*>(orow + c) = w0;
          if (c + 16 <= p.d) *reinterpret_cast<uint4*>(orow + c + 8) = w1;
        }
        __syncwarp();
      }
      // both accumulators of this row have been read: the next item's first P V may overwrite them
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free(g));
    }
    }  // items
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, kTmemCols);
}

template <int KST, int VST>
constexpr size_t attn2x_smem_bytes() { return 1024 + (size_t)(4 + 2 * KST + 2 * VST + 4) * 128 * 128 + 1024 + 88 + 96 + 24 + 16 + 16; }



template <int DCH, int KST, int VST>
constexpr size_t attn2q_smem_bytes() {
  return 1024 + (size_t)(2 * DCH + KST * DCH + VST * DCH + 4) * 128 * 128 + 8 + 8 * (2 * KST + 2 * VST) + 64 + 16;
}

template <int DCH, int KST, int VST>
constexpr size_t attn_smem_bytes() {
  return 1024 + (size_t)(DCH + KST * DCH + VST * DCH + 2) * 128 * 128 + 24 + 8 * (KST + VST) + 16;
}

// ------------------------------------------------------------------------------------------------------
// xattn: cross-attention with a short context (Nk <= 128 keys: the 77-token prompt).  The generic kernels spend a
// whole CTA (TMEM allocation, barrier set-up, three TMA round trips) on ONE key tile per query tile — 2048 CTAs of
// ~6 us of pure latency at the 64x64 level (97 us per launch for 84 MB of traffic).  Here one CTA owns a (sample,
// head): K and V land in shared memory once, and the head's query tiles stream through a warp-specialised pipeline:
//   warp 9      TMA: K, V, then a QST-deep ring of Q tiles
//   warp 8      MMA: S_t = Q_t K^T into TMEM buffer t&1, then O_{t-1} = P_{t-1} V into accumulator (t-1)&1
//   warps 0-7   two groups of four; group g owns tiles t = g, g+2, ... and S / P / O buffer g: softmax of tile t (a
//               thread owns a row; one pass — all keys are in the tile: no running max, no rescale), then, once P V_t
//               has landed, O / l -> bf16 -> global; the other group's exponentials run meanwhile
// ------------------------------------------------------------------------------------------------------
static constexpr int kXAThreads = 320;

// NC = key columns the softmax code is compiled for (80: the 77-token prompt, 128: anything up to a full tile) — a
// compile-time bound keeps the softmax one straight-line block the compiler can software-pipeline (one warp per
// scheduler runs it: with warp-uniform branches around every 16-key block it ran at 3800 cycles per tile, 6x the MUFU floor)
template <int DCH, int KS, int DV, int QST, int NC>
__global__ void __launch_bounds__(kXAThreads, 1)
xattn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
             const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using namespace tc05;
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t kChunk = 128 * 128;
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kOStride = 128;  // accumulator b at TMEM column 256 + 128 b
  static_assert(DV <= 128, "two S buffers (2 x 128 columns) and two accumulators (2 x 128) fill the 512 TMEM columns");
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = base;                        // DCH chunks
  const uint32_t sV = sK + DCH * kChunk;           // DCH chunks
  const uint32_t sQ = sV + DCH * kChunk;           // QST x DCH chunks
  const uint32_t sP = sQ + QST * DCH * kChunk;     // 2 buffers x 2 chunks
  const uint32_t sL = sP + 4 * kChunk;             // float [2][128]: row sums
  const uint32_t bars = sL + 1024u;
  const uint32_t kv_full = bars;
  auto q_full = [&](int s) { return bars + 8u + 8u * s; };
  auto q_empty = [&](int s) { return bars + 8u + 8u * (QST + s); };
  const uint32_t bb = bars + 8u + 16u * QST;
  auto s_full = [&](int b) { return bb + 8u * b; };
  auto s_free = [&](int b) { return bb + 16u + 8u * b; };
  auto p_full = [&](int b) { return bb + 32u + 8u * b; };
  auto o_full = [&](int b) { return bb + 48u + 8u * b; };
  auto o_free = [&](int b) { return bb + 64u + 8u * b; };
  const uint32_t tmem_slot = bb + 80u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x, b = blockIdx.y;
  // blockIdx.z splits the head's query tiles over several CTAs when (heads x batch) alone would leave most SMs idle
  // (UNet batch 2: 16 CTAs); each CTA loads K / V for itself (a few KB)
  const int nqt_all = (p.Nq + 127) / 128;
  const int per_cta = (nqt_all + (int)gridDim.z - 1) / (int)gridDim.z;
  const int t0 = (int)blockIdx.z * per_cta;
  const int nqt = max(0, min(nqt_all, t0 + per_cta) - t0);
  const int ncols = (p.Nk + 15) & ~15;  // keys rounded up to the MMA granularity (TMA zero-fills the missing rows)

  pdl_trigger();
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(kv_full, 1);
    for (int s = 0; s < QST; ++s) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full(i), 1); mbar_init(s_free(i), 4);
      mbar_init(p_full(i), 4); mbar_init(o_full(i), 1); mbar_init(o_free(i), 4);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_wait();  // Q comes from the projection GEMM of this step

  if (warp == 9) {
    // ===== TMA producer =====
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * DCH * kChunk);
#pragma unroll
      for (int c = 0; c < DCH; ++c) {
        tma_load_4d(sK + c * kChunk, &tmK, kv_full, 64 * c, head, 0, b);
        tma_load_4d(sV + c * kChunk, &tmV, kv_full, 64 * c, head, 0, b);
      }
      for (int t = 0; t < nqt; ++t) {
        const int st = t % QST;
        if (t >= QST) mbar_wait(q_empty(st), (uint32_t)(t / QST - 1) & 1u);
        mbar_expect_tx(q_full(st), DCH * kChunk);
#pragma unroll
        for (int c = 0; c < DCH; ++c) tma_load_4d(sQ + (st * DCH + c) * kChunk, &tmQ, q_full(st), 64 * c, head, (t0 + t) * 128, b);
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc_qk = make_idesc_bf16(128, ncols, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, DV, 0, 1);  // B = V is MN-major
      const uint64_t dk = make_smem_desc_sw128(sK, 16, 1024);
      const int ksteps = ncols >> 4;
      mbar_wait(kv_full, 0);
      for (int t = 0; t <= nqt; ++t) {
        if (t < nqt) {
          const int st = t % QST, sb = t & 1, u = t >> 1;
          mbar_wait(q_full(st), (uint32_t)(t / QST) & 1u);
          if (u >= 1) mbar_wait(s_free(sb), (uint32_t)(u - 1) & 1u);  // S of tile t-2 is in registers
          fence_after_sync();
          const uint64_t dq = make_smem_desc_sw128(sQ + st * DCH * kChunk, 16, 1024);
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            const uint64_t off = (uint64_t)((k >> 2) * (kChunk >> 4) + (k & 3) * 2);
            mma_f16_ss(tmem + 128u * sb, dq + off, dk + off, idesc_qk, k != 0);
          }
          mma_commit(s_full(sb));
          mma_commit(q_empty(st));
        }
        if (t >= 1) {
          const int tp = t - 1, pb = tp & 1, u = tp >> 1;
          mbar_wait(p_full(pb), (uint32_t)u & 1u);  // P of tile t-1 is in shared memory (and accumulator pb has been drained)
          fence_after_sync();
          const uint64_t dp = make_smem_desc_sw128(sP + (uint32_t)(2 * pb) * kChunk, 16, 1024);
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = dp + (uint64_t)((k >> 2) * (kChunk >> 4) + (k & 3) * 2);
            const uint64_t db = make_smem_desc_sw128(sV + (uint32_t)k * 2048u, kChunk, 1024);
            mma_f16_ss(tmem + 256u + kOStride * pb, da, db, idesc_pv, k != 0);
          }
          mma_commit(o_full(pb));
        }
      }
    }
    __syncwarp();
  } else {
    // ===== softmax + epilogue: two groups of four warps; group g owns the tiles t = g, g + 2, ... and with them S / P / O
    // buffer g.  A thread is a query row: softmax of tile t, then (once P V_t has landed) O / l -> bf16 -> global.  While one
    // group waits for its P V or stores its rows, the other group's exponentials keep the MUFU busy (one group alone ran
    // a tile in 2700 cycles against a MUFU floor of 640).
    const int g = warp >> 2, q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    const int sw = row & 7;
    const int sb = g;  // buffer index == group
    uint8_t* rowp = gen + (sP - base) + (uint32_t)(2 * sb) * kChunk + row * 128;
    for (int t = g, u = 0; t < nqt; t += 2, ++u) {
      mbar_wait(s_full(sb), (uint32_t)u & 1u);
      fence_after_sync();
      constexpr int NL = (NC + 31) / 32 * 32;  // columns loaded (whole 32-column TMEM loads)
      uint32_t sv[NL];
      tmem_ld32_at<0>(tmem + 128u * sb + lane_off, sv);
      if (NL > 32) tmem_ld32_at<(NL > 32 ? 32 : 0)>(tmem + 128u * sb + lane_off + 32, sv);
      if (NL > 64) tmem_ld32_at<(NL > 64 ? 64 : 0)>(tmem + 128u * sb + lane_off + 64, sv);
      if (NL > 96) tmem_ld32_at<(NL > 96 ? 96 : 0)>(tmem + 128u * sb + lane_off + 96, sv);
      tmem_ld_wait();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free(sb));
      // keys past the context (and columns the MMA never wrote) -> -inf: everything below is unpredicated straight-line code
#pragma unroll
      for (int i = 0; i < NC; ++i)
        if (i >= p.Nk) sv[i] = 0xff800000u;
      float mxa = -INFINITY, mxb = -INFINITY;  // two independent chains
#pragma unroll
      for (int c = 0; c < NC; c += 8) {
        mxa = fmax3(mxa, __uint_as_float(sv[c]), __uint_as_float(sv[c + 1]));
        mxb = fmax3(mxb, __uint_as_float(sv[c + 2]), __uint_as_float(sv[c + 3]));
        mxa = fmax3(mxa, __uint_as_float(sv[c + 4]), __uint_as_float(sv[c + 5]));
        mxb = fmax3(mxb, __uint_as_float(sv[c + 6]), __uint_as_float(sv[c + 7]));
      }
      const float scale = p.scale_log2;
      const float neg_m = -fmaxf(mxa, mxb) * scale;
      float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
      // (the P buffer and accumulator of this group are free: its previous tile's epilogue finished in program order)
#pragma unroll
      for (int c = 0; c < NC; c += 8) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = ex2f(fmaf(__uint_as_float(sv[c + i]), scale, neg_m));
        ls0 += e[0] + e[1]; ls1 += e[2] + e[3]; ls2 += e[4] + e[5]; ls3 += e[6] + e[7];
        uint4 w;
        w.x = pack_bf16(e[0], e[1]); w.y = pack_bf16(e[2], e[3]);
        w.z = pack_bf16(e[4], e[5]); w.w = pack_bf16(e[6], e[7]);
        const int chunk = c >> 6, un = (c & 63) >> 3;
        *reinterpret_cast<uint4*>(rowp + chunk * kChunk + ((un ^ sw) << 4)) = w;
      }
      const float inv_l = 1.f / ((ls0 + ls1) + (ls2 + ls3));
      fence_proxy_async_smem();  // P (generic proxy) -> visible to the tensor core's async proxy
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(sb));
      // ---- epilogue of the same tile ----
      mbar_wait(o_full(sb), (uint32_t)u & 1u);
      fence_after_sync();
      uint32_t o[DV];
#pragma unroll
      for (int c = 0; c < DV; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + 256u + kOStride * sb + lane_off + c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) o[c + i] = v[i];
      }
      tmem_ld_wait();
      fence_before_sync();
      const int q = (t0 + t) * 128 + row;
      if (q < p.Nq) {
        bf16* orow = p.out + ((long long)b * p.Nq + q) * p.ldo + head * p.d;
#pragma unroll
        for (int c = 0; c < DV; c += 8) {
          if (c + 8 <= p.d) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(o[c]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
            w.y = pack_bf16(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
            w.z = pack_bf16(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
            w.w = pack_bf16(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c) = w;
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, kTmemCols);
}

template <int DCH, int QST>
constexpr size_t xattn_smem_bytes() {
  return 1024 + (size_t)(2 * DCH + QST * DCH + 4) * 128 * 128 + 1024 + 8 + 16 * QST + 80 + 16;
}

// ------------------------------------------------------------------------------------------------------
// vattn: the VAE mid-block AttentionBlock (layers.py:28-59) as a flash kernel — ONE head of size 512, N = H*W tokens
// (4096 at 512x512, 9216 at 768x768), softmax(Q K^T / sqrt(512)) V + b_v.  The reference materialises the N x N
// score matrix (layers.py:46-50); round 1 of this engine did too (3 GEMMs + a row softmax per sample, 100 MB of fp32
// scores per sample at 512x512, 510 MB at 768x768).  Here nothing of size N x N exists:
//   * a CTA owns 128 query rows and ONE HALF (256 columns) of the output: TMEM holds two S buffers (2 x 128 columns)
//     and the 128 x 256 fp32 accumulator — exactly the 512 columns there are;
//   * Q (8 chunks of 64 columns, 128 KB) stays in shared memory; K and V stream through a 4-slot ring of 16 KB chunks in
//     the order the tensor core consumes them: K_0 | K_1 V_0 | K_2 V_1 | ... (K tile: 8 chunks, V half tile: 4 chunks);
//   * warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 softmax (a thread owns a query row) and epilogue.  S is double
//     buffered, so Q K_{j+1}^T runs while the softmax of tile j is computed; P V_j follows it on the tensor pipe.
// Tensor work per key tile: 2048 + 1024 cycles against ~1100 cycles of softmax: tensor / operand-delivery bound.
// ------------------------------------------------------------------------------------------------------
static constexpr int kVAThreads = 192;
static constexpr int kVARing = 4;

struct VAttnParams {
  int N;             // tokens per sample (queries == keys)
  float scale_log2;  // 512^-1/2 * log2(e)
  const float* v_bias;  // [512] added after the attention (rows of softmax sum to 1), may be null
  bf16* out;         // [B*N][ldo]
  long long ldo;
};

__global__ void __launch_bounds__(kVAThreads, 1)
vattn_kernel(const __grid_constant__ CUtensorMap tmQKV, const VAttnParams p) {
  using namespace tc05;
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t kChunk = 128 * 128;  // [128 rows][64 bf16] swizzled = 16 KB
  constexpr int D = 512, QCH = D / 64, VCH = 4;  // Q / K chunks per tile, V chunks per half tile
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;                       // 8 chunks
  const uint32_t sR = sQ + QCH * kChunk;          // ring: kVARing chunks
  const uint32_t sP = sR + kVARing * kChunk;      // 2 chunks (128 rows x 128 keys)
  const uint32_t bars = sP + 2 * kChunk;
  const uint32_t q_full = bars;
  auto r_full = [&](int s) { return bars + 8u + 8u * s; };
  auto r_empty = [&](int s) { return bars + 8u + 8u * (kVARing + s); };
  const uint32_t bb = bars + 8u + 16u * kVARing;
  auto s_full = [&](int i) { return bb + 8u * i; };
  auto s_free = [&](int i) { return bb + 16u + 8u * i; };
  const uint32_t p_full = bb + 32u, o_done = bb + 40u;
  const uint32_t tmem_slot = bb + 48u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, half = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.N + 127) / 128;
  auto keys_in_tile = [&](int j) {
    int n = p.N - j * 128;
    n = n > 128 ? 128 : n;
    return (n + 15) & ~15;
  };

  pdl_trigger();
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < kVARing; ++s) { mbar_init(r_full(s), 1); mbar_init(r_empty(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(s_full(i), 1); mbar_init(s_free(i), 4); }
    mbar_init(p_full, 4);
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_wait();  // q | k | v come from the projection GEMM

  if (warp == 0) {
    // ===== TMA producer: Q once, then the chunk stream K_0 | K_1 V_0 | ... | K_{n-1} V_{n-2} | V_{n-1} =====
    if (elect_one()) {
      mbar_expect_tx(q_full, QCH * kChunk);
#pragma unroll
      for (int c = 0; c < QCH; ++c) tma_load_3d(sQ + c * kChunk, &tmQKV, q_full, 64 * c, q0, b);
      uint32_t slot = 0, ph = 0;
      auto push = [&](int col, int row) {
        mbar_wait(r_empty(slot), ph ^ 1u);
        mbar_expect_tx(r_full(slot), kChunk);
        tma_load_3d(sR + slot * kChunk, &tmQKV, r_full(slot), col, row, b);
        if (++slot == kVARing) { slot = 0; ph ^= 1u; }
      };
      for (int j = 0; j <= nkv; ++j) {
        if (j < nkv)
          for (int c = 0; c < QCH; ++c) push(D + 64 * c, j * 128);                          // K_j
        if (j >= 1)
          for (int c = 0; c < VCH; ++c) push(2 * D + half * 256 + 64 * c, (j - 1) * 128);   // V_{j-1}, this CTA's half
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);  // B = V chunk is MN-major
      const uint64_t dp = make_smem_desc_sw128(sP, 16, 1024);
      uint32_t slot = 0, ph = 0;
      mbar_wait(q_full, 0);
      for (int j = 0; j <= nkv; ++j) {
        if (j < nkv) {
          const int sb = j & 1;
          if (j >= 2) mbar_wait(s_free(sb), (uint32_t)((j >> 1) - 1) & 1u);  // S of tile j-2 is in registers
          fence_after_sync();
          const uint32_t idesc = make_idesc_bf16(128, keys_in_tile(j), 0, 0);
          for (int c = 0; c < QCH; ++c) {
            mbar_wait(r_full(slot), ph);
            fence_after_sync();
            const uint64_t dq = make_smem_desc_sw128(sQ + c * kChunk, 16, 1024);
            const uint64_t dk = make_smem_desc_sw128(sR + slot * kChunk, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_f16_ss(tmem + 128u * sb, dq + 2 * k, dk + 2 * k, idesc, (c | k) != 0);
            mma_commit(r_empty(slot));
            if (++slot == kVARing) { slot = 0; ph ^= 1u; }
          }
          mma_commit(s_full(sb));
        }
        if (j >= 1) {
          const int jj = j - 1;
          mbar_wait(p_full, (uint32_t)jj & 1u);  // P_jj is in shared memory (and the accumulator has been rescaled if needed)
          fence_after_sync();
          const int ksteps = keys_in_tile(jj) >> 4;
          for (int c = 0; c < VCH; ++c) {
            mbar_wait(r_full(slot), ph);
            fence_after_sync();
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t da = dp + (uint64_t)((k >> 2) * (kChunk >> 4) + (k & 3) * 2);
              const uint64_t db = make_smem_desc_sw128(sR + slot * kChunk + (uint32_t)k * 2048u, kChunk, 1024);
              mma_f16_ss(tmem + 256u + 64u * c, da, db, idesc_pv, (jj | k) != 0);
            }
            mma_commit(r_empty(slot));
            if (++slot == kVARing) { slot = 0; ph ^= 1u; }
          }
          mma_commit(o_done);
        }
      }
    }
    __syncwarp();
  } else {
    // ===== softmax + epilogue: thread = query row =====
    const int quad = warp & 3;  // a warp reaches TMEM lanes [32 (warp % 4), +32): warps 2..5 own quadrants 2, 3, 0, 1
    const int row = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int sw = row & 7;
    uint8_t* rowp = gen + (sP - base) + row * 128;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const int sb = j & 1;
      const int nvalid = min(128, p.N - j * 128);
      mbar_wait(s_full(sb), (uint32_t)(j >> 1) & 1u);
      fence_after_sync();
      uint32_t sv[128];
      const uint32_t tS = tmem + 128u * sb + lane_off;
      tmem_ld32_at<0>(tS, sv);
      tmem_ld32_at<32>(tS + 32, sv);
      tmem_ld32_at<64>(tS + 64, sv);
      tmem_ld32_at<96>(tS + 96, sv);
      tmem_ld_wait();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free(sb));
      if (nvalid < 128) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= nvalid) sv[i] = 0xff800000u;  // -inf
      }
      float mx = fmax3(__uint_as_float(sv[0]), __uint_as_float(sv[1]), __uint_as_float(sv[2]));
      float mx2 = fmax3(__uint_as_float(sv[3]), __uint_as_float(sv[4]), __uint_as_float(sv[5]));
#pragma unroll
      for (int i = 6; i + 3 < 128; i += 4) {
        mx = fmax3(mx, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
        mx2 = fmax3(mx2, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
      }
      mx = fmax3(mx, mx2, fmaxf(__uint_as_float(sv[126]), __uint_as_float(sv[127])));
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const bool grow = (m_new - m_run) > 8.f;  // the running max only moves when it grows by more than 2^8 (see attn2q)
      if (j > 0) {  // P V_{j-1} complete: P may be overwritten, the accumulator may be rescaled
        mbar_wait(o_done, (uint32_t)(j - 1) & 1u);
        fence_after_sync();
      }
      if (__any_sync(0xffffffffu, grow)) {
        const float alpha = grow ? ex2f(m_run - m_new) : 1.f;
        if (grow) { m_run = m_new; l_run *= alpha; }
        if (j > 0) {
#pragma unroll 4
          for (int c = 0; c < 256; c += 16) {
            uint32_t v[16];
            tmem_ld16(tmem + 256u + lane_off + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st16(tmem + 256u + lane_off + c, v);
          }
          tmem_st_wait();
        }
      }
      const float neg_m = (m_run == -INFINITY) ? 0.f : -m_run;
      float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
      for (int c = 0; c < 128; c += 8) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = ex2f(fmaf(__uint_as_float(sv[c + i]), p.scale_log2, neg_m));
        ls0 += (e[0] + e[1]) + (e[2] + e[3]);
        ls1 += (e[4] + e[5]) + (e[6] + e[7]);
        uint4 w;
        w.x = pack_bf16(e[0], e[1]); w.y = pack_bf16(e[2], e[3]);
        w.z = pack_bf16(e[4], e[5]); w.w = pack_bf16(e[6], e[7]);
        const int chunk = c >> 6, un = (c & 63) >> 3;
        *reinterpret_cast<uint4*>(rowp + chunk * kChunk + ((un ^ sw) << 4)) = w;
      }
      l_run += ls0 + ls1;
      fence_proxy_async_smem();  // P (generic proxy) -> visible to the tensor core's async proxy
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: O / l + b_v -> bf16 ----
    mbar_wait(o_done, (uint32_t)(nkv - 1) & 1u);
    fence_after_sync();
    const float inv_l = 1.f / l_run;
    const int q = q0 + row;
    const bool ok = q < p.N;
    bf16* orow = p.out + ((long long)b * p.N + q) * p.ldo + half * 256;
    const float* vb = p.v_bias ? p.v_bias + half * 256 : nullptr;
#pragma unroll 2
    for (int c = 0; c < 256; c += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + 256u + lane_off + c, v);
      tmem_ld_wait();
      float o[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(v[i]) * inv_l + (vb ? __ldg(vb + c + i) : 0.f);
      if (ok) {
        uint4 w0, w1;
        w0.x = pack_bf16(o[0], o[1]);   w0.y = pack_bf16(o[2], o[3]);
        w0.z = pack_bf16(o[4], o[5]);   w0.w = pack_bf16(o[6], o[7]);
        w1.x = pack_bf16(o[8], o[9]);   w1.y = pack_bf16(o[10], o[11]);
        w1.z = pack_bf16(o[12], o[13]); w1.w = pack_bf16(o[14], o[15]);
        *reinterpret_cast<uint4*>(orow + c) = w0;
        *reinterpret_cast<uint4*>(orow + c + 8) = w1;
      }
      __syncwarp();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

constexpr size_t vattn_smem_bytes() { return 1024 + (size_t)(8 + kVARing + 2) * 128 * 128 + 8 + 16 * kVARing + 48 + 16 + 16; }

// qkv: [B][N][ld] bf16 with q at column 0, k at 512, v at 1024; out: [B][N][ldo] (512 columns)
inline void launch_vattn(cudaStream_t stream, const bf16* qkv, long long ld, int B, int N, const float* v_bias, bf16* out, long long ldo) {
  static_assert(vattn_smem_bytes() <= 232448, "vattn must fit 227 KB of shared memory");
  VAttnParams p;
  p.N = N;
  p.scale_log2 = (float)(1.4426950408889634 / sqrt(512.0));
  p.v_bias = v_bias;
  p.out = out; p.ldo = ldo;
  uint64_t dims[3] = {1536, (uint64_t)N, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)N * ld * 2};
  uint32_t box[3] = {64, 128, 1};
  uint32_t es[3] = {1, 1, 1};
  CUtensorMap tm = make_tmap_bf16(qkv, 3, dims, strides, box, es);
  dim3 grid((unsigned)ceil_div(N, 128), 2, (unsigned)B);
  launch_pdl(vattn_kernel, grid, dim3(kVAThreads), vattn_smem_bytes(), stream, 1, tm, p);
  SDTF_CUDA(cudaGetLastError());
}

// Q/K/V token matrices: [B][N][ld] bf16; head h of Q at q + h*dstride etc.
struct AttnArgs {
  const bf16 *q, *k, *v;
  long long ldq, ldk, ldv;
  int B, heads, Nq, Nk, d, dstride;
  bf16* out;
  long long ldo;
  bool legacy = false;  // force the one-tile-per-CTA kernel (A/B measurements)
};

// (column within head, head, token, batch); box = 64 columns x 1 head x 128 tokens: columns past d are zero-filled
inline CUtensorMap make_tok_tmap(const bf16* base, int d, int dstride, int heads, int N, int B, long long ld) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)N, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)dstride * 2, (uint64_t)ld * 2, (uint64_t)N * ld * 2};
  uint32_t box[4] = {64, 1, 128, 1};
  uint32_t es[4] = {1, 1, 1, 1};
  return make_tmap_bf16(base, 4, dims, strides, box, es);
}

template <int DCH, int KS, int DV, int KST, int VST>
inline void init_attn_t() {
  SDTF_CUDA(cudaFuncSetAttribute(attn_kernel<DCH, KS, DV, KST, VST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)attn_smem_bytes<DCH, KST, VST>()));
}
// called once per process before any launch (and before any stream capture)
inline void init_attn_kernels() {
  init_attn_t<1, 3, 48, 2, 2>();
  SDTF_CUDA(cudaFuncSetAttribute(attn2q_kernel<1, 3, 48, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)attn2q_smem_bytes<1, 3, 3>()));
  static_assert(attn2q_smem_bytes<2, 2, 1>() <= 232448, "d = 80 two-tile attention must fit 227 KB of shared memory");
  SDTF_CUDA(cudaFuncSetAttribute(attn2q_kernel<2, 5, 80, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)attn2q_smem_bytes<2, 2, 1>()));
  SDTF_CUDA(cudaFuncSetAttribute(attn2h_kernel<3, 48, 16, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn2h_smem_bytes()));
  SDTF_CUDA(cudaFuncSetAttribute(attn2h_kernel<3, 48, 16, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn2h_smem_bytes()));
  SDTF_CUDA(cudaFuncSetAttribute(attn2x_kernel<5, 96, 3, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn2x_smem_bytes<2, 1>()));
  SDTF_CUDA(cudaFuncSetAttribute(attn2x_kernel<5, 96, 3, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn2x_smem_bytes<1, 2>()));
  init_attn_t<2, 5, 80, 2, 2>();
  init_attn_t<3, 10, 160, 1, 1>();
  SDTF_CUDA(cudaFuncSetAttribute(vattn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vattn_smem_bytes()));
  static_assert(xattn_smem_bytes<2, 3>() <= 232448, "d = 80 cross-attention must fit 227 KB of shared memory");
  SDTF_CUDA(cudaFuncSetAttribute(xattn_kernel<1, 3, 48, 4, 80>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xattn_smem_bytes<1, 4>()));
  SDTF_CUDA(cudaFuncSetAttribute(xattn_kernel<1, 3, 48, 4, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xattn_smem_bytes<1, 4>()));
  SDTF_CUDA(cudaFuncSetAttribute(xattn_kernel<2, 5, 80, 3, 80>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xattn_smem_bytes<2, 3>()));
  SDTF_CUDA(cudaFuncSetAttribute(xattn_kernel<2, 5, 80, 3, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xattn_smem_bytes<2, 3>()));
}

template <int DCH, int KS, int DV, int KST, int VST>
inline void launch_attn_t(cudaStream_t stream, const AttnArgs& a, const AttnParams& p, const CUtensorMap& tq,
                          const CUtensorMap& tk, const CUtensorMap& tv) {
  auto kern = attn_kernel<DCH, KS, DV, KST, VST>;
  constexpr size_t smem = attn_smem_bytes<DCH, KST, VST>();
  dim3 grid((unsigned)ceil_div(a.Nq, 128), (unsigned)a.heads, (unsigned)a.B);
  launch_pdl(kern, grid, dim3(128), smem, stream, 1, tq, tk, tv, p);
  SDTF_CUDA(cudaGetLastError());
}

inline void launch_attn(cudaStream_t stream, const AttnArgs& a) {
  AttnParams p;
  p.Nq = a.Nq; p.Nk = a.Nk; p.d = a.d; p.dstride = a.dstride;
  p.scale_log2 = (float)(1.4426950408889634 / sqrt((double)a.d));
  p.out = a.out; p.ldo = a.ldo;
  static const int attn_debug = getenv("SDTF_ATTN_DEBUG") ? atoi(getenv("SDTF_ATTN_DEBUG")) : 0;
  p.debug = attn_debug;
  static const int attn_profile = getenv("SDTF_ATTN_PROFILE") ? atoi(getenv("SDTF_ATTN_PROFILE")) : 0;
  static long long* prof_buf = nullptr;
  p.prof = nullptr;
  if (attn_profile && a.d == 40 && !a.legacy) {
    if (!prof_buf) SDTF_CUDA(cudaMalloc((void**)&prof_buf, 16 * sizeof(long long)));
    SDTF_CUDA(cudaMemsetAsync(prof_buf, 0, 16 * sizeof(long long), stream));
    p.prof = prof_buf;
  }
  SDTF_CHECK(a.dstride % 8 == 0, "attention: heads must start on 16-byte boundaries");
  CUtensorMap tq = make_tok_tmap(a.q, a.d, a.dstride, a.heads, a.Nq, a.B, a.ldq);
  CUtensorMap tk = make_tok_tmap(a.k, a.d, a.dstride, a.heads, a.Nk, a.B, a.ldk);
  CUtensorMap tv = make_tok_tmap(a.v, a.d, a.dstride, a.heads, a.Nk, a.B, a.ldv);
  // short contexts (the 77-token prompt): one CTA per (sample, head), K / V resident, query tiles pipelined
  static const int use_xattn = getenv("SDTF_XATTN") ? atoi(getenv("SDTF_XATTN")) : 1;
  if (use_xattn && !a.legacy && a.Nk <= 128 && a.Nq >= 256 && (a.d == 40 || a.d == 80)) {
    const int nqt = ceil_div(a.Nq, 128), hb = a.heads * a.B;
    int zsplit = hb >= 96 ? 1 : (148 + hb - 1) / hb;  // fill the machine when (heads x batch) is small
    if (zsplit > nqt) zsplit = nqt;
    dim3 grid((unsigned)a.heads, (unsigned)a.B, (unsigned)zsplit);
    if (a.d == 40) {
      if (a.Nk <= 80) launch_pdl(xattn_kernel<1, 3, 48, 4, 80>, grid, dim3(kXAThreads), xattn_smem_bytes<1, 4>(), stream, 1, tq, tk, tv, p);
      else launch_pdl(xattn_kernel<1, 3, 48, 4, 128>, grid, dim3(kXAThreads), xattn_smem_bytes<1, 4>(), stream, 1, tq, tk, tv, p);
    } else {
      if (a.Nk <= 80) launch_pdl(xattn_kernel<2, 5, 80, 3, 80>, grid, dim3(kXAThreads), xattn_smem_bytes<2, 3>(), stream, 1, tq, tk, tv, p);
      else launch_pdl(xattn_kernel<2, 5, 80, 3, 128>, grid, dim3(kXAThreads), xattn_smem_bytes<2, 3>(), stream, 1, tq, tk, tv, p);
    }
    return;
  }
  if (a.d == 40) {
    if (a.legacy) {
      launch_attn_t<1, 3, 48, 2, 2>(stream, a, p, tq, tk, tv);
    } else {
      dim3 grid((unsigned)ceil_div(a.Nq, 256), (unsigned)a.heads, (unsigned)a.B);
      static const int use_2q = getenv("SDTF_ATTN_2Q") ? atoi(getenv("SDTF_ATTN_2Q")) : 0;  // A/B: previous full-row kernel
      if (!use_2q) {
        SDTF_CHECK(a.d < 48, "attn2h keeps the softmax denominator in accumulator column d: needs d < DV");
        // SDTF_ATTN_SETREG=0 (A/B): the kernel without the per-role register budgets, pinned addresses and direct P stores.
        // Tuning history of the two other knobs: 3 of every 8 exponential pairs on the FMA pipe (0.725 / 0.749 / 0.702 /
        // 0.754 ms for scalar code, 2, 3, 4 of 8; 0.679 / 0.617 / 0.639 / 0.703 for 2, 3, 4, 5 in the final kernel) and the first 16 keys' exponentials before the wait for P V_{j-1}
        // (0.625 / 0.617 / 0.645 / 0.682 ms for 8 / 16 / 24 / 32), profiles/r02_b_attn2h_packed_exp.log, r02_u_*.log.
        static const int setreg = getenv("SDTF_ATTN_SETREG") ? atoi(getenv("SDTF_ATTN_SETREG")) : 1;
        // SDTF_ATTN_PERSIST: 0 (A/B) = one item per CTA, N > 1 = at most N CTAs (sanitizer runs: several items per CTA at tiny sizes)
        static const int persist = getenv("SDTF_ATTN_PERSIST") ? atoi(getenv("SDTF_ATTN_PERSIST")) : 1;
        p.heads = a.heads; p.n_qt = ceil_div(a.Nq, 256); p.n_items = p.n_qt * a.heads * a.B;
        const int cap = persist > 1 ? persist : sm_count();
        grid = dim3((unsigned)(persist && p.n_items > cap ? cap : p.n_items));
        const size_t sm = attn2h_smem_bytes();
        if (setreg) launch_pdl(attn2h_kernel<3, 48, 16, 3, true>, grid, dim3(kAHThreads), sm, stream, 1, tq, tk, tv, p);
        else launch_pdl(attn2h_kernel<3, 48, 16, 3>, grid, dim3(kAHThreads), sm, stream, 1, tq, tk, tv, p);
        SDTF_CUDA(cudaGetLastError());
        return;
      }
      attn2q_kernel<1, 3, 48, 3, 3><<<grid, kA2Threads, attn2q_smem_bytes<1, 3, 3>(), stream>>>(tq, tk, tv, p);
      SDTF_CUDA(cudaGetLastError());
      if (p.prof) {  // debug: per-role wait cycles, averaged per CTA
        SDTF_CUDA(cudaStreamSynchronize(stream));
        long long h[16];
        SDTF_CUDA(cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost));
        const double n = (double)grid.x * grid.y * grid.z;
        fprintf(stderr,
                "[attn2q-prof] per CTA: tma wait k_empty %.0f v_empty %.0f / %.0f | mma(g0) wait k_full %.0f s_empty %.0f v_full %.0f p_full "
                "%.0f / %.0f | softmax(w0) wait s_full %.0f o_done %.0f loop-body %.0f / %.0f | cta total %.0f (nkv %d)\n",
                h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n, h[7] / n, h[8] / n, h[9] / n, h[10] / n, h[11] / n,
                h[12] / n, (a.Nk + 127) / 128);
      }
    }
  } else if (a.d == 80) {
    if (a.legacy) {
      launch_attn_t<2, 5, 80, 2, 2>(stream, a, p, tq, tk, tv);
    } else {  // two query tiles per CTA, warp-specialised (softmax of one tile overlaps the MMAs of the other)
      dim3 grid((unsigned)ceil_div(a.Nq, 256), (unsigned)a.heads, (unsigned)a.B);
      static const int use_2x = getenv("SDTF_ATTN_2X") ? atoi(getenv("SDTF_ATTN_2X")) : 1;  // A/B: 0 = full-row kernel (round 1)
      if (use_2x) {  // work items on a 1-D grid; persistent (SDTF_ATTN_PERSIST as for d = 40) when an item has >= 2 key tiles
        static const int persist = getenv("SDTF_ATTN_PERSIST") ? atoi(getenv("SDTF_ATTN_PERSIST")) : 1;
        p.heads = a.heads; p.n_qt = ceil_div(a.Nq, 256); p.n_items = p.n_qt * a.heads * a.B;
        const int cap = persist > 1 ? persist : sm_count();
        grid = dim3((unsigned)(persist && a.Nk > 128 && p.n_items > cap ? cap : p.n_items));
      }
      if (use_2x == 2) launch_pdl(attn2x_kernel<5, 96, 3, 2, 1>, grid, dim3(kAHThreads), attn2x_smem_bytes<2, 1>(), stream, 1, tq, tk, tv, p);
      else if (use_2x) launch_pdl(attn2x_kernel<5, 96, 3, 1, 2>, grid, dim3(kAHThreads), attn2x_smem_bytes<1, 2>(), stream, 1, tq, tk, tv, p);
      else launch_pdl(attn2q_kernel<2, 5, 80, 2, 1>, grid, dim3(kA2Threads), attn2q_smem_bytes<2, 2, 1>(), stream, 1, tq, tk, tv, p);
    }
  } else if (a.d == 160) {
    launch_attn_t<3, 10, 160, 1, 1>(stream, a, p, tq, tk, tv);
  } else {
    throw Error("attention: unsupported head size " + std::to_string(a.d));
  }
}

}  // namespace sdtf
