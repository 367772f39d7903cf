// gemm3.cuh — persistent, warp-specialised implicit-GEMM kernel with a TMA epilogue and optional CTA pairs
// (same math and operand layout as gemm.cuh; reference call sites: layers.py:17-25 PaddedConv2D,
// diffusion_model.py:30,38,62,67,90,102-108,146).
//
// What bounds this kernel on B200 is not the tensor pipe but operand delivery: a 128 x 160 x 64 step needs 36 KB
// from L2 per 320 tensor-core cycles.  Two things here attack that:
//   * CG = 2: a CTA pair (tcgen05 cta_group::2) owns a 256 x BN tile.  Each CTA loads its own 128 activation rows
//     and only HALF of the weight tile; the pair's MMA (M = 256) reads both halves.  L2 -> SM bytes per FLOP drop
//     by 28 % at BN = 160 and the smaller stage buys a deeper TMA ring (7 stages instead of 5).
//   * the epilogue no longer computes addresses: accumulator rows go TMEM -> registers -> bf16 into a 64-byte-swizzled
//     128 x 32 staging tile, which leaves as ONE TMA tensor store (clipped at the tensor edges by the hardware);
//     residual tiles arrive the same way (TMA load into the staging buffer two passes ahead, summed in place).
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (pair: leader CTA only) and TMEM owner, warps 2-5 = epilogue.
// Pipelines: smem ring full/empty (TMA <-> MMA), two TMEM accumulators tfull/tempty (MMA <-> epilogue), four
// staging buffers tracked with bulk async-groups (epilogue <-> TMA store) and res_full barriers (TMA load -> epilogue).
#pragma once
#include "gemm.cuh"

namespace sdtf {

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_tanh_fast(float g) {
  const float u = g * 0.7978845608f * (1.f + 0.044715f * g * g);
  return 0.5f * g * (1.f + tanh_approx(u));
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct Gemm3Extra {
  int m_tiles, n_tiles;  // 128-row M tiles, BN-wide N tiles
  int acc_stride;        // TMEM column distance between the two accumulator stages
  int ncols;             // output columns per tile (BN, or BN/2 for GEGLU)
  int log_rows_per_b;    // log2(bw*bh): tile row >> this = sample offset inside the tile
};

static constexpr int kG3Threads = 192;
static constexpr int kG3Bufs = 4;                 // staging buffers
static constexpr int kG3ResAhead = 2;             // residual prefetch distance (passes)
static constexpr uint32_t kG3BufBytes = 128 * 64; // 128 rows x 32 bf16

template <int CG>
__global__ void __launch_bounds__(kG3Threads, 1)
conv_gemm3_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                  const __grid_constant__ CUtensorMap tmRes, const GemmParams p, const Gemm3Extra x) {
  extern __shared__ uint8_t smem_raw[];
  using namespace tc05;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t b_rows = (uint32_t)p.BN / CG;  // weight rows this CTA loads per stage
  const uint32_t stage_bytes = kATileBytes + b_rows * 128u;
  const uint32_t stg_off = (uint32_t)p.stages * stage_bytes;
  const uint32_t bar_off = stg_off + kG3Bufs * kG3BufBytes;
  const uint32_t bar_base = smem_base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + 2 + a); };
  auto res_bar = [&](int b) { return bar_base + 8u * (2 * p.stages + 4 + b); };
  const uint32_t slot_off = bar_off + 8u * (2 * p.stages + 4 + kG3Bufs);
  const uint32_t tmem_slot = smem_base + slot_off;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + slot_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int kchunks = p.kc0 + p.kc1;
  const int iters = p.taps * kchunks;
  const int m_units = (x.m_tiles + CG - 1) / CG;
  const int total_units = m_units * x.n_tiles;
  const int unit0 = (int)blockIdx.x / CG, unit_step = (int)gridDim.x / CG;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmOut);
    if (p.kc1) prefetch_tmap(&tmA1);
    if (p.res) prefetch_tmap(&tmRes);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128 * CG);
    }
    for (int b = 0; b < kG3Bufs; ++b) mbar_init(res_bar(b), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_pair(tmem_slot, (uint32_t)p.tmem_cols);
    else tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  }
  fence_before_sync();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // peer barriers initialised before any remote arrive / multicast commit
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // unit u -> this CTA's tile: n tile and the pixel box origin (out of range for the phantom tile of an odd pair)
  auto tile_coords = [&](int u, int& n_tile, int& x0, int& y0, int& b0) {
    n_tile = u % x.n_tiles;
    int mt = (u / x.n_tiles) * CG + (int)rank;
    const int tx = mt % p.tiles_x; mt /= p.tiles_x;
    const int ty = mt % p.tiles_y; mt /= p.tiles_y;
    x0 = tx * p.bw; y0 = ty * p.bh; b0 = mt * p.bn;
  };

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer =====
      int it = 0;
      for (int u = unit0; u < total_units; u += unit_step) {
        int n_tile, x0, y0, b0;
        tile_coords(u, n_tile, x0, y0, b0);
        const int nrow0 = n_tile * p.BN + (int)(rank * b_rows);
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
            if (CG == 1) {
              mbar_expect_tx(full_bar(st), stage_bytes);
              if (kc < p.kc0) tma_load_4d(sa, &tmA0, full_bar(st), kc * kBK, cx, cy, b0);
              else            tma_load_4d(sa, &tmA1, full_bar(st), (kc - p.kc0) * kBK, cx, cy, b0);
              tma_load_3d(sb, &tmB, full_bar(st), kc * kBK, nrow0, tap);
            } else {
              if (rank == 0) mbar_expect_tx(full_bar(st), 2u * stage_bytes);  // both CTAs' bytes land on the leader's barrier
              if (kc < p.kc0) tma_load_4d_pair(sa, &tmA0, full_bar(st), kc * kBK, cx, cy, b0);
              else            tma_load_4d_pair(sa, &tmA1, full_bar(st), (kc - p.kc0) * kBK, cx, cy, b0);
              tma_load_3d_pair(sb, &tmB, full_bar(st), kc * kBK, nrow0, tap);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      // ===== MMA issuer (pair: leader CTA only) =====
      const uint32_t idesc = make_idesc_bf16(kBM * CG, p.BN, 0, 0);
      int it = 0, lt = 0;
      for (int u = unit0; u < total_units; u += unit_step, ++lt) {
        const int acc = lt & 1;
        mbar_wait(tempty_bar(acc), ((uint32_t)(lt >> 1) & 1u) ^ 1u);  // both CTAs' epilogues drained this accumulator
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * x.acc_stride);
        for (int i = 0; i < iters; ++i, ++it) {
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
            if (CG == 2) mma_f16_ss_pair(d_tmem, da, db, idesc, (i | k) != 0);
            else mma_f16_ss(d_tmem, da, db, idesc, (i | k) != 0);
          }
          if (CG == 2) mma_commit_pair(empty_bar(st)); else mma_commit(empty_bar(st));
        }
        if (CG == 2) mma_commit_pair(tfull_bar(acc)); else mma_commit(tfull_bar(acc));
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps 2..5: TMEM lane quadrant = warp % 4 =====
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const bool issuer = (threadIdx.x == 64);
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const bool geglu = (p.act == ACT_GEGLU);
    const int ncols = x.ncols;
    const int Nout = geglu ? p.N / 2 : p.N;
    const int passes = (ncols + 31) >> 5;
    const bool has_res = p.res != nullptr;
    const int sw = (row >> 1) & 3;  // 64-byte swizzle: 16-byte chunk c of row r lives at chunk c ^ ((r >> 1) & 3)
    auto pass_col = [&](int ps) { return (ps * 32 + 32 <= ncols || ncols < 32) ? ps * 32 : ncols - 32; };

    // residual prefetch cursor (issuer only): pass pf_g = (unit pf_u, pass pf_ps)
    int pf_u = unit0, pf_ps = 0, pf_g = 0;
    auto issue_res = [&]() {
      if (pf_u >= total_units) return;
      int n_tile, x0, y0, b0;
      tile_coords(pf_u, n_tile, x0, y0, b0);
      const int buf = pf_g % kG3Bufs;
      mbar_expect_tx(res_bar(buf), kG3BufBytes);
      tma_load_4d(smem_base + stg_off + (uint32_t)buf * kG3BufBytes, &tmRes, res_bar(buf), n_tile * ncols + pass_col(pf_ps), x0,
                  y0, b0);
      ++pf_g;
      if (++pf_ps == passes) { pf_ps = 0; pf_u += unit_step; }
    };
    if (has_res && issuer)
      for (int k = 0; k < kG3ResAhead; ++k) issue_res();

    int lt = 0, g = 0;
    for (int u = unit0; u < total_units; u += unit_step, ++lt) {
      int n_tile, x0, y0, b0;
      tile_coords(u, n_tile, x0, y0, b0);
      const int acc = lt & 1;
      const uint32_t t_row = tmem_base + (uint32_t)(acc * x.acc_stride) + lane_off;
      int ob = b0 + (row >> x.log_rows_per_b);
      ob = ob < p.B ? ob : p.B - 1;
      mbar_wait(tfull_bar(acc), (uint32_t)(lt >> 1) & 1u);
      fence_after_sync();
      for (int ps = 0; ps < passes; ++ps, ++g) {
        const int buf = g % kG3Bufs;
        const int tc = pass_col(ps);          // column inside the tile's output slice
        const int col = n_tile * ncols + tc;  // global output column
        uint32_t v[32];
        float f[32];
        tmem_ld32(t_row + (uint32_t)tc, v);
        if (geglu) {
          uint32_t gv[32];
          tmem_ld32(t_row + (uint32_t)(ncols + tc), gv);
          tmem_ld_wait();
          const float* bp = p.bias + n_tile * p.BN + tc;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), bg = bv;
            if (p.bias) {
              bv = __ldg(reinterpret_cast<const float4*>(bp + i));
              bg = __ldg(reinterpret_cast<const float4*>(bp + ncols + i));
            }
            f[i] = (__uint_as_float(v[i]) + bv.x) * gelu_tanh_fast(__uint_as_float(gv[i]) + bg.x);
            f[i + 1] = (__uint_as_float(v[i + 1]) + bv.y) * gelu_tanh_fast(__uint_as_float(gv[i + 1]) + bg.y);
            f[i + 2] = (__uint_as_float(v[i + 2]) + bv.z) * gelu_tanh_fast(__uint_as_float(gv[i + 2]) + bg.z);
            f[i + 3] = (__uint_as_float(v[i + 3]) + bv.w) * gelu_tanh_fast(__uint_as_float(gv[i + 3]) + bg.w);
          }
        } else {
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * p.out_scale;
          if (col + 32 <= Nout) {
            if (p.bias) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + col + i));
                f[i] += bv.x; f[i + 1] += bv.y; f[i + 2] += bv.z; f[i + 3] += bv.w;
              }
            }
            if (p.temb) {
              const float* tp = p.temb + (long long)ob * p.temb_ld + col;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 tv = __ldg(reinterpret_cast<const float4*>(tp + i));
                f[i] += tv.x; f[i + 1] += tv.y; f[i + 2] += tv.z; f[i + 3] += tv.w;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (col + i < Nout) {
                if (p.bias) f[i] += __ldg(p.bias + col + i);
                if (p.temb) f[i] += __ldg(p.temb + (long long)ob * p.temb_ld + col + i);
              }
            }
          }
        }
        if (ps + 1 == passes) {  // accumulator fully read: hand it back to the MMA warp (of the leader CTA)
          fence_before_sync();
          if (CG == 2) mbar_arrive_cluster(tempty_bar(acc), 0);
          else mbar_arrive(tempty_bar(acc));
        }
        uint8_t* my_row = smem_gen + stg_off + (uint32_t)buf * kG3BufBytes + (uint32_t)row * 64u;
        if (has_res) {
          mbar_wait(res_bar(buf), (uint32_t)(g / kG3Bufs) & 1u);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 r4 = *reinterpret_cast<const uint4*>(my_row + ((c ^ sw) << 4));
            const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&rw[i]);
              f[8 * c + 2 * i] += __bfloat162float(h.x);
              f[8 * c + 2 * i + 1] += __bfloat162float(h.y);
            }
          }
        }
        if (p.act == ACT_SILU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = silu_f(f[i]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 o;
          o.x = pack_bf16(f[8 * c + 0], f[8 * c + 1]);
          o.y = pack_bf16(f[8 * c + 2], f[8 * c + 3]);
          o.z = pack_bf16(f[8 * c + 4], f[8 * c + 5]);
          o.w = pack_bf16(f[8 * c + 6], f[8 * c + 7]);
          *reinterpret_cast<uint4*>(my_row + ((c ^ sw) << 4)) = o;
        }
        fence_proxy_async_smem();  // staged tile (generic proxy) -> visible to the TMA store (async proxy)
        // without residual loads gating the buffers: the store that last used the NEXT pass's buffer must be done reading
        if (!has_res && issuer) bulk_wait_read<kG3Bufs - 2>();
        epi_bar_sync();
        if (issuer) {
          tma_store_4d(&tmOut, smem_base + stg_off + (uint32_t)buf * kG3BufBytes, col, x0, y0, b0);
          bulk_commit();
          if (has_res) {
            bulk_wait_read<kG3Bufs - kG3ResAhead>();  // buffer of pass g + ResAhead is free again
            issue_res();
          }
        }
      }
    }
    if (issuer) bulk_wait_all();
  }

  // teardown: everyone (in both CTAs of a pair) done with TMEM before the allocating warp frees it
  fence_before_sync();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 1) {
    if (CG == 2) tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

}  // namespace sdtf
