// gemm2.cuh — persistent, warp-specialised version of the implicit-GEMM kernel (same math and operand layout as
// gemm.cuh; reference call sites: layers.py:17-25 PaddedConv2D, diffusion_model.py:30,38,62,67,90,102-108,146).
//
// Why a second kernel: with one output tile per CTA (gemm.cuh) the short-K contractions of the transformer blocks
// (K = 320/640: 5-10 k-iterations, ~0.8 us of tensor-core work per tile) spend most of their time in the per-CTA
// prologue (TMEM alloc, barrier init, first TMA round trip) and in an epilogue whose row-per-thread global accesses
// touch 32 cache lines per warp instruction.  Here:
//   * grid = min(tiles, #SMs); every CTA walks tiles t = blockIdx.x, +gridDim.x, ... (n fastest, so CTAs running at
//     the same time share A tiles and stream the whole weight matrix through L2 together);
//   * warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue.  The smem ring and the two TMEM accumulator
//     stages run across tile boundaries: the producer prefetches tile i+1 and the tensor core computes it while the
//     epilogue warps drain tile i;
//   * epilogue: the residual tile is prefetched into shared memory with cp.async (coalesced 16-byte chunks), each
//     epilogue thread then owns one accumulator row: TMEM -> registers, + bias, + time-embedding column, + residual
//     (fp32, single rounding), GEGLU gate / SiLU, -> bf16 written IN PLACE over the residual tile; after a named
//     barrier the tile leaves as coalesced 16-byte stores.  Two staging buffers alternate so the next residual tile
//     is in flight while the current one is being stored.
#pragma once
#include "gemm.cuh"

namespace sdtf {

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_tanh_fast(float g) {
  const float u = g * 0.7978845608f * (1.f + 0.044715f * g * g);
  return 0.5f * g * (1.f + tanh_approx(u));
}

struct Gemm2Extra {
  int m_tiles, n_tiles;  // tile grid (m tiles = tiles_x * tiles_y * tiles_b)
  int W;                 // output columns handled per epilogue pass (multiple of 16)
  int passes;            // passes per tile (1 or 2)
  int pitch;             // staging row pitch in bytes = W*2 + 16
  int acc_stride;        // TMEM column distance between the two accumulator stages
};

static constexpr int kG2Threads = 192;

__global__ void __launch_bounds__(kG2Threads, 1)
conv_gemm2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB, const GemmParams p, const Gemm2Extra x) {
  extern __shared__ uint8_t smem_raw[];
  using namespace tc05;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t stage_bytes = kATileBytes + (uint32_t)p.BN * 128u;
  const uint32_t stg_off = (uint32_t)p.stages * stage_bytes;          // 2 staging buffers
  const uint32_t stg_bytes = 128u * (uint32_t)x.pitch;
  const uint32_t pix_off = stg_off + 2u * stg_bytes;                  // int row_pix[2][128]
  const uint32_t bar_off = pix_off + 2u * 128u * 4u;
  const uint32_t bar_base = smem_base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 8u * (2 * p.stages + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kchunks = p.kc0 + p.kc1;
  const int iters = p.taps * kchunks;
  const int total_tiles = x.m_tiles * x.n_tiles;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmB);
    if (p.kc1) prefetch_tmap(&tmA1);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  auto tile_coords = [&](int t, int& n_tile, int& x0, int& y0, int& b0) {
    n_tile = t % x.n_tiles;
    int mt = t / x.n_tiles;
    const int tx = mt % p.tiles_x; mt /= p.tiles_x;
    const int ty = mt % p.tiles_y; mt /= p.tiles_y;
    x0 = tx * p.bw; y0 = ty * p.bh; b0 = mt * p.bn;
  };

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer =====
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int n_tile, x0, y0, b0;
        tile_coords(t, n_tile, x0, y0, b0);
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
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer =====
      const uint32_t idesc = make_idesc_bf16(kBM, p.BN, 0, 0);
      int it = 0, lt = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
        const int acc = lt & 1;
        mbar_wait(tempty_bar(acc), ((uint32_t)(lt >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
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
            mma_f16_ss(d_tmem, da, db, idesc, (i | k) != 0);
          }
          mma_commit(empty_bar(st));
        }
        mma_commit(tfull_bar(acc));
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps 2..5: TMEM lane quadrant = warp % 4 =====
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int et = threadIdx.x - 64;  // 0..127
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const bool geglu = (p.act == ACT_GEGLU);
    const int ncols = geglu ? p.BN / 2 : p.BN;  // output columns of one tile
    const int Nout = geglu ? p.N / 2 : p.N;
    const int cpr = x.W >> 3;                   // 16-byte chunks per staged row
    const int rpi = 32 / cpr;                   // rows handled per warp iteration in the coalesced phases
    const int sub = lane / cpr, chk = lane - sub * cpr;
    const bool lane_on = sub < rpi;
    const int ew = warp - 2;                    // 0..3
    const bf16* res = p.res;
    int* row_pix_all = reinterpret_cast<int*>(smem_gen + pix_off);

    // pixel index of tile row r (or -1 outside the tensor)
    auto pix_of_row = [&](int r, int x0, int y0, int b0) -> int {
      int rr = r;
      const int lx = rr % p.bw; rr /= p.bw;
      const int ly = rr % p.bh; rr /= p.bh;
      const int ox = x0 + lx, oy = y0 + ly, ob = b0 + rr;
      if (ox >= p.W || oy >= p.H || ob >= p.B) return -1;
      return (ob * p.H + oy) * p.W + ox;
    };
    // coalesced residual prefetch of (tile t, pass ps) into staging buffer `buf`; also publishes that tile's row->pixel table
    auto prefetch = [&](int t, int ps, int buf) {
      int n_tile, x0, y0, b0;
      tile_coords(t, n_tile, x0, y0, b0);
      int* row_pix = row_pix_all + buf * 128;
      row_pix[et] = pix_of_row(et, x0, y0, b0);
      if (res != nullptr && lane_on) {
        const int col = n_tile * ncols + ps * x.W + chk * 8;
        const uint32_t sbase = smem_base + stg_off + (uint32_t)buf * stg_bytes + (uint32_t)chk * 16u;
        for (int r = ew * rpi + sub; r < 128; r += 4 * rpi) {
          const int pix = pix_of_row(r, x0, y0, b0);
          if (pix >= 0 && col + 8 <= Nout) cp_async16(sbase + (uint32_t)r * x.pitch, res + (long long)pix * p.res_ld + col);
        }
      }
      cp_async_commit();
    };

    int lt = 0, gp = 0;  // local tile counter, global pass counter
    if ((int)blockIdx.x < total_tiles) prefetch(blockIdx.x, 0, 0);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
      int n_tile, x0, y0, b0;
      tile_coords(t, n_tile, x0, y0, b0);
      const int acc = lt & 1;
      const uint32_t t_row = tmem_base + (uint32_t)(acc * x.acc_stride) + lane_off;
      for (int ps = 0; ps < x.passes; ++ps, ++gp) {
        const int buf = gp & 1;
        cp_async_wait_all();
        if (ps == 0) {
          mbar_wait(tfull_bar(acc), (uint32_t)(lt >> 1) & 1u);
          fence_after_sync();
        }
        epi_bar_sync();  // residual + row table of this pass visible; everyone is done storing the other buffer
        // next pass (possibly of the next tile) starts flowing into the other buffer
        if (ps + 1 < x.passes) prefetch(t, ps + 1, buf ^ 1);
        else if (t + (int)gridDim.x < total_tiles) prefetch(t + gridDim.x, 0, buf ^ 1);
        const int* row_pix = row_pix_all + buf * 128;
        uint8_t* my_row = smem_gen + stg_off + (uint32_t)buf * stg_bytes + (uint32_t)row * x.pitch;
        const int my_pix = row_pix[row];
        const int ob = my_pix >= 0 ? my_pix / (p.H * p.W) : 0;
        const bool use_res = (res != nullptr) && my_pix >= 0;
        // ---- phase 1: accumulator row -> bf16 in the staging buffer ----
        for (int c = 0; c < x.W; c += 16) {
          const int tc = ps * x.W + c;            // column inside the tile's output slice
          const int col = n_tile * ncols + tc;    // global output column
          uint32_t v[16];
          float f[16];
          tmem_ld16(t_row + (uint32_t)tc, v);
          if (geglu) {
            uint32_t g[16];
            tmem_ld16(t_row + (uint32_t)(ncols + tc), g);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float val = __uint_as_float(v[i]), gate = __uint_as_float(g[i]);
              if (p.bias) {
                val += __ldg(p.bias + n_tile * p.BN + tc + i);
                gate += __ldg(p.bias + n_tile * p.BN + ncols + tc + i);
              }
              f[i] = val * gelu_tanh_fast(gate);
            }
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * p.out_scale;
            if (col + 16 <= Nout) {
              if (p.bias) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                  const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + col + i));
                  f[i] += bv.x; f[i + 1] += bv.y; f[i + 2] += bv.z; f[i + 3] += bv.w;
                }
              }
              if (p.temb) {
                const float* tp = p.temb + (long long)ob * p.temb_ld + col;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                  const float4 tv = __ldg(reinterpret_cast<const float4*>(tp + i));
                  f[i] += tv.x; f[i + 1] += tv.y; f[i + 2] += tv.z; f[i + 3] += tv.w;
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (col + i < Nout) {
                  if (p.bias) f[i] += __ldg(p.bias + col + i);
                  if (p.temb) f[i] += __ldg(p.temb + (long long)ob * p.temb_ld + col + i);
                }
              }
            }
          }
          uint4* sp = reinterpret_cast<uint4*>(my_row + c * 2);
          if (use_res) {
            const uint4 r0 = sp[0], r1 = sp[1];
            const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&rw[i]);
              f[2 * i] += __bfloat162float(h.x);
              f[2 * i + 1] += __bfloat162float(h.y);
            }
          }
          if (p.act == ACT_SILU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = silu_f(f[i]);
          }
          uint4 o0, o1;
          o0.x = pack_bf16(f[0], f[1]);   o0.y = pack_bf16(f[2], f[3]);
          o0.z = pack_bf16(f[4], f[5]);   o0.w = pack_bf16(f[6], f[7]);
          o1.x = pack_bf16(f[8], f[9]);   o1.y = pack_bf16(f[10], f[11]);
          o1.z = pack_bf16(f[12], f[13]); o1.w = pack_bf16(f[14], f[15]);
          sp[0] = o0;
          sp[1] = o1;
        }
        if (ps + 1 == x.passes) {  // accumulator fully read: hand it back to the MMA warp
          fence_before_sync();
          mbar_arrive(tempty_bar(acc));
        }
        epi_bar_sync();
        // ---- phase 2: coalesced 16-byte stores ----
        if (lane_on) {
          const int col = n_tile * ncols + ps * x.W + chk * 8;
          if (col + 8 <= Nout) {
            const uint8_t* sbase = smem_gen + stg_off + (uint32_t)buf * stg_bytes + (uint32_t)chk * 16u;
            bf16* obase = reinterpret_cast<bf16*>(p.out) + col;
            for (int r = ew * rpi + sub; r < 128; r += 4 * rpi) {
              const int pix = row_pix[r];
              if (pix >= 0)
                *reinterpret_cast<uint4*>(obase + (long long)pix * p.out_ld) = *reinterpret_cast<const uint4*>(sbase + (uint32_t)r * x.pitch);
            }
          }
        }
      }
    }
    cp_async_wait_all();
  }

  // teardown: everyone done with TMEM before the allocating warp frees it
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace sdtf
