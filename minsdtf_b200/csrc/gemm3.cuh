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
//   * N tiles wider than one MMA (BN up to 512 = two N <= 256 instructions on the same A stage) cut the bytes per
//     FLOP further; above 256 columns the accumulator is single-buffered (512 TMEM columns), which long-K convs
//     amortise.  Measured (ncu, 3x3 conv): tensor pipe 50 % busy at 256x160 tiles, 80 % at 256x256 — the L2 -> SM
//     path delivers ~50 B/clk/SM, so tile area per byte is what sets the rate.
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (pair: leader CTA only) and TMEM owner, warps 2-3 = store warps (TMA
// stores of staged sub-tiles, residual prefetch / buffer grants; one per epilogue set), warps 4-11 = epilogue (two sets).
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM accumulators tfull/tempty (MMA <-> epilogue), four staging
// buffers: grant (res_bar: buffer free, residual landed) -> epilogue -> stg_full -> store warp -> bulk-group wait.
#pragma once
#include "gemm.cuh"

namespace sdtf {

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// barrier of ONE epilogue warp set (128 threads; ids 1 and 2): the two sets never wait for each other — a tile's passes
// split 3 / 2 between them, and with a common barrier per tile the lighter set idled a pass per tile
__device__ __forceinline__ void epi_set_bar_sync(int set) { asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory"); }
__device__ __forceinline__ float gelu_tanh_fast(float g) {
  const float u = g * 0.7978845608f * (1.f + 0.044715f * g * g);
  return 0.5f * g * (1.f + tanh_approx(u));
}

// u / d for the small non-negative integers of the tile bookkeeping as one multiply-high: q = (u * ceil(2^32 / d)) >> 32,
// exact while u * d < 2^32 (the host checks it).  A run-time integer division is ~25 SASS instructions, and the tile
// decode has six of them at four call sites: a sixth of the kernel's code before this.
struct FastDiv {
  uint32_t mul = 0, d = 1;
#ifdef __CUDACC__
  __device__ __forceinline__ int div(int u) const { return d == 1 ? u : (int)__umulhi((uint32_t)u, mul); }
#endif
  static FastDiv make(int d) {
    FastDiv f;
    f.d = (uint32_t)(d < 1 ? 1 : d);
    f.mul = f.d == 1 ? 0u : (uint32_t)((0x100000000ull + f.d - 1) / f.d);
    return f;
  }
};

struct Gemm3Extra {
  int m_tiles, n_tiles;  // 128-row M tiles, BN-wide N tiles
  int acc_stride;        // TMEM column distance between accumulator stages
  int acc_bufs;          // 2: double-buffered accumulator (BN <= 256), 1: single (BN up to 512)
  int n_mma;             // MMA instructions along N per k-step (BN / n_mma columns each, <= 256)
  int ncols;             // output columns per tile (BN, or BN/2 for GEGLU)
  int log_rows_per_b;    // log2(bw*bh): tile row >> this = sample offset inside the tile
  int log_bw;            // log2(bw) (tile shapes are powers of two)
  int nbufs;             // staging buffers in use (<= kG3MaxBufs)
  int vec_rows;          // rows of the per-tile epilogue vector staged in smem: samples a tile spans (time-embedding conv) or 1
  int vec_width;         // its width in floats (tile columns, rounded up to 32)
  int splits;            // split-K: the k-iterations of a tile are cut into `splits` ranges, one unit each (1: off)
  int iters_split;       // k-iterations per range (the last range may be shorter)
  float* partial;        // split-K: fp32 partial sums [splits][B*H*W][N] (the reduce kernel applies the epilogue)
  long long* prof;       // optional [gridDim.x][16] cycle counters per role (null: off); see gemm_host.cuh
  int debug;             // timing experiments only (results are garbage): 1 = skip the MMA instructions, 2 = skip the TMA loads,
                         // 4 = no proxy fence after staging, 8 = no staging stores
  FastDiv d_ntiles, d_munits, d_tx, d_ty;  // divisors of the tile decode: n_tiles, ceil(m_tiles / CG), tiles_x, tiles_y
  int head_stride;       // > 0: output columns are heads of this many columns, tmOut is 5-D (make_epi_tmap_heads) and clips each head
};

// cycle accounting for the role loops (only when x.prof is set): t += clock spent inside a wait
#define G3_TIMED(prof_on, acc, stmt)            \
  do {                                          \
    if (prof_on) {                              \
      const long long _t0 = clock64();          \
      stmt;                                     \
      acc += clock64() - _t0;                   \
    } else {                                    \
      stmt;                                     \
    }                                           \
  } while (0)

static constexpr int kG3Threads = 384;
static constexpr int kG3MaxBufs = 8;              // staging buffers: x.nbufs of them (4, or 8 when residual tiles must be prefetched deep)
static constexpr uint32_t kG3BufBytes = 128 * 64; // 128 rows x 32 bf16

// MODE: which epilogue variants an instantiation carries.  The general kernel (every variant behind run-time flags) is 5 800
// SASS instructions (93 KB) and ncu showed its epilogue warps stalled on instruction fetch (no_inst 18 % of the samples of
// the HBM-bound 1x1 GEMMs): a launch only ever needs one variant, so each gets its own, smaller kernel.  Same box, isolated:
// 320->320 +res @64x64 39 -> 33 us, q|k|v 89 -> 72 us, 3x3 320->320 96 -> 91 us (profiles/r02_j_lean_bench_cases.log).
enum : int {
  G3_GENERAL = 0,     // everything: split-K partial sums, quick-GELU, per-role profiler, and all of the below
  G3_LEAN = 1,        // bias / time embedding / residual / SiLU
  G3_GEGLU = 2,       // + GEGLU gate
  G3_LN_PRODUCE = 3,  // + LayerNorm statistics of the output rows
  G3_LN_APPLY = 4,    // + LayerNorm applied to the accumulator (gamma folded into the weights)
  G3_GEGLU_LN = 5,    // GEGLU gate on a LayerNorm-folded projection (SDTF_LN_FOLD=1)
};
template <int CG, int MODE = G3_GENERAL>
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
  const int kG3Bufs = x.nbufs;  // 4 or 8
  const int nb_mask = kG3Bufs - 1, nb_shift = kG3Bufs == 8 ? 3 : 2;
  const uint32_t bar_off = stg_off + (uint32_t)kG3Bufs * kG3BufBytes;
  const uint32_t bar_base = smem_base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + 2 + a); };
  auto res_bar = [&](int b) { return bar_base + 8u * (2 * p.stages + 4 + b); };
  auto stg_bar = [&](int b) { return bar_base + 8u * (2 * p.stages + 4 + kG3Bufs + b); };
  const uint32_t slot_off = bar_off + 8u * (2 * p.stages + 4 + 2 * kG3Bufs);
  const uint32_t vec_off = slot_off + 16u;  // float [2 sets][2][vec_rows][vec_width]: bias (+ time-embedding) of the current tile, one copy per epilogue warp set
  const uint32_t tmem_slot = smem_base + slot_off;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + slot_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int kchunks = p.kc0 + p.kc1;
  const int iters = p.taps * kchunks;
  const int m_units = (x.m_tiles + CG - 1) / CG;
  const int total_units = m_units * x.n_tiles * x.splits;
  constexpr bool kAll = MODE == G3_GENERAL;
  constexpr bool kGegluOnly = MODE == G3_GEGLU || MODE == G3_GEGLU_LN;
  constexpr bool kLnOut = kAll || MODE == G3_LN_PRODUCE,
                 kLnIn = kAll || MODE == G3_LN_APPLY || MODE == G3_GEGLU_LN;
  const bool partial_mode = kAll && x.partial != nullptr;
  const int unit0 = (int)blockIdx.x / CG, unit_step = (int)gridDim.x / CG;
  const bool prof_on = kAll && x.prof != nullptr;
  long long* prof = prof_on ? x.prof + (long long)blockIdx.x * 16 : nullptr;
  const long long t_start = prof_on ? clock64() : 0;

  pdl_trigger();
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
      mbar_init(tempty_bar(a), 8 * CG);  // one arrive per epilogue warp
    }
    for (int b = 0; b < kG3Bufs; ++b) {
      mbar_init(res_bar(b), 1);
      mbar_init(stg_bar(b), 4);
    }
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
  // split-K: unit u covers k-iterations [i0, i0 + n_it) of its tile
  auto k_range = [&](int u, int& i0, int& n_it) {
    const int ks = x.d_munits.div(x.d_ntiles.div(u));
    i0 = ks * x.iters_split;
    n_it = iters - i0 < x.iters_split ? iters - i0 : x.iters_split;
  };
  pdl_wait();  // activations, residual and the time-embedding table come from earlier kernels of the stream

  // unit u -> this CTA's tile: n tile and the pixel box origin (out of range for the phantom tile of an odd pair)
  auto tile_coords = [&](int u, int& n_tile, int& x0, int& y0, int& b0) {
    const int un = x.d_ntiles.div(u);
    n_tile = u - un * x.n_tiles;
    int mt = (un - x.d_munits.div(un) * m_units) * CG + (int)rank;
    int q = x.d_tx.div(mt);
    const int tx = mt - q * p.tiles_x;
    mt = q;
    q = x.d_ty.div(mt);
    const int ty = mt - q * p.tiles_y;
    x0 = tx * p.bw; y0 = ty * p.bh; b0 = q * p.bn;
  };

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer =====
      // (one thread: everything per k-iteration is kept incremental — no divisions, no address rebuilds; measured:
      //  the naive loop cost ~570 cycles per iteration of scalar work and was the bottleneck of the whole kernel)
      long long w_empty = 0;
      uint32_t st = 0, ph = 0;  // ring slot and the parity of its current fill
      uint32_t sa = smem_base, fb = full_bar(0), eb = empty_bar(0);
      const uint32_t chunk_rows = b_rows / (uint32_t)x.n_mma;  // weight rows per MMA chunk held by this CTA
      const uint32_t chunk_stride = chunk_rows * 128u;
      const int chunk = p.BN / x.n_mma;
      const uint32_t tx_bytes = (uint32_t)CG * stage_bytes;
      const uint32_t n_stages = (uint32_t)p.stages;
      for (int u = unit0; u < total_units; u += unit_step) {
        int n_tile, x0, y0, b0;
        tile_coords(u, n_tile, x0, y0, b0);
        const int nrow0 = n_tile * p.BN + (int)(rank * chunk_rows);
        const int cx0 = x0 * p.stride - p.pad_x, cy0 = y0 * p.stride - p.pad_y;
        int i0, n_it;
        k_range(u, i0, n_it);
        int tap = 0, kc = 0, r = 0, sx = 0;  // (divisions only for a split-K range that starts inside the filter; none per iteration)
        if (i0 != 0) {
          tap = i0 / kchunks; kc = i0 - tap * kchunks;
          r = tap / p.tap_w; sx = tap - r * p.tap_w;
        }
        for (int it = 0; it < n_it; ++it) {
          const int cx = cx0 + sx, cy = cy0 + r;
          G3_TIMED(prof_on, w_empty, mbar_wait(eb, ph ^ 1u));
          const uint32_t sb = sa + kATileBytes;
          const CUtensorMap* ta = kc < p.kc0 ? &tmA0 : &tmA1;
          const int ka = (kc < p.kc0 ? kc : kc - p.kc0) * kBK;
          if (x.debug & 2) {
            if (rank == 0) mbar_arrive(fb);
          } else if (CG == 1) {
            mbar_expect_tx(fb, tx_bytes);
            tma_load_4d(sa, ta, fb, ka, cx, cy, b0);
            tma_load_3d(sb, &tmB, fb, kc * kBK, nrow0, tap);
            if (x.n_mma == 2) tma_load_3d(sb + chunk_stride, &tmB, fb, kc * kBK, nrow0 + chunk, tap);
          } else {
            if (rank == 0) mbar_expect_tx(fb, tx_bytes);  // both CTAs' bytes land on the leader's barrier
            tma_load_4d_pair(sa, ta, fb, ka, cx, cy, b0);
            tma_load_3d_pair(sb, &tmB, fb, kc * kBK, nrow0, tap);
            if (x.n_mma == 2) tma_load_3d_pair(sb + chunk_stride, &tmB, fb, kc * kBK, nrow0 + chunk, tap);
          }
          sa += stage_bytes; fb += 8u; eb += 8u;
          if (++st == n_stages) { st = 0; ph ^= 1u; sa = smem_base; fb = full_bar(0); eb = empty_bar(0); }
          if (++kc == kchunks) {
            kc = 0; ++tap;
            if (++sx == p.tap_w) { sx = 0; ++r; }
          }
        }
      }
      if (prof_on) { prof[0] = w_empty; prof[1] = clock64() - t_start; }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      // ===== MMA issuer (pair: leader CTA only) =====
      const int chunk = p.BN / x.n_mma;
      const uint32_t idesc = make_idesc_bf16(kBM * CG, chunk, 0, 0);
      // UMMA descriptors differ between ring slots / k-steps / N chunks only in their 14-bit start-address field
      // (bytes >> 4), so they are built once and advanced by integer adds
      const uint64_t da0 = make_smem_desc_sw128(smem_base, 16, 1024);
      const uint64_t db0 = make_smem_desc_sw128(smem_base + kATileBytes, 16, 1024);
      const uint64_t stage_step = (uint64_t)(stage_bytes >> 4);
      const uint64_t chunk_step = (uint64_t)(((uint32_t)(chunk / CG) * 128u) >> 4);
      const uint32_t n_stages = (uint32_t)p.stages;
      const bool two = x.n_mma == 2, skip = (x.debug & 1) != 0;
      long long w_full = 0, w_tempty = 0;
      uint32_t st = 0, ph = 0;
      uint32_t fb = full_bar(0), eb = empty_bar(0);
      uint64_t da = da0, db = db0;
      uint32_t acc = 0, acc_ph = 0;
      for (int u = unit0; u < total_units; u += unit_step) {
        // both CTAs' epilogues drained this accumulator
        G3_TIMED(prof_on, w_tempty, mbar_wait(tempty_bar(acc), acc_ph ^ 1u));
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)x.acc_stride;
        const uint32_t d_tmem2 = d_tmem + (uint32_t)chunk;
        int i0, n_it;
        k_range(u, i0, n_it);
        for (int i = 0; i < n_it; ++i) {
          G3_TIMED(prof_on, w_full, mbar_wait(fb, ph));
          fence_after_sync();
          if (!skip) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint32_t accum = (i | k) != 0;
              if (CG == 2) {
                mma_f16_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, accum);
                if (two) mma_f16_ss_pair(d_tmem2, da + 2 * k, db + chunk_step + 2 * k, idesc, accum);
              } else {
                mma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, accum);
                if (two) mma_f16_ss(d_tmem2, da + 2 * k, db + chunk_step + 2 * k, idesc, accum);
              }
            }
          }
          if (CG == 2) mma_commit_pair(eb); else mma_commit(eb);
          da += stage_step; db += stage_step; fb += 8u; eb += 8u;
          if (++st == n_stages) { st = 0; ph ^= 1u; da = da0; db = db0; fb = full_bar(0); eb = empty_bar(0); }
        }
        if (CG == 2) mma_commit_pair(tfull_bar(acc)); else mma_commit(tfull_bar(acc));
        if (++acc == (uint32_t)x.acc_bufs) { acc = 0; acc_ph ^= 1u; }
      }
      if (prof_on) { prof[2] = w_full; prof[3] = w_tempty; prof[4] = clock64() - t_start; }
    }
    __syncwarp();
  } else if (warp == 2 || warp == 3) {
    // ===== store warps: TMA stores of staged sub-tiles; grant staging buffers (with the residual tile when there is one).
    // One elected thread each: warp 2 serves the passes of even global index (epilogue set 0 and the even buffers), warp 3
    // the odd ones.  A single store thread was what bounded the short-K linears: ~700 cycles of serial bookkeeping per
    // pass (wait, coordinates, store, commit, read-wait, grant) against one pass per ~720 cycles measured at 320 -> 1536
    // (per-role counters: that thread busy 84 %, the MMA warp waiting for accumulators 41 % of the time).
    // (split-K partial sums leave straight from the epilogue warps' registers: nothing to stage)
    if (!partial_mode && elect_one()) {
      const int par = warp - 2;
      const int ncols = x.ncols;
      const int passes = (ncols + 31) >> 5;
      const bool has_res = p.res != nullptr;
      auto pass_col = [&](int ps) { return (ps * 32 + 32 <= ncols || ncols < 32) ? ps * 32 : ncols - 32; };
      const int my_units = unit0 < total_units ? (total_units - unit0 + unit_step - 1) / unit_step : 0;
      const int total_passes = my_units * passes;
      // grant cursor: pass gg = (unit gu, pass gps) of this thread's parity may use buffer gg & nb_mask
      int gu = unit0, gps = par, gg = par;
      while (gps >= passes) { gps -= passes; gu += unit_step; }
      auto grant = [&]() {
        if (gg >= total_passes) return;
        const int buf = gg & nb_mask;
        if (has_res) {
          int n_tile, x0, y0, b0;
          tile_coords(gu, n_tile, x0, y0, b0);
          mbar_expect_tx(res_bar(buf), kG3BufBytes);
          tma_load_4d(smem_base + stg_off + (uint32_t)buf * kG3BufBytes, &tmRes, res_bar(buf), n_tile * ncols + pass_col(gps),
                      x0, y0, b0);
        } else {
          mbar_arrive(res_bar(buf));
        }
        gg += 2; gps += 2;
        while (gps >= passes) { gps -= passes; gu += unit_step; }
      };
      for (int k = par; k < kG3Bufs; k += 2) grant();
      long long w_stg = 0, w_read = 0;
      int su = unit0, sps = par, issued = 0;
      while (sps >= passes) { sps -= passes; su += unit_step; }
      int n_tile = 0, x0 = 0, y0 = 0, b0 = 0, cu = -1;
      for (int sg = par; sg < total_passes; sg += 2) {
        if (su != cu) { tile_coords(su, n_tile, x0, y0, b0); cu = su; }
        const int buf = sg & nb_mask;
        // all 4 epilogue warps of the set staged (and fenced) their rows
        G3_TIMED(prof_on, w_stg, mbar_wait(stg_bar(buf), (uint32_t)(sg >> nb_shift) & 1u));
        const int col = n_tile * ncols + pass_col(sps);
        const uint32_t src = smem_base + stg_off + (uint32_t)buf * kG3BufBytes;
        if (x.head_stride) tma_store_5d(&tmOut, src, col % x.head_stride, col / x.head_stride, x0, y0, b0);
        else tma_store_4d(&tmOut, src, col, x0, y0, b0);
        bulk_commit();
        if (issued >= 1) {
          // this thread's previous store no longer reads its buffer: it can host that pass + kG3Bufs
          G3_TIMED(prof_on, w_read, bulk_wait_read<1>());
          grant();
        }
        ++issued;
        sps += 2;
        while (sps >= passes) { sps -= passes; su += unit_step; }
      }
      bulk_wait_all();
      if (prof_on && par == 0) { prof[5] = w_stg; prof[6] = w_read; prof[7] = clock64() - t_start; }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===== epilogue warps 4..11: two sets of four (set = passes of even / odd global index), TMEM lane quadrant = warp % 4.
    // Two epilogue warps per SM sub-partition hide each other's TMEM-load / shared-memory / MUFU latencies.
    const int quad = warp & 3;
    const int set = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const bool geglu = kGegluOnly || (kAll && p.act == ACT_GEGLU);  // (the GEGLU instantiations carry no other path)
    const int ncols = x.ncols;
    const int Nout = geglu ? p.N / 2 : p.N;
    float2* const ln_out = kLnOut ? p.ln_out : nullptr;
    const int passes = (ncols + 31) >> 5;
    const bool has_res = p.res != nullptr;
    const int sw = (row >> 1) & 3;  // 64-byte swizzle: 16-byte chunk c of row r lives at chunk c ^ ((r >> 1) & 3)
    auto pass_col = [&](int ps) { return (ps * 32 + 32 <= ncols || ncols < 32) ? ps * 32 : ncols - 32; };

    long long w_tfull = 0, w_grant = 0, w_ld = 0, w_stage = 0, w_vec = 0;
    const int et = (int)threadIdx.x - 128;  // 0..255 (set 0: 0..127)
    const bool ln_apply = kLnIn && p.ln_in != nullptr;
    const bool has_vec = p.bias != nullptr || p.temb != nullptr || ln_apply;
    const int rpb = 1 << x.log_rows_per_b;
    const int vrows = x.vec_rows, vwidth = x.vec_width;
    const int Nvec = geglu ? p.N : Nout;
    float* vec = reinterpret_cast<float*>(smem_gen + vec_off) + (size_t)set * 2 * x.vec_rows * x.vec_width;  // this set's double buffer
    int lt = 0;
    for (int u = unit0; u < total_units; u += unit_step, ++lt) {
      int n_tile, x0, y0, b0;
      tile_coords(u, n_tile, x0, y0, b0);
      const int acc = x.acc_bufs == 2 ? (lt & 1) : 0;
      const uint32_t t_row = tmem_base + (uint32_t)(acc * x.acc_stride) + lane_off;
      const int g0 = lt * passes;                       // global index of this tile's pass 0
      const int ps0 = (g0 ^ set) & 1;                   // first pass of this tile that belongs to this set
      const int ps_last = ps0 + ((passes - 1 - ps0) & ~1);  // last one (< ps0 if the set has none)
      int vr = row >> x.log_rows_per_b;  // sample of this row inside the tile
      vr = vr < vrows ? vr : vrows - 1;
      if (ln_apply) vr = 0;              // LayerNorm consumer: vector row 0 = c0 (bias), row 1 = c1
      // global pixel index of this thread's row (LayerNorm statistics are indexed by it)
      long long ln_m = -1;
      if (ln_apply || ln_out) {
        const int rr = row & (rpb - 1);
        const int pb = b0 + (row >> x.log_rows_per_b), py = y0 + (rr >> x.log_bw), px = x0 + (rr & (p.bw - 1));
        if (pb < p.B && py < p.H && px < p.W) ln_m = ((long long)pb * p.H + py) * p.W + px;
      }
      float ln_a = 1.f, ln_b = 0.f;
      if (ln_apply && ln_m >= 0) {  // issued before the accumulator wait: the loads hide behind the main loop
        float s1 = 0.f, s2 = 0.f;
        for (int sl = 0; sl < p.ln_slots; ++sl) {
          const float2 t2 = __ldg(p.ln_in + (long long)sl * p.ln_rows + ln_m);
          s1 += t2.x; s2 += t2.y;
        }
        const float mean = s1 * p.ln_inv_c;
        const float var = fmaxf(s2 * p.ln_inv_c - mean * mean, 0.f);
        ln_a = rsqrtf(var + 1e-5f);
        ln_b = -ln_a * mean;
      }

      // this tile's per-column vector (bias, + the time-embedding row of each sample the tile spans): fetched from
      // global BEFORE waiting for the accumulator so the latency hides behind the main loop, then staged in smem
      float4 pre[4];
      const int vcol = 4 * (et & 127);
      if (has_vec && vcol < vwidth) {
        const int gcol = n_tile * (geglu ? p.BN : ncols) + vcol;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < vrows && gcol < Nvec) {
            if (ln_apply && r == 1) a = __ldg(reinterpret_cast<const float4*>(p.ln_c1 + gcol));
            else if (p.bias) a = __ldg(reinterpret_cast<const float4*>(p.bias + gcol));
            if (p.temb && !ln_apply) {
              int bb = b0 + r;
              bb = bb < p.B ? bb : p.B - 1;
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.temb + (long long)bb * p.temb_ld + gcol));
              a.x += t4.x; a.y += t4.y; a.z += t4.z; a.w += t4.w;
            }
          }
          pre[r] = a;
        }
      }
      G3_TIMED(prof_on, w_tfull, mbar_wait(tfull_bar(acc), (uint32_t)(x.acc_bufs == 2 ? lt >> 1 : lt) & 1u));
      fence_after_sync();
      float* vtile = vec + (size_t)(lt & 1) * vrows * vwidth;
      const long long t_v0 = prof_on ? clock64() : 0;
      if (has_vec) {
        if (vcol < vwidth) {
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (r < vrows) *reinterpret_cast<float4*>(vtile + r * vwidth + vcol) = pre[r];
        }
        epi_set_bar_sync(set);  // vector visible to the set's four warps; also: they are done reading the tile before last's copy
      }
      if (prof_on) w_vec += clock64() - t_v0;
      const float* vrow = vtile + vr * vwidth;
      if (ps0 >= passes) {  // no pass of this tile is ours: nothing to read from the accumulator
        fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(tempty_bar(acc), 0);
          else mbar_arrive(tempty_bar(acc));
        }
      }
      if (partial_mode) {
        // split-K: raw fp32 accumulator rows -> partial[ks][pixel][column]; bias / time embedding / residual / bf16
        // rounding happen once, in splitk_reduce_kernel
        const int ks = x.d_munits.div(x.d_ntiles.div(u));
        const int rpb = 1 << x.log_rows_per_b;
        const int rr = row & (rpb - 1);
        const int pb = b0 + (row >> x.log_rows_per_b), py = y0 + (rr >> x.log_bw), px = x0 + (rr & (p.bw - 1));
        const bool row_ok = pb < p.B && py < p.H && px < p.W;
        float* prow = x.partial + ((long long)ks * p.B * p.H * p.W + ((long long)pb * p.H + py) * p.W + px) * p.N +
                      (long long)n_tile * ncols;
        for (int ps = ps0; ps < passes; ps += 2) {
          const int tc = pass_col(ps);
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)tc, v);
          tmem_ld_wait();
          if (ps == ps_last) {
            fence_before_sync();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2) mbar_arrive_cluster(tempty_bar(acc), 0);
              else mbar_arrive(tempty_bar(acc));
            }
          }
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              if (n_tile * ncols + tc + i < p.N)
                *reinterpret_cast<uint4*>(prow + tc + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        }
        continue;
      }
      for (int ps = ps0; ps < passes; ps += 2) {
        const int g = g0 + ps;
        const int buf = g & nb_mask;
        const int tc = pass_col(ps);          // column inside the tile's output slice
        uint32_t v[32];
        float f[32];
        const long long t_ld0 = prof_on ? clock64() : 0;
        tmem_ld32(t_row + (uint32_t)tc, v);
        if (prof_on && !geglu) { tmem_ld_wait(); w_ld += clock64() - t_ld0; }
        if (geglu) {
          uint32_t gv[32];
          tmem_ld32(t_row + (uint32_t)(ncols + tc), gv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), bg = bv;
            if (has_vec) {
              bv = *reinterpret_cast<const float4*>(vrow + tc + i);
              bg = *reinterpret_cast<const float4*>(vrow + ncols + tc + i);
            }
            if (ln_apply) {  // value / gate = rstd (acc - mean c1) + c0
              const float4 cv = *reinterpret_cast<const float4*>(vrow + vwidth + tc + i);
              const float4 cg = *reinterpret_cast<const float4*>(vrow + vwidth + ncols + tc + i);
              bv.x = fmaf(ln_b, cv.x, bv.x); bv.y = fmaf(ln_b, cv.y, bv.y); bv.z = fmaf(ln_b, cv.z, bv.z); bv.w = fmaf(ln_b, cv.w, bv.w);
              bg.x = fmaf(ln_b, cg.x, bg.x); bg.y = fmaf(ln_b, cg.y, bg.y); bg.z = fmaf(ln_b, cg.z, bg.z); bg.w = fmaf(ln_b, cg.w, bg.w);
            }
            // value * gelu(gate) on packed pairs: 8 FMA-pipe instructions + 2 MUFU.TANH per two outputs (scalar: 22 + 2)
            const f32x2 la2 = pack2(ln_a, ln_a);
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const int k = i + 2 * h2;
              const f32x2 bv2 = h2 ? pack2(bv.z, bv.w) : pack2(bv.x, bv.y), bg2 = h2 ? pack2(bg.z, bg.w) : pack2(bg.x, bg.y);
              f32x2 val = pack2(__uint_as_float(v[k]), __uint_as_float(v[k + 1]));
              f32x2 gate = pack2(__uint_as_float(gv[k]), __uint_as_float(gv[k + 1]));
              if (ln_apply) { val = fma2(la2, val, bv2); gate = fma2(la2, gate, bg2); }
              else { val = add2(val, bv2); gate = add2(gate, bg2); }
              // gelu(g) = 0.5 g (1 + tanh(g (0.79788456 + 0.03567741 g^2)))
              const f32x2 inner = fma2(mul2(gate, gate), pack2(0.0356774081f, 0.0356774081f), pack2(0.7978845608f, 0.7978845608f));
              float u0, u1;
              unpack2(mul2(inner, gate), u0, u1);
              const f32x2 th = pack2(tanh_approx(u0), tanh_approx(u1));
              const f32x2 hv = mul2(mul2(gate, pack2(0.5f, 0.5f)), val);  // 0.5 g v
              unpack2(fma2(hv, th, hv), f[k], f[k + 1]);
            }
          }
        } else {
          tmem_ld_wait();
          if (ln_apply) {  // y = rstd (acc - mean c1[n]) + c0[n]
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 c0v = *reinterpret_cast<const float4*>(vrow + tc + i);
              const float4 c1v = *reinterpret_cast<const float4*>(vrow + vwidth + tc + i);
              f[i] = fmaf(ln_a, __uint_as_float(v[i]), fmaf(ln_b, c1v.x, c0v.x));
              f[i + 1] = fmaf(ln_a, __uint_as_float(v[i + 1]), fmaf(ln_b, c1v.y, c0v.y));
              f[i + 2] = fmaf(ln_a, __uint_as_float(v[i + 2]), fmaf(ln_b, c1v.z, c0v.z));
              f[i + 3] = fmaf(ln_a, __uint_as_float(v[i + 3]), fmaf(ln_b, c1v.w, c0v.w));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * p.out_scale;
            if (has_vec) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(vrow + tc + i);
                f[i] += bv.x; f[i + 1] += bv.y; f[i + 2] += bv.z; f[i + 3] += bv.w;
              }
            }
          }
        }
        if (ps == ps_last) {  // this warp's last read of the accumulator: hand it back to the MMA warp (of the leader CTA)
          fence_before_sync();
          __syncwarp();
          if (lane == 0) {  // (an arrive per thread would serialise 128 shared-memory atomics on one barrier)
            if (CG == 2) mbar_arrive_cluster(tempty_bar(acc), 0);
            else mbar_arrive(tempty_bar(acc));
          }
        }
        uint8_t* my_row = smem_gen + stg_off + (uint32_t)buf * kG3BufBytes + (uint32_t)row * 64u;
        G3_TIMED(prof_on, w_grant, mbar_wait(res_bar(buf), (uint32_t)(g >> nb_shift) & 1u));  // buffer granted (residual landed)
        if (has_res) {
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
        if (kAll && p.act == ACT_QGELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = qgelu_f(f[i]);
        }
        float ln_s1 = 0.f, ln_s2 = 0.f;
        const long long t_st0 = prof_on ? clock64() : 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 o;
          o.x = pack_bf16(f[8 * c + 0], f[8 * c + 1]);
          o.y = pack_bf16(f[8 * c + 2], f[8 * c + 3]);
          o.z = pack_bf16(f[8 * c + 4], f[8 * c + 5]);
          o.w = pack_bf16(f[8 * c + 6], f[8 * c + 7]);
          if (!(x.debug & 8)) *reinterpret_cast<uint4*>(my_row + ((c ^ sw) << 4)) = o;
          if (ln_out) {  // statistics of what the consumer will read: the bf16-rounded values, summed in column order
            const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&ow[i]);
              const float a0 = __bfloat162float(h2.x), a1 = __bfloat162float(h2.y);
              ln_s1 += a0 + a1;
              ln_s2 = fmaf(a0, a0, fmaf(a1, a1, ln_s2));
            }
          }
        }
        // LayerNorm producer: one (sum, sum of squares) slot per 32-column pass, indexed by the pass's GLOBAL column
        // (column / 32).  Coarser slots (per N tile) would be cheaper but their grouping follows BN, which follows the
        // number of M tiles, i.e. the batch: a sample's statistics must not depend on what it is batched with.
        if (ln_out && ln_m >= 0)
          ln_out[(long long)((n_tile * ncols + tc) >> 5) * p.ln_rows + ln_m] = make_float2(ln_s1, ln_s2);
        // (deferring this hand-over by one pass, so that the fence finds the stores long complete, was slower: with two
        //  staging buffers per set the later store delays the grant the set needs two passes on — 63.5 -> 68 us at 320 -> 1536)
        if (!(x.debug & 4)) fence_proxy_async_smem();  // staged row (generic proxy) -> visible to the TMA store (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(stg_bar(buf));
        if (prof_on) w_stage += clock64() - t_st0;
      }
    }
    if (prof_on && threadIdx.x == 128) { prof[8] = w_tfull; prof[9] = w_grant; prof[10] = clock64() - t_start; prof[11] = lt; prof[12] = w_ld; prof[13] = w_stage; prof[14] = w_vec; }
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
