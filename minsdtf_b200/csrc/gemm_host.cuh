// gemm_host.cuh — host-side launcher for conv_gemm_kernel: picks the 128-pixel tile shape and BN, builds the
// TMA tensor maps and launches.  One call = one convolution / Dense layer of the reference graph.
#pragma once
#include "common.cuh"
#include "gemm.cuh"
#include "gemm3.cuh"
#include <stdlib.h>

#include <vector>

namespace sdtf {

// Packed weight as the kernel wants it: [taps][N][K] bf16 (K contiguous), plus fp32 bias.
struct PackedWeight {
  bf16* w = nullptr;
  float* bias = nullptr;  // may be null
  float* ln_c1 = nullptr; // LayerNorm folded in: rows carry gamma, bias = c0, ln_c1[n] = sum_k of the packed row (WeightStore::fold_ln)
  int ln_channels = 0;    // channels the folded LayerNorm normalises (= true K)
  int K = 0, N = 0, kh = 1, kw = 1;
  int geglu_half = 0;  // >0: rows are interleaved [value x half | gate x half] per N tile of 2*half
};

struct ConvArgs {
  View a0, a1;  // a1.p == nullptr => single source; K = a0.C + a1.C
  const PackedWeight* w = nullptr;
  int stride = 1;
  int pad_t = 0, pad_l = 0;
  int outH = 0, outW = 0;  // output spatial size (batch = a0.B)
  // epilogue
  const float* temb = nullptr;
  int temb_ld = 0;
  const bf16* res = nullptr;
  long long res_ld = 0;
  void* out = nullptr;
  long long out_ld = 0;
  int out_head_d = 0, out_head_stride = 0;  // > 0: output columns are heads `out_head_stride` apart, only the first `out_head_d` are stored
  // LayerNorm folded into the surrounding GEMMs (gemm.cuh GemmParams::ln_*)
  float2* ln_out = nullptr;      // producer: per-row partial (sum, sum of squares) slots of the output, [slots][pixels]
  int* ln_slots_out = nullptr;   //           number of slots this launch filled (one per 32 output columns)
  const float2* ln_in = nullptr; // consumer: partials of its A operand's rows (weights must have been packed with fold_ln)
  int ln_slots = 0;
  int batch_class = 0;  // samples of the whole denoise step when this call only sees part of them (0: a0.B); see plan_splitk
  int out_step = 1;  // > 1: output pixel (y, x) of this launch is pixel (step y, step x) of the tensor at `out` (see upconv)
  bool out_fp32 = false;
  int act = ACT_NONE;
  float out_scale = 1.f;
  int force_bn = 0;  // testing / tuning
  float* splitk_ws = nullptr;  // split-K scratch: conv_splitk_plan(a).splits * pixels * N floats (null: no split-K)
  bool force_v1 = false;  // one-tile-per-CTA kernel (gemm.cuh) instead of the persistent one (gemm2.cuh)
};

struct TileShape {
  int bw, bh, bn;
};

// choose bw*bh*bn = 128 (powers of two) minimising the number of 128-pixel tiles
inline TileShape choose_tile(int W, int H, int B) {
  TileShape best{128, 1, 1};
  long long best_tiles = -1;
  for (int bw = 128; bw >= 1; bw >>= 1) {
    for (int bh = 128 / bw; bh >= 1; bh >>= 1) {
      int bn = 128 / (bw * bh);
      long long tiles = (long long)ceil_div(W, bw) * ceil_div(H, bh) * ceil_div(B, bn);
      // prefer wide rows (contiguous pixels) on ties
      if (best_tiles < 0 || tiles < best_tiles) {
        best_tiles = tiles;
        best = {bw, bh, bn};
      }
    }
  }
  return best;
}

inline int choose_bn(int N, long long m_tiles, int act) {
  if (act == ACT_GEGLU) return 256;
  if (N <= 16) return 16;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N % 160 == 0) {
    // 148 SMs x 2 resident CTAs: drop to 80-wide tiles when the grid would not fill the machine
    if (m_tiles * (N / 160) < 148 && N % 80 == 0) return 80;
    return 160;
  }
  if (N % 128 == 0) return 128;
  if (N % 96 == 0) return 96;
  if (N <= 128) return ((N + 15) / 16) * 16;
  return 128;  // ragged tail masked in the epilogue
}

inline size_t conv_smem_bytes(int BN, int stages) {
  return 1024 + (size_t)stages * (kATileBytes + (size_t)BN * 128) + 8 * (2 * stages + 1) + 16;
}

inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}
// SDTF_GEMM = 1: one-tile-per-CTA kernel only (gemm.cuh); 3 (default): persistent TMA-epilogue kernel (gemm3.cuh)
inline int gemm_version() {
  static int v = env_int("SDTF_GEMM", 3);
  return v;
}
inline bool gemm_v1_forced() { return gemm_version() == 1; }
// tuning overrides for experiments: SDTF_GEMM_CG = 1|2 forces the CTA-group size, SDTF_GEMM_BN the N tile
inline int gemm_force_cg() {
  static int v = env_int("SDTF_GEMM_CG", 0);
  return v;
}
inline int gemm_force_bn() {
  static int v = env_int("SDTF_GEMM_BN", 0);
  return v;
}

struct G3Plan {
  int cg = 1, bn = 0, n_mma = 1, bufs = 2;
};
// Pick (CTA-group size, N tile) for the persistent kernel with a small cost model.  One k-iteration (64 channels of
// one tap) costs max(tensor cycles = 2*BN, operand bytes / L2->SM delivery rate); measured delivery is ~50 B/clk/SM
// with every SM pulling (ncu: 256x160 tiles keep the tensor pipe 50 % busy, 256x256 tiles 80 %).  Tiles wider than
// 256 columns are two MMAs per k-step and a single-buffered accumulator: their epilogue is not hidden.
inline G3Plan plan_gemm3(int N, long long m_tiles, int act, int force_bn, int iters, bool whole_passes = false) {
  G3Plan best;
  int cands[40], nc = 0;
  const bool geglu = act == ACT_GEGLU;
  if (geglu) cands[nc++] = 256;
  else if (force_bn) cands[nc++] = force_bn;
  else if (gemm_force_bn() && N % gemm_force_bn() == 0) cands[nc++] = gemm_force_bn();
  else if (N <= 16) cands[nc++] = 16;
  else if (N <= 32) cands[nc++] = 32;
  else if (N <= 64) cands[nc++] = 64;
  else {
    for (int bn = 512; bn >= 64; bn -= 32)
      if (N % bn == 0 && (bn <= 256 || bn % 64 == 0)) cands[nc++] = bn;
    if (N % 80 == 0) cands[nc++] = 80;
    if (nc == 0) cands[nc++] = N <= 128 ? ((N + 15) / 16) * 16 : 128;  // ragged tail clipped by the TMA store
  }
  double best_cost = 1e30;
  for (int pass = 0; pass < 2 && best.bn == 0; ++pass)
    for (int i = 0; i < nc; ++i) {
      const int bn = cands[i];
      if (whole_passes && bn % 32 != 0 && nc > 1) continue;  // LayerNorm statistics: no 32-column pass may overlap another
      const long long n_tiles = (N + bn - 1) / bn;
      const int n_mma = bn > 256 ? 2 : 1;
      for (int cg = 1; cg <= 2; ++cg) {
        if (pass == 0 && gemm_force_cg() && cg != gemm_force_cg()) continue;
        if (cg == 2 && (m_tiles < 2 || (bn / n_mma / 2) % 8 != 0)) continue;
        if ((bn / n_mma) % 16 != 0) continue;
        const size_t stage = 16384 + (size_t)bn * 128 / cg;
        if (3 * stage + 72000 > 232448) continue;  // at least 3 stages beside the staging buffers
        const long long units = ((m_tiles + cg - 1) / cg) * n_tiles;
        const long long slots = 148 / cg;
        const long long waves = (units + slots - 1) / slots;
        const double busy = (double)(units < slots ? units : slots) * cg / 148.0;  // fraction of SMs pulling from L2
        const double rate = 50.0 / (busy > 0.3 ? busy : 0.3);
        const double mma = 2.0 * bn;
        const double l2 = (double)stage / rate;
        const double t_iter = (mma > l2 ? mma : l2) + 30.0;
        const int ncols = geglu ? bn / 2 : bn;
        const double t_epi = 400.0 + ((ncols + 31) / 32) * 450.0;
        const double t_main = iters * t_iter;
        const double t_unit = bn > 256 ? t_main + t_epi : (t_main > t_epi ? t_main : t_epi) + 200.0;
        const double cost = (double)waves * t_unit * (cg == 2 && m_tiles % 2 ? 1.02 : 1.0);
        if (cost < best_cost) { best_cost = cost; best.cg = cg; best.bn = bn; best.n_mma = n_mma; best.bufs = bn > 256 ? 1 : 2; }
      }
    }
  return best;
}
static constexpr size_t kSmemLimit = 232448;  // 227 KB per CTA on sm_100

// ---- split-K ----
// Layers with few output tiles and a long K loop (the 8x8 level at batch 16: 1024 pixels x 1280 channels, K = 9 x 1280
// or 9 x 2560) fill 80 of 148 SMs, and each CTA is bound by L2->SM operand delivery (~46 B/clk/SM): 54 us for a conv
// that is 20 us of tensor work.  Their k-iterations are cut into `splits` ranges, one unit each, with wide tiles
// (BN 320 / 256: fewer operand bytes per FLOP); fp32 partial sums go to scratch and splitk_reduce_kernel adds them in
// index order (deterministic) and applies the usual epilogue once.
struct SplitKPlan {
  int splits = 1, cg = 1, bn = 0, n_mma = 1, iters_split = 0;
};
inline int splitk_enabled() {
  static int v = env_int("SDTF_SPLITK", 1);
  return v;
}
// The decision and the K partition depend on the layer (output pixels per sample, N, K) and on the BATCH CLASS only
// (small: <= kSmallBatch samples per call, i.e. one or two prompts with their CFG twins; large: anything above) — never
// on the batch size inside a class: a sample's result must not depend on what it is batched with (fp32 partial sums
// round differently from one accumulator), see test_full_size_batching_and_determinism.  Rules:
//   * the 8x8 level and below (<= 64 output pixels per sample): always, 4 ranges;
//   * the 16x16 level (<= 256 pixels per sample) in the small-batch class only: at UNet batch 2 those convs are 4 M tiles
//     x 180 k-iterations — 32 CTAs walking a 60 us serial K loop; with 4 ranges 128 CTAs share it;
//   * spatial filters only (the linears run on token-shaped views that fold the batch into the pixel count), >= 60
//     k-iterations.
// The N tile does NOT change any sum (every output element sees the same k order), so it is free to follow the batch:
// the widest tile that still yields >= 96 CTAs (fewer operand bytes per FLOP), down to 64 columns when one M tile is all
// there is (UNet batch 2 at 8x8: 80 CTAs stream the 29 MB of weights instead of 16).
static constexpr int kSmallBatch = 4;
inline SplitKPlan plan_splitk(int N, long long m_tiles, int iters, int act, bool bf16_out, long long pixels_per_sample, int taps,
                              int batch) {
  SplitKPlan pl;
  if (!splitk_enabled() || !bf16_out || (act != ACT_NONE && act != ACT_SILU) || iters < 60 || N % 8 != 0) return pl;
  if (taps == 1) return pl;
  static const int small_on = env_int("SDTF_SPLITK_SMALL", 1);  // A/B: 0 = round-1 rule (8x8 level only)
  const bool small = small_on && batch <= kSmallBatch;
  if (pixels_per_sample > (small ? 256 : 64)) return pl;
  static const int cands[5] = {320, 256, 160, 128, 64};
  static const int force_bn = env_int("SDTF_SPLITK_BN", 0), n_splits = env_int("SDTF_SPLITK_N", 4);  // tuning
  const int cg = m_tiles >= 2 ? 2 : 1;
  const int iters_split = (iters + n_splits - 1) / n_splits;
  const int splits = (iters + iters_split - 1) / iters_split;
  int pick = 0;
  for (int bn : cands) {
    if (N % bn || (force_bn && bn != force_bn)) continue;
    const int n_mma = bn > 256 ? 2 : 1;
    if ((bn / n_mma) % 16 != 0 || (cg == 2 && (bn / n_mma / 2) % 8 != 0)) continue;
    pick = bn;  // the narrowest valid tile so far; stop at the widest one that fills the machine
    const long long ctas = ((m_tiles + cg - 1) / cg) * (N / bn) * splits * cg;
    if (ctas >= 96 || !small_on) break;
  }
  if (!pick) return pl;
  pl.iters_split = iters_split;
  pl.splits = splits;
  pl.cg = cg; pl.bn = pick; pl.n_mma = pick > 256 ? 2 : 1;
  return pl;
}

__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, long long M, int N, long long rows_per_b,
                     const float* __restrict__ bias, const float* __restrict__ temb, int temb_ld, const bf16* __restrict__ res,
                     long long res_ld, bf16* __restrict__ out, long long out_ld, int act, float out_scale) {
  pdl_trigger();
  pdl_wait();
  const int nv = N >> 2;
  const long long total = M * nv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / nv;
    const int c = (int)(i - m * nv) << 2;
    float4 a = __ldcs(reinterpret_cast<const float4*>(partial + m * N + c));
    for (int s = 1; s < splits; ++s) {
      const float4 b = __ldcs(reinterpret_cast<const float4*>(partial + ((long long)s * M + m) * N + c));
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    a.x *= out_scale; a.y *= out_scale; a.z *= out_scale; a.w *= out_scale;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) v = __ldg(reinterpret_cast<const float4*>(bias + c));
    if (temb) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(temb + (m / rows_per_b) * temb_ld + c));
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    if (res) {
      const uint2 r = *reinterpret_cast<const uint2*>(res + m * res_ld + c);
      const __nv_bfloat162 r0 = *reinterpret_cast<const __nv_bfloat162*>(&r.x), r1 = *reinterpret_cast<const __nv_bfloat162*>(&r.y);
      a.x += __bfloat162float(r0.x); a.y += __bfloat162float(r0.y); a.z += __bfloat162float(r1.x); a.w += __bfloat162float(r1.y);
    }
    if (act == ACT_SILU) { a.x = silu_f(a.x); a.y = silu_f(a.y); a.z = silu_f(a.z); a.w = silu_f(a.w); }
    uint2 o;
    o.x = tc05::pack_bf16(a.x, a.y);
    o.y = tc05::pack_bf16(a.z, a.w);
    *reinterpret_cast<uint2*>(out + m * out_ld + c) = o;
  }
}

// called once per process before any launch (and before any stream capture)
inline void init_gemm_kernels() {
  SDTF_CUDA(cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  SDTF_CUDA(cudaFuncSetAttribute(conv_gemm3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  SDTF_CUDA(cudaFuncSetAttribute(conv_gemm3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  SDTF_CUDA(cudaFuncSetAttribute(conv_gemm3_kernel<1, G3_LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  SDTF_CUDA(cudaFuncSetAttribute(conv_gemm3_kernel<2, G3_LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  SDTF_CUDA(cudaFuncSetAttribute(conv_gemm3_kernel<1, G3_GEGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  SDTF_CUDA(cudaFuncSetAttribute(conv_gemm3_kernel<2, G3_GEGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  SDTF_CUDA(cudaFuncSetAttribute(conv_gemm3_kernel<1, G3_LN_PRODUCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  SDTF_CUDA(cudaFuncSetAttribute(conv_gemm3_kernel<2, G3_LN_PRODUCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  SDTF_CUDA(cudaFuncSetAttribute(conv_gemm3_kernel<1, G3_GEGLU_LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  SDTF_CUDA(cudaFuncSetAttribute(conv_gemm3_kernel<2, G3_GEGLU_LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  SDTF_CUDA(cudaFuncSetAttribute(conv_gemm3_kernel<1, G3_LN_APPLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  SDTF_CUDA(cudaFuncSetAttribute(conv_gemm3_kernel<2, G3_LN_APPLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  sm_count();
}

// floats of split-K scratch launch_conv would use for this conv (0: the layer is not split)
inline size_t conv_splitk_floats(const ConvArgs& a) {
  const PackedWeight& w = *a.w;
  if (a.out_fp32 || a.force_v1 || a.force_bn || gemm_v1_forced() || a.out_step != 1) return 0;
  const TileShape ts = choose_tile(a.outW, a.outH, a.a0.B);
  const long long m_tiles = (long long)ceil_div(a.outW, ts.bw) * ceil_div(a.outH, ts.bh) * ceil_div(a.a0.B, ts.bn);
  const int iters = w.kh * w.kw * (ceil_div(a.a0.C, 64) + (a.a1.p ? ceil_div(a.a1.C, 64) : 0));
  const SplitKPlan sk = plan_splitk(w.N, m_tiles, iters, a.act, true, (long long)a.outH * a.outW, w.kh * w.kw, a.batch_class ? a.batch_class : a.a0.B);
  return sk.splits > 1 ? (size_t)sk.splits * a.a0.B * a.outH * a.outW * w.N : 0;
}

inline void launch_conv(cudaStream_t stream, const ConvArgs& a) {
  const PackedWeight& w = *a.w;
  const int K = a.a0.C + (a.a1.p ? a.a1.C : 0);
  SDTF_CHECK(K == w.K || (a.a1.p == nullptr && ((K + 7) / 8) * 8 == w.K), "conv: K mismatch between activation and weight");
  SDTF_CHECK(a.a1.p == nullptr || a.a0.C % 64 == 0, "conv: first concat source must be a multiple of 64 channels");
  const int B = a.a0.B;
  GemmParams p{};
  p.W = a.outW; p.H = a.outH; p.B = B;
  TileShape ts = choose_tile(p.W, p.H, p.B);
  p.bw = ts.bw; p.bh = ts.bh; p.bn = ts.bn;
  p.tiles_x = ceil_div(p.W, p.bw);
  p.tiles_y = ceil_div(p.H, p.bh);
  const int tiles_b = ceil_div(p.B, p.bn);
  const long long m_tiles = (long long)p.tiles_x * p.tiles_y * tiles_b;
  p.taps = w.kh * w.kw; p.tap_w = w.kw;
  p.pad_x = a.pad_l; p.pad_y = a.pad_t; p.stride = a.stride;
  p.kc0 = ceil_div(a.a0.C, 64);
  p.kc1 = a.a1.p ? ceil_div(a.a1.C, 64) : 0;
  p.N = w.N;
  p.bias = w.bias;
  p.temb = a.temb; p.temb_ld = a.temb_ld;
  p.res = a.res; p.res_ld = a.res_ld;
  p.out = a.out; p.out_ld = a.out_ld; p.out_fp32 = a.out_fp32 ? 1 : 0;
  p.act = a.act;
  p.out_scale = a.out_scale;
  p.ln_out = nullptr; p.ln_in = nullptr; p.ln_c1 = nullptr; p.ln_slots = 0; p.ln_rows = (long long)B * a.outH * a.outW; p.ln_inv_c = 0.f;
  const bool geglu = a.act == ACT_GEGLU;
  const int Nout = geglu ? w.N / 2 : w.N;
  const int iters = p.taps * (p.kc0 + p.kc1);

  // ---- persistent kernel (gemm3.cuh) whenever the output is bf16 and TMA-addressable ----
  const bool aligned = Nout % 8 == 0 && a.out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 &&
                       (a.res == nullptr || (a.res_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a.res) & 15) == 0));
  if (!a.out_fp32 && !a.force_v1 && !gemm_v1_forced() && aligned) {
    G3Plan plan = plan_gemm3(w.N, m_tiles, a.act, a.force_bn, iters, a.ln_out != nullptr);
    SplitKPlan sk;
    if (a.splitk_ws && !a.force_bn) sk = plan_splitk(w.N, m_tiles, iters, a.act, true, (long long)a.outH * a.outW, w.kh * w.kw, a.batch_class ? a.batch_class : B);
    const bool split = sk.splits > 1;
    if (split) { plan.cg = sk.cg; plan.bn = sk.bn; plan.n_mma = sk.n_mma; plan.bufs = sk.bn > 256 ? 1 : 2; }
    p.BN = plan.bn;
    const int ncols = geglu ? p.BN / 2 : p.BN;
    const int n_tiles = ceil_div(w.N, p.BN);
    // a last pass narrower than 32 columns is shifted back over columns the tile already wrote: fine unless the
    // residual is read from the tensor being written
    const bool overlap_ok = ncols % 32 == 0 || ncols < 32 || a.res != a.out;
    const bool vec_ok = split || a.temb == nullptr || p.bn <= 4;  // the staged epilogue vector holds up to 4 sample rows
    if (vec_ok && (ncols % 32 == 0 || (ncols < 32 && n_tiles == 1) || (ncols > 32 && overlap_ok))) {
      if (split) {  // partial sums only: bias / time embedding / residual / activation move to the reduce kernel
        p.bias = nullptr; p.temb = nullptr; p.res = nullptr; p.act = ACT_NONE; p.out_scale = 1.f;
      }
      if (geglu) SDTF_CHECK(w.geglu_half * 2 == p.BN && w.N % p.BN == 0, "GEGLU weight packing must match BN");
      if (a.ln_out) {
        SDTF_CHECK(!split && a.out_step == 1 && a.out_head_stride == 0 && !geglu && ncols % 32 == 0 && Nout % 32 == 0,
                   "LayerNorm statistics: plain bf16 output in whole 32-column passes only");
        p.ln_out = a.ln_out;
        if (a.ln_slots_out) *a.ln_slots_out = Nout / 32;
      }
      if (a.ln_in) {
        SDTF_CHECK(w.ln_c1 != nullptr && w.ln_channels > 0 && a.ln_slots > 0 && !split && a.temb == nullptr && a.out_scale == 1.f,
                   "LayerNorm-folded GEMM: weights must be packed with fold_ln, statistics supplied");
        p.ln_in = a.ln_in; p.ln_slots = a.ln_slots; p.ln_c1 = w.ln_c1; p.ln_inv_c = 1.f / (float)w.ln_channels;
      }
      Gemm3Extra x{};
      x.m_tiles = (int)m_tiles;
      x.n_tiles = n_tiles;
      x.ncols = ncols;
      int acc = 32;
      while (acc < p.BN) acc <<= 1;
      x.acc_stride = acc;
      x.acc_bufs = plan.bufs;
      x.n_mma = plan.n_mma;
      p.tmem_cols = plan.bufs * acc;
      SDTF_CHECK(p.tmem_cols <= 512, "gemm3: accumulators exceed TMEM");
      int lg = 0;
      while ((1 << lg) < p.bw * p.bh) ++lg;
      x.log_rows_per_b = lg;
      int lbw = 0;
      while ((1 << lbw) < p.bw) ++lbw;
      SDTF_CHECK((1 << lbw) == p.bw, "gemm3: tile width must be a power of two");
      x.log_bw = lbw;
      x.splits = split ? sk.splits : 1;
      x.iters_split = split ? sk.iters_split : iters;
      x.partial = split ? a.splitk_ws : nullptr;
      x.vec_rows = a.ln_in ? 2 : ((a.temb && !split) ? p.bn : 1);
      x.vec_width = ((geglu ? p.BN : ncols) + 31) / 32 * 32;
      const size_t vec_bytes = (size_t)2 * 2 * x.vec_rows * x.vec_width * 4;  // per epilogue warp set, double buffered
      const int cg = plan.cg;
      const size_t stage_bytes = kATileBytes + (size_t)(p.BN / cg) * 128;
      x.nbufs = (a.res && !split) ? kG3MaxBufs : 4;  // residual tiles are TMA-prefetched nbufs-1 passes ahead: latency needs depth
      // (eight buffers without a residual, where they cost no ring stage: no change — 64.0 against 64.1 us at 320 -> 1536)
      const int kG3Bufs = x.nbufs;
      const size_t fixed = 1024 + (size_t)kG3Bufs * kG3BufBytes + 8 * (2 * 10 + 4 + 2 * kG3Bufs) + 32 + vec_bytes;
      int st = (int)((kSmemLimit - fixed) / stage_bytes);
      if (st > 10) st = 10;
      if (st > x.iters_split) st = x.iters_split < 2 ? 2 : x.iters_split;
      SDTF_CHECK(st >= 2, "gemm3: tile does not fit shared memory");
      p.stages = st;
      const size_t smem = 1024 + (size_t)st * stage_bytes + (size_t)kG3Bufs * kG3BufBytes + 8 * (2 * st + 4 + 2 * kG3Bufs) + 32 + vec_bytes;
      CUtensorMap tmA0 = make_act_tmap(a.a0, p.bw, p.bh, p.bn, a.stride);
      CUtensorMap tmA1 = a.a1.p ? make_act_tmap(a.a1, p.bw, p.bh, p.bn, a.stride) : tmA0;
      CUtensorMap tmB = make_weight_tmap(w.w, w.K, w.N, p.taps, p.BN / plan.n_mma / cg);
      {
        const long long m_units = (m_tiles + cg - 1) / cg, total = m_units * n_tiles * x.splits;
        long long dmax = n_tiles > m_units ? n_tiles : m_units;
        if (p.tiles_x > dmax) dmax = p.tiles_x;
        if (p.tiles_y > dmax) dmax = p.tiles_y;
        SDTF_CHECK((total + 2 * cg) * dmax < 0x100000000ll, "gemm3: tile count too large for the multiply-high tile decode");
        x.d_ntiles = FastDiv::make(n_tiles); x.d_munits = FastDiv::make((int)m_units);
        x.d_tx = FastDiv::make(p.tiles_x); x.d_ty = FastDiv::make(p.tiles_y);
      }
      x.head_stride = 0;
      CUtensorMap tmOut;
      if (a.out_head_stride > 0) {
        SDTF_CHECK(a.out_head_stride % 32 == 0 && ncols % 32 == 0 && Nout % a.out_head_stride == 0 && a.out_step == 1 && !split && !a.res,
                   "conv: head-clipped output needs 32-column passes that do not straddle heads");
        x.head_stride = a.out_head_stride;
        tmOut = make_epi_tmap_heads(reinterpret_cast<const bf16*>(a.out), a.out_head_d, a.out_head_stride, Nout / a.out_head_stride, p.W, p.H,
                                    p.B, a.out_ld, p.bw, p.bh, p.bn);
      } else {
        tmOut = make_epi_tmap(reinterpret_cast<const bf16*>(a.out), Nout, p.W, p.H, p.B, a.out_ld, p.bw, p.bh, p.bn, a.out_step);
      }
      CUtensorMap tmRes = (a.res && !split) ? make_epi_tmap(a.res, Nout, p.W, p.H, p.B, a.res_ld, p.bw, p.bh, p.bn) : tmOut;
      const long long units = ((m_tiles + cg - 1) / cg) * n_tiles * x.splits;
      const long long slots = sm_count() / cg;
      const unsigned grid = (unsigned)((units < slots ? units : slots) * cg);
      // SDTF_GEMM_PROFILE=1: per-role wait-cycle accounting, printed after the launch (debug; synchronises)
      static const int profile = env_int("SDTF_GEMM_PROFILE", 0);
      static const int debug = env_int("SDTF_GEMM_DEBUG", 0);
      x.debug = debug;
      static long long* prof_buf = nullptr;
      if (profile) {
        if (!prof_buf) SDTF_CUDA(cudaMalloc((void**)&prof_buf, sizeof(long long) * 16 * 160));
        SDTF_CUDA(cudaMemsetAsync(prof_buf, 0, sizeof(long long) * 16 * 160, stream));
        x.prof = prof_buf;
      }
      // one instantiation per epilogue variant (gemm3.cuh MODE); SDTF_GEMM_LEAN=0 (A/B): every launch on the general kernel
      static const int lean_on = env_int("SDTF_GEMM_LEAN", 1);
      const bool plain_act = a.act == ACT_NONE || a.act == ACT_SILU;
      int mode = G3_GENERAL;
      if (lean_on && !split && !profile) {
        if (geglu && !a.ln_in && !a.ln_out) mode = G3_GEGLU;
        else if (geglu && a.ln_in && !a.ln_out) mode = G3_GEGLU_LN;
        else if (plain_act && a.ln_out && !a.ln_in) mode = G3_LN_PRODUCE;
        else if (plain_act && a.ln_in && !a.ln_out) mode = G3_LN_APPLY;
        else if (plain_act && !a.ln_in && !a.ln_out) mode = G3_LEAN;
      }
#define SDTF_G3_LAUNCH(M)                                                                                                        \
  do {                                                                                                                           \
    if (cg == 2) launch_pdl(conv_gemm3_kernel<2, M>, dim3(grid), dim3(kG3Threads), smem, stream, 2, tmA0, tmA1, tmB, tmOut, tmRes, p, x); \
    else launch_pdl(conv_gemm3_kernel<1, M>, dim3(grid), dim3(kG3Threads), smem, stream, 1, tmA0, tmA1, tmB, tmOut, tmRes, p, x);        \
  } while (0)
      switch (mode) {
        case G3_LEAN: SDTF_G3_LAUNCH(G3_LEAN); break;
        case G3_GEGLU: SDTF_G3_LAUNCH(G3_GEGLU); break;
        case G3_GEGLU_LN: SDTF_G3_LAUNCH(G3_GEGLU_LN); break;
        case G3_LN_PRODUCE: SDTF_G3_LAUNCH(G3_LN_PRODUCE); break;
        case G3_LN_APPLY: SDTF_G3_LAUNCH(G3_LN_APPLY); break;
        default: SDTF_G3_LAUNCH(G3_GENERAL); break;
      }
#undef SDTF_G3_LAUNCH
      SDTF_CUDA(cudaGetLastError());
      if (split) {
        const long long M = (long long)p.B * p.H * p.W;
        long long blocks = ceil_div_ll(M * (w.N / 4), 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        launch_pdl(splitk_reduce_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, 1, (const float*)a.splitk_ws, x.splits, M, w.N,
                   (long long)p.H * p.W, w.bias, a.temb, a.temb_ld, a.res, a.res_ld, reinterpret_cast<bf16*>(a.out), a.out_ld, a.act,
                   a.out_scale);
      }
      if (profile) {
        SDTF_CUDA(cudaStreamSynchronize(stream));
        std::vector<long long> h(16 * 160);
        SDTF_CUDA(cudaMemcpy(h.data(), prof_buf, sizeof(long long) * 16 * 160, cudaMemcpyDeviceToHost));
        double a[16] = {0};
        for (unsigned b = 0; b < grid; ++b)
          for (int k = 0; k < 16; ++k) a[k] += (double)h[b * 16 + k] / grid;
        fprintf(stderr,
                "[gemm3-prof] N %d iters %d CG %d BN %d | producer wait_empty %.0f / %.0f | mma wait_full %.0f wait_tempty %.0f / %.0f | "
                "store wait_stg %.0f wait_read %.0f / %.0f | epi wait_tfull %.0f wait_grant %.0f tmem_ld %.0f pack+stage+fence %.0f vec+bar %.0f / %.0f tiles "
                "%.1f (cycles, avg per CTA; mma counters are per leader CTA x1/CG)\n",
                w.N, iters, cg, p.BN, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[12], a[13], a[14], a[10], a[11]);
      }
      static const int verbose = env_int("SDTF_GEMM_VERBOSE", 0);
      if (verbose)
        fprintf(stderr, "[gemm3] M-tiles %lld N %d K-iters %d -> CG %d BN %d (x%d MMA, %d acc) stages %d grid %u split-K %d\n", m_tiles, w.N,
                iters, cg, p.BN, plan.n_mma, plan.bufs, st, grid, x.splits);
      return;
    }
  }

  // ---- one tile per CTA (gemm.cuh): fp32 outputs, unaligned views ----
  SDTF_CHECK(a.ln_in == nullptr && a.ln_out == nullptr, "conv: LayerNorm folding needs the TMA-store kernel");
  SDTF_CHECK(a.out_step == 1 && a.out_head_stride == 0, "conv: strided / head-clipped output needs the TMA-store kernel (bf16, 16-byte aligned output)");
  p.BN = a.force_bn ? a.force_bn : choose_bn(w.N, m_tiles, a.act);
  if (geglu) SDTF_CHECK(w.geglu_half * 2 == p.BN && w.N % p.BN == 0, "GEGLU weight packing must match BN");
  SDTF_CHECK(p.BN % 16 == 0 && p.BN >= 16 && p.BN <= 256, "BN must be a multiple of 16 in [16,256]");
  int cols = 32;
  while (cols < p.BN) cols <<= 1;
  p.tmem_cols = cols;
  // stages: as deep as fits ~110 KB so that two CTAs stay resident per SM (the second one's main loop
  // overlaps the first one's epilogue)
  const size_t stage_bytes = kATileBytes + (size_t)p.BN * 128;
  int stages = (int)((110 * 1024) / stage_bytes);
  if (stages > 6) stages = 6;
  if (stages < 2) stages = 2;
  if (stages > iters) stages = iters < 1 ? 1 : iters;
  p.stages = stages;
  CUtensorMap tmA0 = make_act_tmap(a.a0, p.bw, p.bh, p.bn, a.stride);
  CUtensorMap tmA1 = a.a1.p ? make_act_tmap(a.a1, p.bw, p.bh, p.bn, a.stride) : tmA0;
  CUtensorMap tmB = make_weight_tmap(w.w, w.K, w.N, p.taps, p.BN);
  const size_t smem = conv_smem_bytes(p.BN, p.stages);
  dim3 grid((unsigned)m_tiles, (unsigned)ceil_div(w.N, p.BN), 1);
  conv_gemm_kernel<<<grid, 128, smem, stream>>>(tmA0, tmA1, tmB, p);
  SDTF_CUDA(cudaGetLastError());
}

}  // namespace sdtf
