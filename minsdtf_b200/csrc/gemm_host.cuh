// gemm_host.cuh — host-side launcher for conv_gemm_kernel: picks the 128-pixel tile shape and BN, builds the
// TMA tensor maps and launches.  One call = one convolution / Dense layer of the reference graph.
#pragma once
#include "common.cuh"
#include "gemm.cuh"
#include "gemm2.cuh"
#include <stdlib.h>

namespace sdtf {

// Packed weight as the kernel wants it: [taps][N][K] bf16 (K contiguous), plus fp32 bias.
struct PackedWeight {
  bf16* w = nullptr;
  float* bias = nullptr;  // may be null
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
  bool out_fp32 = false;
  int act = ACT_NONE;
  float out_scale = 1.f;
  int force_bn = 0;  // testing / tuning
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
  if (act == ACT_GEGLU) return 160;
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

inline int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    SDTF_CUDA(cudaGetDevice(&dev));
    SDTF_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}
inline bool gemm_v1_forced() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SDTF_GEMM_V1");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
static constexpr size_t kSmemLimit = 232448;  // 227 KB per CTA on sm_100

// called once per process before any launch (and before any stream capture)
inline void init_gemm_kernels() {
  SDTF_CUDA(cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  SDTF_CUDA(cudaFuncSetAttribute(conv_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
  sm_count();
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
  p.BN = a.force_bn ? a.force_bn : choose_bn(w.N, m_tiles, a.act);
  if (a.act == ACT_GEGLU) SDTF_CHECK(w.geglu_half * 2 == p.BN && w.N % p.BN == 0, "GEGLU weight packing must match BN");
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
  const int iters = p.taps * (p.kc0 + p.kc1);
  if (stages > iters) stages = iters < 1 ? 1 : iters;
  p.stages = stages;
  p.bias = w.bias;
  p.temb = a.temb; p.temb_ld = a.temb_ld;
  p.res = a.res; p.res_ld = a.res_ld;
  p.out = a.out; p.out_ld = a.out_ld; p.out_fp32 = a.out_fp32 ? 1 : 0;
  p.act = a.act;
  p.out_scale = a.out_scale;

  CUtensorMap tmA0 = make_act_tmap(a.a0, p.bw, p.bh, p.bn, a.stride);
  CUtensorMap tmA1 = a.a1.p ? make_act_tmap(a.a1, p.bw, p.bh, p.bn, a.stride) : tmA0;
  CUtensorMap tmB = make_weight_tmap(w.w, w.K, w.N, p.taps, p.BN);

  // ---- persistent kernel (gemm2.cuh) whenever the output is bf16 and 16-byte chunkable ----
  {
    const bool geglu = a.act == ACT_GEGLU;
    const int ncols = geglu ? p.BN / 2 : p.BN;
    const int Nout = geglu ? w.N / 2 : w.N;
    Gemm2Extra x{};
    x.W = ncols <= 128 ? ncols : ncols / 2;
    x.passes = ncols <= 128 ? 1 : 2;
    const bool aligned = Nout % 8 == 0 && a.out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 &&
                         (a.res == nullptr || (a.res_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a.res) & 15) == 0));
    if (!a.out_fp32 && !a.force_v1 && !gemm_v1_forced() && aligned && x.W % 16 == 0 && x.W * x.passes == ncols) {
      x.m_tiles = (int)m_tiles;
      x.n_tiles = ceil_div(w.N, p.BN);
      x.pitch = x.W * 2 + 16;
      int acc = 32;
      while (acc < p.BN) acc <<= 1;
      x.acc_stride = acc;
      p.tmem_cols = 2 * acc;
      const size_t fixed = 1024 + 2 * (size_t)128 * x.pitch + 1024 + 8 * (2 * 8 + 5) + 16;
      int st2 = (int)((kSmemLimit - fixed) / stage_bytes);
      if (st2 > 8) st2 = 8;
      SDTF_CHECK(st2 >= 2, "gemm2: tile does not fit shared memory");
      p.stages = st2;
      const size_t smem2 = 1024 + (size_t)st2 * stage_bytes + 2 * (size_t)128 * x.pitch + 1024 + 8 * (2 * st2 + 5) + 16;
      const long long total = (long long)x.m_tiles * x.n_tiles;
      const unsigned grid2 = (unsigned)(total < sm_count() ? total : sm_count());
      conv_gemm2_kernel<<<grid2, kG2Threads, smem2, stream>>>(tmA0, tmA1, tmB, p, x);
      SDTF_CUDA(cudaGetLastError());
      return;
    }
  }
  const size_t smem = conv_smem_bytes(p.BN, p.stages);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    SDTF_CUDA(cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    smem_set = 200 * 1024;
  }
  dim3 grid((unsigned)m_tiles, (unsigned)ceil_div(w.N, p.BN), 1);
  conv_gemm_kernel<<<grid, 128, smem, stream>>>(tmA0, tmA1, tmB, p);
  SDTF_CUDA(cudaGetLastError());
}

}  // namespace sdtf
