// ops.cuh — the HBM-bound kernels of the path: GroupNorm(+SiLU), LayerNorm, nearest upsample, row softmax,
// the tiny time-embedding linears, the fused CFG + scheduler step, weight packing and dtype conversions.
// All use 128-bit loads/stores on NHWC bf16 and warp-shuffle reductions.
#pragma once
#include "common.cuh"

namespace sdtf {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __bfloat162float(h[i].x);
    f[2 * i + 1] = __bfloat162float(h[i].y);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162 h;
  h = __floats2bfloat162_rn(f[0], f[1]); u.x = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[2], f[3]); u.y = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[4], f[5]); u.z = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[6], f[7]); u.w = *reinterpret_cast<uint32_t*>(&h);
  return u;
}

// ------------------------------------------------------------------------------------------------------
// GroupNorm(32 groups, eps 1e-5, biased variance) [+ SiLU]   (reference: keras GroupNormalization at
// diffusion_model.py:27-28,32-33,57,277-278; layers.py:32,66-79; image_decoder.py:51-52)
// ONE kernel per GroupNorm.  A sample is cut into nblk contiguous pixel ranges, one CTA each (nblk depends on
// (H*W, C) only, so a sample's statistics do not depend on what it is batched with), and all CTAs of the launch are
// co-resident (grid <= 2 CTAs per SM; larger batches are walked in rounds by the same CTAs):
//   phase 1  per-thread sum / sum-of-squares of its 8-channel vector over the CTA's pixels (16-byte, warp-contiguous
//            loads, four in flight) -> smem -> fixed-order per-channel, then per-group reduction -> one fp64 partial
//            per (CTA, group), published to global memory;
//   barrier  per-sample ticket counter + generation word (sense reversal); every CTA of the sample then adds the nblk
//            partials in index order — DETERMINISTIC and identical in every CTA;
//   phase 2  y = silu?((x - mean_g) * rstd_g * gamma_c + beta_c) over the same pixel range: the re-read hits L2 (the
//            producer conv just wrote the tensor), so HBM sees one read and one write per element.
// The previous two-kernel version (statistics, then apply) paid two launches, a serial last-CTA reduction and a
// second cold ramp per GroupNorm: 3.2 ms of a 18.8 ms denoise step for 61 GroupNorms.
// ------------------------------------------------------------------------------------------------------
static constexpr int kGnMaxBlk = 128;

struct GnScratch {
  double* partial = nullptr;    // [B][kGnMaxBlk][64]
  unsigned* counters = nullptr; // [B], zero between launches (self-resetting)
  unsigned* gens = nullptr;     // [B], barrier generation (monotonic)
  float* stats = nullptr;       // [B][64]: mean, rstd per group (kept for debugging / tests)
  int chains = 4;               // partial-sum chains per statistic (SDTF_GN_CHAINS=1: the round-1 single chain, for A/B)
  int share = 1;                // launches of this scratch may run next to (share - 1) other GroupNorm kernels: each gets
                                // 1/share of the co-resident CTA slots (two spinning kernels must fit the device TOGETHER)
};

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(512, 2)
gn_fused_kernel(const bf16* __restrict__ x, long long ld, int C, long long HW, int pix_per_cta, int B, GnScratch sc,
                const float* __restrict__ gamma, const float* __restrict__ beta, int silu, bf16* __restrict__ y, long long ldy) {
  extern __shared__ __align__(16) float gn_sm[];  // sums [lanes][C], then squares [lanes][C]
  __shared__ float s_stats[64];
  __shared__ double s_chain[4][64];
  pdl_trigger();
  pdl_wait();  // x, and the barrier words of the previous GroupNorm
  const int vecs = C >> 3, gs = C >> 5;
  const int lanes = blockDim.x / vecs;  // pixels processed in parallel by this CTA
  const int cv = threadIdx.x % vecs, pl = threadIdx.x / vecs;
  const int nblk = gridDim.x;
  float* smS = gn_sm;
  float* smQ = gn_sm + lanes * C;
  const long long p0 = (long long)blockIdx.x * pix_per_cta;
  const long long p1 = min(HW, p0 + pix_per_cta);
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    unsigned gen0 = 0;
    if (threadIdx.x == 0) gen0 = ld_acquire_gpu_u32(sc.gens + b);  // cannot advance before this CTA has arrived
    // ---------------- phase 1 ----------------
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    const bf16* xb = x + ((long long)b * HW) * ld + cv * 8;
    long long p = p0 + pl;
    auto accum = [&](const uint4& u) {
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f[i];
        q[i] = fmaf(f[i], f[i], q[i]);
      }
    };
    for (; p + 7LL * lanes < p1; p += 8LL * lanes) {  // eight independent 16-byte loads in flight per thread
      uint4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] = *reinterpret_cast<const uint4*>(xb + (p + (long long)k * lanes) * ld);
#pragma unroll
      for (int k = 0; k < 8; ++k) accum(u[k]);
    }
    for (; p + 3LL * lanes < p1; p += 4LL * lanes) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = *reinterpret_cast<const uint4*>(xb + (p + (long long)k * lanes) * ld);
#pragma unroll
      for (int k = 0; k < 4; ++k) accum(u[k]);
    }
    for (; p < p1; p += lanes) accum(*reinterpret_cast<const uint4*>(xb + p * ld));
    {
      float4* dS = reinterpret_cast<float4*>(smS + pl * C + cv * 8);
      float4* dQ = reinterpret_cast<float4*>(smQ + pl * C + cv * 8);
      dS[0] = make_float4(s[0], s[1], s[2], s[3]); dS[1] = make_float4(s[4], s[5], s[6], s[7]);
      dQ[0] = make_float4(q[0], q[1], q[2], q[3]); dQ[1] = make_float4(q[4], q[5], q[6], q[7]);
    }
    __syncthreads();
    // per channel over the pixel lanes, fixed order (row 0 of each array receives the result)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float S = smS[c], Q = smQ[c];
      for (int l = 1; l < lanes; ++l) {
        S += smS[l * C + c];
        Q += smQ[l * C + c];
      }
      smS[c] = S;
      smQ[c] = Q;
    }
    __syncthreads();
    // per group: one warp per group at a time, strided partial sums + a fixed shuffle tree
    {
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
      for (int g = warp; g < 32; g += nwarps) {
        float S = 0.f, Q = 0.f;
        for (int c = lane; c < gs; c += 32) {
          S += smS[g * gs + c];
          Q += smQ[g * gs + c];
        }
        S = warp_sum(S);
        Q = warp_sum(Q);
        if (lane == 0) {
          double* dst = sc.partial + (((long long)b * kGnMaxBlk + blockIdx.x) * 32 + g) * 2;
          dst[0] = (double)S;
          dst[1] = (double)Q;
        }
      }
    }
    __threadfence();
    __syncthreads();
    // ---------------- per-sample barrier ----------------
    if (threadIdx.x == 0) {
      const unsigned ticket = atomicAdd(&sc.counters[b], 1u);
      if (ticket == (unsigned)nblk - 1) {
        sc.counters[b] = 0;
        __threadfence();
        atomicAdd(&sc.gens[b], 1u);
      } else {
        unsigned spins = 0;
        while (ld_acquire_gpu_u32(sc.gens + b) == gen0) {
          if (++spins > (1u << 28)) __trap();  // a CTA of this sample is not resident: launch configuration bug
        }
      }
      __threadfence();
    }
    __syncthreads();
    // every CTA of the sample adds the nblk partials — in a FIXED order (four interleaved chains of consecutive k, then the
    // four chain sums in index order), so all CTAs get identical statistics and the result depends on nblk only.  Four
    // chains with the loads issued ahead instead of one: a single chain was nblk dependent L2 round trips (64 x ~0.4 us
    // in the small-batch class — most of the kernel's 32 us there).
    if (threadIdx.x < 256) {  // (blockDim.x >= 256 for every channel count: launch_groupnorm)
      const int pair = threadIdx.x & 63, chain = threadIdx.x >> 6;
      const double* src = sc.partial + ((long long)b * kGnMaxBlk * 32) * 2 + pair;
      const int per = sc.chains == 1 ? (chain == 0 ? nblk : 0) : (nblk + 3) >> 2, k0 = chain * per, k1 = min(nblk, k0 + per);
      double acc = 0;
      int k = k0;
      for (; k + 4 <= k1; k += 4) {
        const double v0 = __ldcg(src + (long long)k * 64), v1 = __ldcg(src + (long long)(k + 1) * 64);
        const double v2 = __ldcg(src + (long long)(k + 2) * 64), v3 = __ldcg(src + (long long)(k + 3) * 64);
        acc += v0; acc += v1; acc += v2; acc += v3;
      }
      for (; k < k1; ++k) acc += __ldcg(src + (long long)k * 64);
      s_chain[chain][pair] = acc;
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      const int w = threadIdx.x & 1;  // thread 2g + w holds statistic w (0: sum, 1: sum of squares) of group g
      const double acc = ((s_chain[0][threadIdx.x] + s_chain[1][threadIdx.x]) + s_chain[2][threadIdx.x]) + s_chain[3][threadIdx.x];
      const double other = __shfl_xor_sync(0xffffffffu, acc, 1);
      const double S = w ? other : acc, Q = w ? acc : other;
      const double n = (double)HW * gs;
      const double m = S / n;
      double var = Q / n - m * m;
      if (var < 0) var = 0;
      const float r = w ? (float)(1.0 / sqrt(var + 1e-5)) : (float)m;
      s_stats[threadIdx.x] = r;
      if (blockIdx.x == 0) sc.stats[(long long)b * 64 + threadIdx.x] = r;
    }
    __syncthreads();
    // ---------------- phase 2 ----------------
    float a8[8], h8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = cv * 8 + k, g = c / gs;
      const float a = s_stats[2 * g + 1] * __ldg(gamma + c);
      a8[k] = a;
      h8[k] = __ldg(beta + c) - s_stats[2 * g] * a;
    }
    bf16* yb = y + ((long long)b * HW) * ldy + cv * 8;
    auto xform = [&](const uint4& u) {
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float h = fmaf(f[k], a8[k], h8[k]);
        f[k] = silu ? __fdividef(h, 1.f + __expf(-h)) : h;
      }
      return pack8(f);
    };
    p = p0 + pl;
    for (; p + 7LL * lanes < p1; p += 8LL * lanes) {
      uint4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] = *reinterpret_cast<const uint4*>(xb + (p + (long long)k * lanes) * ld);
#pragma unroll
      for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(yb + (p + (long long)k * lanes) * ldy) = xform(u[k]);
    }
    for (; p + 3LL * lanes < p1; p += 4LL * lanes) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = *reinterpret_cast<const uint4*>(xb + (p + (long long)k * lanes) * ld);
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(yb + (p + (long long)k * lanes) * ldy) = xform(u[k]);
    }
    for (; p < p1; p += lanes) *reinterpret_cast<uint4*>(yb + p * ldy) = xform(*reinterpret_cast<const uint4*>(xb + p * ld));
  }
}

static constexpr int kGnMaxBatch = 256;
static constexpr size_t kGnMaxSmem = 32 * 1024;  // lanes * C * 8 B <= 512 / (C / 8) * C * 8 B (static + dynamic stays under the 48 KB default)
static int g_gn_resident_ctas = 0;  // CTAs of gn_fused_kernel that are co-resident on this device (set by init_norm_kernels)

inline void init_norm_kernels() {
  int dev = 0, sms = 0, occ = 0;
  SDTF_CUDA(cudaGetDevice(&dev));
  SDTF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  SDTF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_fused_kernel, 512, kGnMaxSmem));
  SDTF_CHECK(occ >= 1, "gn_fused_kernel does not fit on an SM");
  g_gn_resident_ctas = sms * (occ > 2 ? 2 : occ);
}

// x: NHWC view (possibly a slice of a wider buffer); y dense [B][HW][C] (ldy may differ).  Returns the launch count.
inline int launch_groupnorm(cudaStream_t st, const View& x, const float* gamma, const float* beta, bool silu, bf16* y,
                            long long ldy, const GnScratch& sc, int batch_class = 0) {
  SDTF_CHECK(x.C % 32 == 0 && x.C % 8 == 0, "GroupNorm needs C % 32 == 0");
  SDTF_CHECK(x.B <= kGnMaxBatch, "GroupNorm: batch too large for the statistics scratch");
  SDTF_CHECK(g_gn_resident_ctas > 0, "init_norm_kernels() was not called");
  const long long HW = (long long)x.H * x.W;
  const int vecs = x.C / 8;
  int threads = (512 / vecs) * vecs;
  SDTF_CHECK(threads >= vecs && threads <= 512 && threads >= 256, "GroupNorm: unsupported channel count");
  const int lanes = threads / vecs;
  const size_t smem = (size_t)lanes * x.C * 2 * sizeof(float);
  SDTF_CHECK(smem <= kGnMaxSmem, "GroupNorm: reduction scratch exceeds the shared-memory budget the occupancy was computed for");
  // CTAs per sample: a function of (HW, C) and the BATCH CLASS only, so statistics do not depend on how samples are batched
  // inside a class (gemm_host.cuh kSmallBatch: <= 4 samples per denoise step; `batch_class` carries the step's sample
  // count when the call at hand only sees part of it — one branch of a CFG pair evaluated on its own).
  // Large batches: ~32 KB of the sample per CTA (four 16-byte vectors per thread: one round of loads per phase) up to 16 CTAs —
  // a UNet batch of 16 is then one co-resident round on 148 SMs; the VAE's 16+ MB samples get 1 MB per CTA, 32 to 64 CTAs
  // (measured: 128 CTAs per sample makes a single image's decode 1.3 ms faster but a batch of 8, walked in four rounds,
  // 0.8 ms slower; 64 keeps the batch at two rounds).  Small batches (one or two prompts: 2 x 16 CTAs left 116 SMs idle and
  // a 2.6 MB sample took 35 us) cut the sample finer: up to 64 CTAs per sample.
  static const int small_on = getenv("SDTF_GN_SMALL") ? atoi(getenv("SDTF_GN_SMALL")) : 1;  // A/B: 0 = round-1 partition
  const int bc = batch_class > 0 ? batch_class : x.B;
  const int cap = (small_on && bc <= 4) ? 64 : 16;
  const long long bytes = HW * x.C * 2;
  long long nblk = ceil_div_ll(bytes, 32 * 1024);
  if (nblk > cap) {
    nblk = cap;
    if (bytes > (16LL << 20)) {
      const long long per = cap > 16 ? 256 * 1024 : 1024 * 1024;  // 1 MB per CTA, 256 KB in the small-batch class
      nblk = ceil_div_ll(bytes, per);
      const long long lo = 32, hi = cap > 64 ? cap : 64;
      if (nblk < lo) nblk = lo;
      if (nblk > hi) nblk = hi;
    }
  }
  if (nblk > HW / lanes) nblk = HW / lanes;
  if (nblk < 1) nblk = 1;
  const long long ppc = ceil_div_ll(HW, nblk);
  nblk = ceil_div_ll(HW, ppc);
  long long by = (g_gn_resident_ctas / (sc.share > 0 ? sc.share : 1)) / nblk;  // samples per round: every CTA of the grid must be resident
  if (by > x.B) by = x.B;
  SDTF_CHECK(by >= 1, "GroupNorm: a sample's CTAs do not fit on the device at once");
  dim3 grid((unsigned)nblk, (unsigned)by);
  static const int chains_env = getenv("SDTF_GN_CHAINS") ? atoi(getenv("SDTF_GN_CHAINS")) : 4;
  GnScratch scc = sc;
  scc.chains = chains_env == 1 ? 1 : 4;
  // The per-sample barrier needs every CTA of the grid resident at once.  A COOPERATIVE launch makes that the driver's
  // promise instead of an assumption about what else runs on the device: if another stream / process / engine holds
  // SMs, the launch is refused (cudaErrorCooperativeLaunchTooLarge) or waits for room — it can no longer start with
  // part of the grid and spin.  (SDTF_GN_COOP=0: plain launch, for A/B timing.)
  static const int coop = getenv("SDTF_GN_COOP") ? atoi(getenv("SDTF_GN_COOP")) : 1;
  if (coop && !pdl_enabled()) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    SDTF_CUDA(cudaLaunchKernelEx(&cfg, gn_fused_kernel, (const bf16*)x.p, (long long)x.ld, x.C, HW, (int)ppc, x.B, scc, gamma, beta,
                                 silu ? 1 : 0, y, ldy));
  } else {
    launch_pdl(gn_fused_kernel, grid, dim3(threads), smem, st, 1, x.p, x.ld, x.C, HW, (int)ppc, x.B, scc, gamma, beta, silu ? 1 : 0, y, ldy);
  }
  SDTF_CUDA(cudaGetLastError());
  return 1;
}

// ------------------------------------------------------------------------------------------------------
// LayerNorm(eps 1e-5) over C, one warp per token (reference: diffusion_model.py:84,86,88)
// ------------------------------------------------------------------------------------------------------
template <int MAXV>  // max 8-channel vectors per lane
__global__ void __launch_bounds__(256)
layernorm_kernel(const bf16* __restrict__ x, long long ld, int C, long long rows, const float* __restrict__ gamma,
                 const float* __restrict__ beta, bf16* __restrict__ y, long long ldy) {
  const int lane = threadIdx.x & 31;
  const int vecs = C >> 3;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  // this lane's slice of gamma / beta stays in registers for every row the warp normalises
  float ga[MAXV][8], be[MAXV][8];
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int cv = lane + 32 * j;
    if (cv < vecs) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + cv * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + cv * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + cv * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + cv * 8 + 4));
      ga[j][0] = g0.x; ga[j][1] = g0.y; ga[j][2] = g0.z; ga[j][3] = g0.w; ga[j][4] = g1.x; ga[j][5] = g1.y; ga[j][6] = g1.z; ga[j][7] = g1.w;
      be[j][0] = b0.x; be[j][1] = b0.y; be[j][2] = b0.z; be[j][3] = b0.w; be[j][4] = b1.x; be[j][5] = b1.y; be[j][6] = b1.z; be[j][7] = b1.w;
    }
  }
  const float inv_c = 1.f / (float)C;
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += nwarps) {
    float f[MAXV][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int cv = lane + 32 * j;
      if (cv < vecs) {
        const uint4 u = *reinterpret_cast<const uint4*>(x + row * ld + cv * 8);
        unpack8(u, f[j]);
#pragma unroll
        for (int k = 0; k < 8; ++k) s += f[j][k];
      }
    }
    const float mean = warp_sum(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int cv = lane + 32 * j;
      if (cv < vecs) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float d = f[j][k] - mean;
          q = fmaf(d, d, q);
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_c + 1e-5f);
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int cv = lane + 32 * j;
      if (cv < vecs) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf((f[j][k] - mean) * rstd, ga[j][k], be[j][k]);
        *reinterpret_cast<uint4*>(y + row * ldy + cv * 8) = pack8(o);
      }
    }
  }
}

// C == 8 * G * VPL: a row is normalised by a group of G lanes (G = 8 / 16 / 32 for C = 320 / 640 / 1280 at VPL = 5), so
// a warp works on 32 / G rows at once with every lane busy (the one-warp-per-row kernel above leaves 3 lanes in 8 idle
// at C = 320 and walks its rows one memory round trip at a time).  gamma / beta sit in shared memory.
template <int G, int VPL>
__global__ void __launch_bounds__(256)
layernorm_group_kernel(const bf16* __restrict__ x, long long ld, int C, long long rows, const float* __restrict__ gamma,
                       const float* __restrict__ beta, bf16* __restrict__ y, long long ldy) {
  extern __shared__ __align__(16) float ln_sm[];  // gamma [C], beta [C]
  pdl_trigger();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    ln_sm[i] = __ldg(gamma + i);
    ln_sm[C + i] = __ldg(beta + i);
  }
  __syncthreads();
  pdl_wait();
  constexpr int R = 32 / G;
  const int lane = threadIdx.x & 31, sub = lane / G, gl = lane % G;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const float inv_c = 1.f / (float)C;
  for (long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R; row0 < rows; row0 += nwarps * R) {
    const long long row = row0 + sub;
    const bool valid = row < rows;
    float f[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      uint4 u = make_uint4(0, 0, 0, 0);
      if (valid) u = *reinterpret_cast<const uint4*>(x + row * ld + (gl + G * j) * 8);
      unpack8(u, f[j]);
    }
#pragma unroll
    for (int j = 0; j < VPL; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) s += f[j][k];
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_c;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = f[j][k] - mean;
        q = fmaf(d, d, q);
      }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * inv_c + 1e-5f);
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int c0 = (gl + G * j) * 8;
      const float4 g0 = *reinterpret_cast<const float4*>(ln_sm + c0), g1 = *reinterpret_cast<const float4*>(ln_sm + c0 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(ln_sm + C + c0), b1 = *reinterpret_cast<const float4*>(ln_sm + C + c0 + 4);
      const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaf((f[j][k] - mean) * rstd, ga[k], be[k]);
      if (valid) *reinterpret_cast<uint4*>(y + row * ldy + c0) = pack8(o);
    }
  }
}

template <int G, int VPL>
inline void launch_layernorm_group(cudaStream_t st, const bf16* x, long long ld, int C, long long rows, const float* gamma,
                                   const float* beta, bf16* y, long long ldy) {
  constexpr int R = 32 / G;
  long long blocks = ceil_div_ll(rows, 8 * R);
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_pdl(layernorm_group_kernel<G, VPL>, dim3((unsigned)blocks), dim3(256), (size_t)C * 2 * sizeof(float), st, 1, x, ld, C, rows, gamma,
             beta, y, ldy);
}

inline void launch_layernorm(cudaStream_t st, const bf16* x, long long ld, int C, long long rows, const float* gamma,
                             const float* beta, bf16* y, long long ldy) {
  SDTF_CHECK(C % 8 == 0 && C <= 32 * 8 * 5, "LayerNorm: C must be a multiple of 8 and <= 1280");
  const int vecs = C / 8;
  if (C == 320) launch_layernorm_group<8, 5>(st, x, ld, C, rows, gamma, beta, y, ldy);
  else if (C == 640) launch_layernorm_group<16, 5>(st, x, ld, C, rows, gamma, beta, y, ldy);
  else if (C == 1280) launch_layernorm_group<32, 5>(st, x, ld, C, rows, gamma, beta, y, ldy);
  else {
    long long blocks = ceil_div_ll(rows, 8);
    if (blocks > 148 * 8) blocks = 148 * 8;  // persistent warps: each keeps its gamma / beta slice and strides over rows
    if (vecs <= 64) layernorm_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(x, ld, C, rows, gamma, beta, y, ldy);
    else if (vecs <= 96) layernorm_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(x, ld, C, rows, gamma, beta, y, ldy);
    else layernorm_kernel<5><<<(unsigned)blocks, 256, 0, st>>>(x, ld, C, rows, gamma, beta, y, ldy);
  }
  SDTF_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------
// skinny linear for the time-embedding chain (diffusion_model.py:184-188 and every ResBlock's
// time_emb_proj :30,47): out[m][n] = act(sum_k x[m][k] * W[n][k] + b[n]), M <= 64 rows, one warp per n.
// ------------------------------------------------------------------------------------------------------
template <int MT>
__global__ void __launch_bounds__(256)
skinny_linear_kernel(const float* __restrict__ x, int M, int K, const bf16* __restrict__ W, const float* __restrict__ bias,
                     int N, int silu, float* __restrict__ out, int ldo) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  for (int m0 = 0; m0 < M; m0 += MT) {
    float acc[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = 0.f;
    for (int k = lane * 8; k < K; k += 256) {
      const uint4 u = *reinterpret_cast<const uint4*>(W + (long long)n * K + k);
      float w[8];
      unpack8(u, w);
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        if (m0 + m < M) {
          const float4 a0 = *reinterpret_cast<const float4*>(x + (long long)(m0 + m) * K + k);
          const float4 a1 = *reinterpret_cast<const float4*>(x + (long long)(m0 + m) * K + k + 4);
          acc[m] += a0.x * w[0] + a0.y * w[1] + a0.z * w[2] + a0.w * w[3] + a1.x * w[4] + a1.y * w[5] + a1.z * w[6] + a1.w * w[7];
        }
      }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      float v = warp_sum(acc[m]);
      if (lane == 0 && m0 + m < M) {
        v += bias ? bias[n] : 0.f;
        if (silu) v = v / (1.f + __expf(-v));
        out[(long long)(m0 + m) * ldo + n] = v;
      }
    }
  }
}
inline void launch_skinny_linear(cudaStream_t st, const float* x, int M, int K, const bf16* W, const float* bias, int N,
                                 bool silu, float* out, int ldo) {
  SDTF_CHECK(K % 8 == 0, "skinny linear: K % 8");
  skinny_linear_kernel<8><<<(unsigned)ceil_div(N, 8), 256, 0, st>>>(x, M, K, W, bias, N, silu ? 1 : 0, out, ldo);
  SDTF_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------
// CLIP text tower pieces (reference: text_encoder.py:22-33 CLIPEmbedding, :58-99 CLIPAttention).  Once per prompt, 77
// tokens: latency-bound, plain CUDA.
// ------------------------------------------------------------------------------------------------------
// x[b][t][:] = token_embedding[tokens[b][t]] + position_embedding[t]  -> fp32 residual stream
__global__ void clip_embed_kernel(const int* __restrict__ tokens, const int* __restrict__ positions /* null: 0..T-1 */, int pos_rows,
                                  const float* __restrict__ tok_emb, const float* __restrict__ pos_emb, int vocab, int max_len, int T,
                                  int C, long long rows, float* __restrict__ out) {
  const long long total = rows * (C >> 2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (C >> 2);
    const int c = (int)(i - r * (C >> 2)) << 2;
    int tok = tokens[r];
    tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
    const float4 a = __ldg(reinterpret_cast<const float4*>(tok_emb + (long long)tok * C + c));
    int pos = (int)(r % T);
    if (positions) pos = positions[pos_rows == 1 ? pos : r];  // (1,T) broadcast over the batch, or (B,T)
    pos = pos < 0 ? 0 : (pos >= max_len ? max_len - 1 : pos);
    const float4 b = __ldg(reinterpret_cast<const float4*>(pos_emb + (long long)pos * C + c));
    *reinterpret_cast<float4*>(out + r * C + c) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// residual stream of the text tower in fp32: x += delta (bf16 GEMM output, optional), then LayerNorm(x) -> bf16 (the
// next GEMM's operand) or fp32 (the final context).  One warp per 768-wide row: 24 values per lane in registers.
// (With a bf16 stream the 24 roundings over 12 layers left 1.7e-2 of error on the context; fp32 leaves < 5e-3.)
template <int VPL>  // float4 vectors per lane: C == 128 * VPL
__global__ void __launch_bounds__(256)
clip_add_ln_kernel(float* __restrict__ x, const bf16* __restrict__ delta, int C, long long rows, const float* __restrict__ gamma,
                   const float* __restrict__ beta, bf16* __restrict__ out_bf16, float* __restrict__ out_f32) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[VPL][4];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int c = (lane + 32 * j) * 4;
    const float4 a = *reinterpret_cast<const float4*>(x + row * C + c);
    v[j][0] = a.x; v[j][1] = a.y; v[j][2] = a.z; v[j][3] = a.w;
    if (delta) {
      const uint2 d = *reinterpret_cast<const uint2*>(delta + row * C + c);
      const __nv_bfloat162 d0 = *reinterpret_cast<const __nv_bfloat162*>(&d.x), d1 = *reinterpret_cast<const __nv_bfloat162*>(&d.y);
      v[j][0] += __bfloat162float(d0.x); v[j][1] += __bfloat162float(d0.y);
      v[j][2] += __bfloat162float(d1.x); v[j][3] += __bfloat162float(d1.y);
      *reinterpret_cast<float4*>(x + row * C + c) = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
    }
    s += (v[j][0] + v[j][1]) + (v[j][2] + v[j][3]);
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = v[j][k] - mean;
      q = fmaf(d, d, q);
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int c = (lane + 32 * j) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c)), b = __ldg(reinterpret_cast<const float4*>(beta + c));
    const float o0 = fmaf((v[j][0] - mean) * rstd, g.x, b.x), o1 = fmaf((v[j][1] - mean) * rstd, g.y, b.y);
    const float o2 = fmaf((v[j][2] - mean) * rstd, g.z, b.z), o3 = fmaf((v[j][3] - mean) * rstd, g.w, b.w);
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * C + c) = make_float4(o0, o1, o2, o3);
    if (out_bf16) {
      uint2 o;
      o.x = tc05::pack_bf16(o0, o1);
      o.y = tc05::pack_bf16(o2, o3);
      *reinterpret_cast<uint2*>(out_bf16 + row * C + c) = o;
    }
  }
}

// causal self-attention over T <= 128 tokens, head size 64: one CTA per (head, sample), a warp per query row.
// qkv: [B*T][3*C] bf16 (q | k | v, q already scaled by d^-1/2 through its weights), out: [B*T][C] bf16
__global__ void __launch_bounds__(128)
clip_causal_attn_kernel(const bf16* __restrict__ qkv, int T, int C, bf16* __restrict__ out) {
  constexpr int D = 64, LD = D + 1;
  extern __shared__ float ca_sm[];  // K [T][65], V [T][65], P [4 warps][128]
  float* sK = ca_sm;
  float* sV = sK + T * LD;
  float* sP = sV + T * LD;
  const int head = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* base = qkv + (long long)b * T * 3 * C + head * D;
  for (int i = threadIdx.x; i < T * D; i += blockDim.x) {
    const int t = i / D, d = i - t * D;
    sK[t * LD + d] = __bfloat162float(base[(long long)t * 3 * C + C + d]);
    sV[t * LD + d] = __bfloat162float(base[(long long)t * 3 * C + 2 * C + d]);
  }
  __syncthreads();
  float* myP = sP + warp * 128;
  for (int q = warp; q < T; q += 4) {
    const float q0 = __bfloat162float(base[(long long)q * 3 * C + lane]);
    const float q1 = __bfloat162float(base[(long long)q * 3 * C + 32 + lane]);
    float sc[4];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = lane + 32 * k;
      const int jj = j < T ? j : T - 1;  // every lane runs the same shuffles; masked keys are discarded below
      float acc = 0.f;
#pragma unroll 16
      for (int d = 0; d < 32; ++d) {
        acc = fmaf(__shfl_sync(0xffffffffu, q0, d), sK[jj * LD + d], acc);
        acc = fmaf(__shfl_sync(0xffffffffu, q1, d), sK[jj * LD + 32 + d], acc);
      }
      if (j > q) acc = -INFINITY;  // causal mask (text_encoder.py:78-81: -inf above the diagonal)
      sc[k] = acc;
      mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float e = sc[k] == -INFINITY ? 0.f : __expf(sc[k] - mx);
      myP[lane + 32 * k] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j <= q; ++j) {
      const float pj = myP[j];
      o0 = fmaf(pj, sV[j * LD + lane], o0);
      o1 = fmaf(pj, sV[j * LD + 32 + lane], o1);
    }
    const float inv = 1.f / sum;
    bf16* orow = out + ((long long)b * T + q) * C + head * D;
    orow[lane] = __float2bfloat16(o0 * inv);
    orow[32 + lane] = __float2bfloat16(o1 * inv);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------------
// Fused CFG combine + CFG-rescale + scheduler update (+ inpaint blend)  — one CTA per sample.
//   reference: stable_diffusion.py:458 (combine), :304-315 (rescale_noise_cfg, population std over h,w,c),
//   scheduler.py:285-312 (DDIM / TCD update, coefficients pre-combined on the host in fp64),
//   stable_diffusion.py:469-475 (inpaint: latent = noised_init(t)*(1-m) + latent*m).
// latent' = ca * latent + cb * eps (+ cn * noise);  writes the fp32 master latent and the bf16 8-channel
// copies (duplicated for the uncond / cond halves of the next UNet batch).
// ------------------------------------------------------------------------------------------------------
struct StepCoef {
  float guidance, rescale;  // guidance <= 0: eps = eps_c (no CFG)
  float ca, cb, cn;         // latent' = ca*latent + cb*eps + cn*noise
  float sig_t, noi_t;       // inpaint re-noising of the init latent at the current t
};

__global__ void __launch_bounds__(512)
cfg_sched_kernel(const float* __restrict__ eps_u, const float* __restrict__ eps_c, const float* latent /* may alias out */,
                 const StepCoef* __restrict__ coefs, const int* __restrict__ step_ptr, const float* __restrict__ noise,
                 const float* __restrict__ mask,
                 const float* __restrict__ init_latent, const float* __restrict__ init_noise, int n_per_sample,
                 float* out, bf16* __restrict__ out_bf16 /* [2B or B][HW][8] */, int B, int dup) {
  const int step = step_ptr ? *step_ptr : 0;
  const StepCoef c = coefs[step];
  const int b = blockIdx.x;
  if (noise) noise += (long long)step * B * n_per_sample;  // per-step TCD noise table
  const long long off = (long long)b * n_per_sample;
  __shared__ float red[2][16];
  __shared__ float stat[4];
  const bool cfg = c.guidance > 0.f && eps_u != nullptr;
  float factor = 1.f;
  if (cfg && c.rescale > 0.f) {
    // population std of eps_c and of the combined eps over the whole sample (two-pass, fp32); 16-byte loads (n_per_sample
    // is a multiple of 4: h*w*4 channels) — the sample's 128 KB stay in L1 for the second and third pass
    const int n4 = n_per_sample >> 2;
    const float4* u4 = reinterpret_cast<const float4*>(eps_u + off);
    const float4* t4 = reinterpret_cast<const float4*>(eps_c + off);
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 u = u4[i], t = t4[i];
      s1 += (t.x + t.y) + (t.z + t.w);
      s2 += ((u.x + c.guidance * (t.x - u.x)) + (u.y + c.guidance * (t.y - u.y))) +
            ((u.z + c.guidance * (t.z - u.z)) + (u.w + c.guidance * (t.w - u.w)));
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, d = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; d += red[1][i]; }
      stat[0] = a / n_per_sample; stat[1] = d / n_per_sample;
    }
    __syncthreads();
    const float m1 = stat[0], m2 = stat[1];
    float q1 = 0.f, q2 = 0.f;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 u = u4[i], t = t4[i];
      const float uu[4] = {u.x, u.y, u.z, u.w}, tt[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float e = uu[k] + c.guidance * (tt[k] - uu[k]);
        q1 += (tt[k] - m1) * (tt[k] - m1);
        q2 += (e - m2) * (e - m2);
      }
    }
    q1 = warp_sum(q1); q2 = warp_sum(q2);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = q1; red[1][threadIdx.x >> 5] = q2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, d = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; d += red[1][i]; }
      const float std_text = sqrtf(a / n_per_sample), std_cfg = sqrtf(d / n_per_sample) + 1e-5f;
      stat[2] = c.rescale * (std_text / std_cfg) + (1.f - c.rescale);
    }
    __syncthreads();
    factor = stat[2];
  }
  for (int i = threadIdx.x; i < n_per_sample; i += blockDim.x) {
    const float t = eps_c[off + i];
    float e = t;
    if (cfg) {
      const float u = eps_u[off + i];
      e = (u + c.guidance * (t - u)) * factor;
    }
    float x = c.ca * latent[off + i] + c.cb * e;
    if (noise) x += c.cn * noise[off + i];
    if (mask) {
      const float mk = mask[i >> 2];  // mask is (h, w, 1), shared by all samples and channels
      const float orig = c.sig_t * init_latent[i] + c.noi_t * init_noise[off + i];
      x = orig * (1.f - mk) + x * mk;
    }
    out[off + i] = x;
    if (out_bf16) {
      const int pix = i >> 2, ch = i & 3;
      const long long hw = n_per_sample >> 2;
      out_bf16[((long long)b * hw + pix) * 8 + ch] = __float2bfloat16(x);
      if (dup) out_bf16[((long long)(b + B) * hw + pix) * 8 + ch] = __float2bfloat16(x);
    }
  }
}

__global__ void step_advance_kernel(int* step) { *step += 1; }
// t_emb table row select: out[b][:] = table[*step][:] for all b
__global__ void temb_select_kernel(const float* __restrict__ table, const int* __restrict__ step_ptr, int dim, int B,
                                   float* __restrict__ out) {
  const int step = *step_ptr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * dim; i += gridDim.x * blockDim.x)
    out[i] = table[(long long)step * dim + (i % dim)];
}
// y(view) += c(dense bf16)   (ControlNet residual injection, diffusion_model.py:230-234)
__global__ void add_inplace_kernel(bf16* __restrict__ y, long long ld, int C, long long pixels, const bf16* __restrict__ c) {
  const int vecs = C >> 3;
  const long long total = pixels * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / vecs;
    const int cv = (int)(i - p * vecs);
    uint4* yp = reinterpret_cast<uint4*>(y + p * ld + cv * 8);
    float a[8], b[8];
    unpack8(*yp, a);
    unpack8(*reinterpret_cast<const uint4*>(c + i * 8), b);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += b[k];
    *yp = pack8(a);
  }
}

// fp32 [B][HW][Cs] -> bf16 [B*(dup?2:1)][HW][Cd] (Cd >= Cs, zero padded), scaled
__global__ void cast_pad_kernel(const float* __restrict__ x, long long pixels, int Cs, int Cd, float scale, bf16* __restrict__ y,
                                int dup) {
  const long long total = pixels * Cd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / Cd;
    const int c = (int)(i - p * Cd);
    const bf16 v = __float2bfloat16(c < Cs ? x[p * Cs + c] * scale : 0.f);
    y[i] = v;
    if (dup) y[total + i] = v;
  }
}
inline void launch_cast_pad(cudaStream_t st, const float* x, long long pixels, int Cs, int Cd, float scale, bf16* y, bool dup) {
  long long blocks = ceil_div_ll(pixels * Cd, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cast_pad_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, pixels, Cs, Cd, scale, y, dup ? 1 : 0);
  SDTF_CUDA(cudaGetLastError());
}

// bf16 NHWC (ld) first Cs channels -> fp32 dense
__global__ void cast_out_kernel(const bf16* __restrict__ x, long long ld, long long pixels, int Cs, float* __restrict__ y) {
  const long long total = pixels * Cs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / Cs;
    y[i] = __bfloat162float(x[p * ld + (i - p * Cs)]);
  }
}
inline void launch_cast_out(cudaStream_t st, const bf16* x, long long ld, long long pixels, int Cs, float* y) {
  long long blocks = ceil_div_ll(pixels * Cs, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cast_out_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, ld, pixels, Cs, y);
  SDTF_CUDA(cudaGetLastError());
}

// decoded fp32 [pixels][3] in ~[-1,1] -> uint8 with the reference's arithmetic:
// clip(((d + 1) * 0.5 [blend with source]) * 255, 0, 255) truncated (stable_diffusion.py:483-486)
__global__ void to_uint8_kernel(const float* __restrict__ d, long long n, const float* __restrict__ src_img,
                                const float* __restrict__ src_mask, long long n_per_sample, uint8_t* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = (d[i] + 1.f) * 0.5f;
    if (src_img) {
      const long long j = i % n_per_sample;
      const float m = src_mask[j / 3];
      v = src_img[j] * (1.f - m) + v * m;
    }
    v = fminf(fmaxf(v * 255.f, 0.f), 255.f);
    out[i] = (uint8_t)v;
  }
}

// ------------------------------------------------------------------------------------------------------
// weight packing: PyTorch layout fp32 [O][I][taps] -> bf16 [taps][Ntot][Kp] rows n0 + [0, nrows)
//   row_map[r] = source output channel for packed row n0 + r (or -1 for a zero row); scale folds constants.
// ------------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ src, int I, int taps, const int* __restrict__ row_map, int nrows,
                                   int n0, int Ntot, int Kp, int k0, float scale, bf16* __restrict__ dst,
                                   const float* __restrict__ kscale = nullptr /* [I]: LayerNorm gamma folded into the columns */) {
  const long long total = (long long)taps * nrows * I;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % I);
    long long r = i / I;
    const int row = (int)(r % nrows);
    const int t = (int)(r / nrows);
    const int so = row_map ? row_map[row] : row;
    float v = so >= 0 ? src[((long long)so * I + k) * taps + t] * scale : 0.f;
    if (kscale) v *= kscale[k];
    dst[((long long)t * Ntot + n0 + row) * Kp + k0 + k] = __float2bfloat16(v);
  }
}

// Nearest-2x upsample followed by a 3x3 convolution (diffusion_model.py:132-139, image_decoder.py:36-47) == four 2x2
// convolutions on the SOURCE grid, one per output parity (py, px): output pixel (2y+py, 2x+px) reads source rows
// {y-1, y} (py = 0) or {y, y+1} (py = 1), and the 3x3 taps that land on the same source pixel are summed:
//   py = 0: tap row 0 <- k0,      tap row 1 <- k1 + k2        py = 1: tap row 0 <- k0 + k1,  tap row 1 <- k2
// (columns alike).  4/9 of the multiply-adds, no upsampled tensor.  src fp32 [O][I][3][3] -> dst bf16
// [class = 2 py + px][tap = 2 r + s][O][Kp], summed in fp32 and rounded once.
__global__ void pack_upconv_kernel(const float* __restrict__ src, int O, int I, int Kp, bf16* __restrict__ dst) {
  const long long total = 16LL * O * I;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % I);
    long long r0 = i / I;
    const int o = (int)(r0 % O);
    const int ct = (int)(r0 / O);  // class * 4 + tap
    const int py = ct >> 3, px = (ct >> 2) & 1, r = (ct >> 1) & 1, sc = ct & 1;
    // 3x3 rows merged into 2x2 row r of parity py: [lo, hi]
    const int ylo = py == 0 ? (r == 0 ? 0 : 1) : (r == 0 ? 0 : 2), yhi = py == 0 ? (r == 0 ? 0 : 2) : (r == 0 ? 1 : 2);
    const int xlo = px == 0 ? (sc == 0 ? 0 : 1) : (sc == 0 ? 0 : 2), xhi = px == 0 ? (sc == 0 ? 0 : 2) : (sc == 0 ? 1 : 2);
    const float* w = src + ((long long)o * I + k) * 9;
    float acc = 0.f;
    for (int y = ylo; y <= yhi; ++y)
      for (int x = xlo; x <= xhi; ++x) acc += w[y * 3 + x];
    dst[((long long)ct * O + o) * Kp + k] = __float2bfloat16(acc);
  }
}

// LayerNorm folded into the linear that follows it: y = LN(x) W^T + b = rstd (x W'^T - mean c1) + c0 with
//   W'[n][k] = bf16(gamma[k] W[n][k])   (packed by pack_weight_kernel with kscale = gamma)
//   c1[n]    = sum_k W'[n][k]           (of the ROUNDED weights: it must cancel exactly what the tensor core summed)
//   c0[n]    = b[n] + sum_k beta[k] W[n][k]
// one warp per packed row; row_map as in pack_weight_kernel (-1: zero row)
__global__ void ln_fold_consts_kernel(const float* __restrict__ src, int I, const int* __restrict__ row_map, int nrows, int n0,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ c1,
                                      float* __restrict__ c0 /* in: bias (or 0), out: c0 */) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const int so = row_map ? row_map[row] : row;
  float s1 = 0.f, s0 = 0.f;
  if (so >= 0)
    for (int k = lane; k < I; k += 32) {
      const float w = src[(long long)so * I + k];
      s1 += __bfloat162float(__float2bfloat16(w * gamma[k]));
      s0 = fmaf(w, beta[k], s0);
    }
  s1 = warp_sum(s1);
  s0 = warp_sum(s0);
  if (lane == 0) {
    c1[n0 + row] = s1;
    c0[n0 + row] += s0;
  }
}

__global__ void vec_add_kernel(float* __restrict__ dst, const float* __restrict__ src, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

__global__ void gather_f32_kernel(const float* __restrict__ src, const int* __restrict__ row_map, int n, float scale,
                                  float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int so = row_map ? row_map[i] : i;
    dst[i] = so >= 0 ? src[so] * scale : 0.f;
  }
}

}  // namespace sdtf
