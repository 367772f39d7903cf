// common.cuh — error handling, tensor views and the TMA tensor-map encoder shared by all launchers.
#pragma once
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace sdtf {

using bf16 = __nv_bfloat16;

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define SDTF_CUDA(expr)                                                                          \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      throw ::sdtf::Error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +   \
                          __FILE__ + ":" + std::to_string(__LINE__));                            \
  } while (0)

#define SDTF_CHECK(cond, msg)                                                                    \
  do {                                                                                           \
    if (!(cond))                                                                                 \
      throw ::sdtf::Error(std::string("check failed: ") + #cond + " — " + (msg) + " at " +       \
                          __FILE__ + ":" + std::to_string(__LINE__));                            \
  } while (0)

// NHWC bf16 activation view.  `ld` is the element distance between consecutive pixels, so a view can be a
// channel slice [c0, c0+C) of a wider buffer (that is how channel concats are laid out: the producers of
// the two halves write straight into one buffer, see unet.cu).
struct View {
  bf16* p = nullptr;
  int B = 0, H = 0, W = 0, C = 0;
  long long ld = 0;
  long long pixels() const { return (long long)B * H * W; }
  View slice(int c0, int c) const {
    View v = *this;
    v.p = p + c0;
    v.C = c;
    return v;
  }
  // same memory seen as a [1][1][B*H*W] token matrix
  View tokens() const {
    View v = *this;
    v.W = B * H * W;
    v.H = 1;
    v.B = 1;
    return v;
  }
};

inline PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SDTF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    SDTF_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// bf16 tensor map, 128-byte swizzle, zero fill out of bounds.  dims/box/estr are innermost-first;
// strides_bytes[i] is the byte stride of dimension i+1.
inline CUtensorMap make_tmap_bf16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                  const uint32_t* box, const uint32_t* estr,
                                  CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  CUtensorMap m;
  SDTF_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
  for (int i = 0; i + 1 < rank; ++i)
    SDTF_CHECK(strides_bytes[i] % 16 == 0, "TMA strides must be multiples of 16 bytes");
  CUresult r = tensor_map_encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                                    reinterpret_cast<const cuuint64_t*>(dims),
                                    reinterpret_cast<const cuuint64_t*>(strides_bytes),
                                    reinterpret_cast<const cuuint32_t*>(box),
                                    reinterpret_cast<const cuuint32_t*>(estr), CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string s = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") rank " + std::to_string(rank);
    for (int i = 0; i < rank; ++i)
      s += " d" + std::to_string(i) + "=" + std::to_string(dims[i]) + "/box" + std::to_string(box[i]) + "/es" +
           std::to_string(estr[i]);
    for (int i = 0; i + 1 < rank; ++i) s += " s" + std::to_string(i) + "=" + std::to_string(strides_bytes[i]);
    throw Error(s);
  }
  return m;
}

// activation map over an NHWC view: dims (C, W, H, B); box (64, bw*stride, bh*stride, bn) with element
// strides (1, stride, stride, 1) => exactly bw x bh x bn pixels x 64 channels = 16 KB per box.
inline CUtensorMap make_act_tmap(const View& v, int bw, int bh, int bn, int stride) {
  uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.W, (uint64_t)v.H, (uint64_t)v.B};
  uint64_t strides[3] = {(uint64_t)v.ld * 2, (uint64_t)v.W * v.ld * 2, (uint64_t)v.H * v.W * v.ld * 2};
  uint32_t box[4] = {64, (uint32_t)(bw * stride), (uint32_t)(bh * stride), (uint32_t)bn};
  uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
  return make_tmap_bf16(v.p, 4, dims, strides, box, es);
}

// epilogue map over an NHWC output / residual view: dims (C, W, H, B); box (32 channels, bw, bh, bn) = 128 pixels x
// 64 bytes, 64-byte swizzle (the layout the GEMM epilogue stages its bf16 sub-tiles in).  Stores are clipped and
// loads zero-filled at the tensor edges by the hardware.
// `step` > 1: the W x H grid addresses every step-th pixel of a (step W) x (step H) tensor (the parity classes of a
// folded nearest-upsample convolution write interleaved output pixels; base points at the class's first pixel).
inline CUtensorMap make_epi_tmap(const bf16* base, int C, int W, int H, int B, long long ld, int bw, int bh, int bn, int step = 1) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t px = (uint64_t)ld * 2 * step, row = (uint64_t)W * step * ld * 2 * step;
  uint64_t strides[3] = {px, row, (uint64_t)H * step * W * step * ld * 2};
  uint32_t box[4] = {32, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
  uint32_t es[4] = {1, 1, 1, 1};
  return make_tmap_bf16(base, 4, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_64B);
}

// epilogue map over an output whose columns are HEADS of `head_stride` columns of which only the first `d` are real
// (q | k | v of the d = 40 attention levels: heads sit 64 columns apart so that a head is one aligned 128-byte row for
// the attention kernel's TMA loads, but the 24 padding columns are never written — the store is clipped at d by the
// hardware — and never read: the load maps clip too).  dims (d, heads, W, H, B); box (32, 1, bw, bh, bn).
inline CUtensorMap make_epi_tmap_heads(const bf16* base, int d, int head_stride, int heads, int W, int H, int B, long long ld, int bw,
                                       int bh, int bn) {
  uint64_t dims[5] = {(uint64_t)d, (uint64_t)heads, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t strides[4] = {(uint64_t)head_stride * 2, (uint64_t)ld * 2, (uint64_t)W * ld * 2, (uint64_t)H * W * ld * 2};
  uint32_t box[5] = {32, 1, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
  uint32_t es[5] = {1, 1, 1, 1, 1};
  return make_tmap_bf16(base, 5, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_64B);
}

// packed weight map [taps][N][K]: dims (K, N, taps), box (64, BN, 1)
inline CUtensorMap make_weight_tmap(const bf16* w, int K, int N, int taps, int BN) {
  uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)taps};
  uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)N * K * 2};
  uint32_t box[3] = {64, (uint32_t)BN, 1};
  uint32_t es[3] = {1, 1, 1};
  return make_tmap_bf16(w, 3, dims, strides, box, es);
}

// ---- programmatic dependent launch (PDL) ----
// Every hot kernel calls pdl_trigger() first (the NEXT kernel of the stream may be scheduled as soon as all CTAs of
// this one have started) and pdl_wait() after its own prologue (barrier init, TMEM allocation, descriptor prefetch,
// constant staging) and before it touches anything a predecessor wrote: launch latency and prologue of kernel i+1
// overlap the tail of kernel i instead of following it.  A kernel launched without the attribute sees both as no-ops.
// Measured on B200 inside the captured 25-step graph (profiles/r01_e_pdl_ab.md): UNet step 18.54 / 18.63 ms without
// vs 18.61 / 18.66 ms with the attribute, VAE decode 22.1 -> 23.6 ms — graph launches already have sub-microsecond
// gaps and the step runs under the 1000 W power cap, so filling the gaps only lowers the clock.  The attribute is
// therefore OFF by default; SDTF_PDL=1 turns it on (back-to-back launches of one kernel outside a graph gain ~9 %).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
inline int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    SDTF_CUDA(cudaGetDevice(&dev));
    SDTF_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}
inline bool pdl_enabled() {
  static const int v = getenv("SDTF_PDL") ? atoi(getenv("SDTF_PDL")) : 0;
  return v != 0;
}
// launch with optional thread-block cluster (cluster_x > 1) and the PDL attribute
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  SDTF_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace sdtf
