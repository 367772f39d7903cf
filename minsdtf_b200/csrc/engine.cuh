// engine.cuh — engine state: device arena, weight store + packing, and the launch context every model graph
// issues its kernels through (dry-run aware, so the workspace high-water mark is measured before anything runs).
#pragma once
#include <nvtx3/nvToolsExt.h>  // header-only: ranges are no-ops unless a tool (ncu --nvtx, nsys) injects itself

#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "attn.cuh"
#include "gemm_host.cuh"
#include "ops.cuh"

namespace sdtf {

// ----------------------------------------------------------------------------------------------------------
// Bump arena over one cudaMalloc block.  Stream order makes reuse after release() safe: every kernel of a
// forward pass is enqueued on the engine's single stream.
// ----------------------------------------------------------------------------------------------------------
struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  bool dry = false;  // dry: only measure
  void* alloc(size_t bytes) {
    size_t a = (off + 1023) & ~(size_t)1023;
    size_t end = a + ((bytes + 1023) & ~(size_t)1023);
    if (!dry && end > cap) throw Error("workspace arena exhausted (internal sizing error)");
    off = end;
    if (off > peak) peak = off;
    // dry runs hand out a dummy non-null, suitably aligned address that is never dereferenced
    return dry ? reinterpret_cast<void*>((uintptr_t)0x10000000 + a) : base + a;
  }
  template <class T>
  T* alloc_n(size_t n) { return reinterpret_cast<T*>(alloc(n * sizeof(T))); }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }
};

// persistent device allocations (weights), freed with the engine
struct DevicePool {
  std::vector<void*> ptrs;
  size_t bytes = 0;
  void* alloc(size_t n) {
    void* p = nullptr;
    SDTF_CUDA(cudaMalloc(&p, n ? n : 16));
    ptrs.push_back(p);
    bytes += n;
    return p;
  }
  ~DevicePool() {
    for (void* p : ptrs) cudaFree(p);
  }
};

struct RawTensor {  // a checkpoint tensor staged on the device as fp32, PyTorch layout
  float* p = nullptr;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

struct UpConvW {  // nearest-2x upsample + 3x3 conv folded into four 2x2 convs (ops.cuh pack_upconv_kernel)
  PackedWeight cls[4];  // parity class 2 py + px
};

struct NormW {
  float* gamma = nullptr;
  float* beta = nullptr;
  int C = 0;
};

// ----------------------------------------------------------------------------------------------------------
// Launch context
// ----------------------------------------------------------------------------------------------------------
// Timing ablation (debug only; results are garbage): SDTF_SKIP=conv,attn,gn,ln,misc makes the listed kernel classes
// no-ops, so the in-graph cost of a class is (step time without the flag) - (step time with it).
enum : int { SKIP_CONV = 1, SKIP_ATTN = 2, SKIP_GN = 4, SKIP_LN = 8, SKIP_MISC = 16 };
inline int skip_mask() {
  static int m = -1;
  if (m < 0) {
    m = 0;
    const char* e = getenv("SDTF_SKIP");
    if (e) {
      const std::string v = e;
      if (v.find("conv") != std::string::npos) m |= SKIP_CONV;
      if (v.find("attn") != std::string::npos) m |= SKIP_ATTN;
      if (v.find("gn") != std::string::npos) m |= SKIP_GN;
      if (v.find("ln") != std::string::npos) m |= SKIP_LN;
      if (v.find("misc") != std::string::npos) m |= SKIP_MISC;
    }
  }
  return m;
}

// SDTF_TRACE=1 (debug, eager launches only): every operator is bracketed by CUDA events on the engine's stream and
// one line per launch goes to stderr — kind, shape, microseconds, TFLOP/s and GB/s of algorithmic work — in graph order,
// with the caches in the state the previous operator left them (unlike an ncu replay).  tools/trace_table.py sums it.
// Per-operator-class totals collected while tracing (sdtf_trace_begin / sdtf_trace_end: bench.py's roofline line is the
// FLOP-weighted rate of the conv / linear kernel over ALL its launches of one denoise step, not one hand-picked shape).
struct TraceTotals {
  long long launches[4] = {0, 0, 0, 0};  // conv, attn, gn, ln
  double us[4] = {0, 0, 0, 0}, flop[4] = {0, 0, 0, 0}, bytes[4] = {0, 0, 0, 0};
  bool collecting = false, quiet = false;
  bool paused = false;  // set around work that is not part of what is being accounted (the once-per-job prologue of sdtf_denoise)
  // quiet collection (bench.py): the event pairs of all operators are only read back at sdtf_trace_end, so the stream never
  // drains between operators — an operator's elapsed time is its kernel(s), not kernel + launch latency + host round trip
  struct Pending { int kind; cudaEvent_t e0, e1; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get_event() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    SDTF_CUDA(cudaEventCreate(&e));
    return e;
  }
  void drain() {  // caller has synchronised the stream
    for (auto& pd : pending) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, pd.e0, pd.e1) == cudaSuccess) us[pd.kind] += ms * 1e3;
      pool.push_back(pd.e0); pool.push_back(pd.e1);
    }
    pending.clear();
  }
};
inline TraceTotals& trace_totals() {
  static TraceTotals t;
  return t;
}
inline int trace_level() {
  static const int v = getenv("SDTF_TRACE") ? atoi(getenv("SDTF_TRACE")) : 0;
  return (v == 0 && trace_totals().collecting) ? 1 : v;
}
inline bool trace_on() { return trace_level() != 0; }

// SDTF_TRACE=2 additionally prints an order-independent fingerprint (sum and xor of the bf16 bit patterns) of the LAST
// sample of every operator output: two runs that should agree bit for bit (a sample alone vs inside a batch, graph vs
// eager) can be diffed operator by operator (tools/batch_invariance.py).
__global__ void fingerprint_kernel(const bf16* __restrict__ x, long long ld, int C, long long pixels, unsigned long long* out) {
  unsigned long long s = 0, xo = 0;
  const long long total = pixels * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / C;
    const unsigned short b = *reinterpret_cast<const unsigned short*>(x + p * ld + (i - p * C));
    s += b;
    xo ^= (unsigned long long)b << ((i & 3) * 16);
  }
  atomicAdd(out, s);
  atomicXor(out + 1, xo);
}

struct Ctx {
  cudaStream_t st = nullptr;
  Arena* ws = nullptr;
  bool dry = false;
  int launches = 0;
  GnScratch gn;
  int batch_class = 0;  // samples per denoise step (both CFG branches) when calls see one branch at a time; 0: the call's own batch

  // SDTF_NVTX=1: every operator launch sits in an NVTX range "kind shape" (e.g. "conv 16x64x64 320->320 k3 s1 +res"), so a
  // profile can be filtered by operator (`ncu --nvtx --nvtx-include "attn*"`) instead of by kernel name and launch index
  static bool nvtx_on() {
    static const int v = getenv("SDTF_NVTX") ? atoi(getenv("SDTF_NVTX")) : 0;
    return v != 0;
  }
  template <class F>
  void traced(const char* kind, const std::string& shape, double flop, double bytes, F&& f) {
    struct Range {
      bool on;
      Range(bool o, const char* k, const std::string& s) : on(o) { if (on) nvtxRangePushA((std::string(k) + " " + s).c_str()); }
      ~Range() { if (on) nvtxRangePop(); }
    } range(nvtx_on() && !dry, kind, shape);
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    const bool capturing = cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
    TraceTotals& tq = trace_totals();
    if (tq.collecting && tq.quiet && !dry && !tq.paused) {
      // bench.py's accounting: one event pair per operator, read back after the job.  Under stream capture the records
      // become EXTERNAL event-record nodes of the step graph, so the durations are those of the graph replay the bench
      // times (eager launches of 20 us kernels are CPU-bound: their event intervals include launch gaps); every replay
      // re-records the same events and the last one is read.
      const int k = kind[0] == 'c' ? 0 : kind[0] == 'a' ? 1 : kind[0] == 'g' ? 2 : 3;
      TraceTotals::Pending pd{k, tq.get_event(), tq.get_event()};
      const unsigned flags = capturing ? cudaEventRecordExternal : cudaEventRecordDefault;
      SDTF_CUDA(cudaEventRecordWithFlags(pd.e0, st, flags));
      f();
      SDTF_CUDA(cudaEventRecordWithFlags(pd.e1, st, flags));
      tq.pending.push_back(pd);
      ++tq.launches[k]; tq.flop[k] += flop; tq.bytes[k] += bytes;
      return;
    }
    if (!trace_on() || dry || capturing || (tq.collecting && tq.quiet)) {
      f();
      return;
    }
    static cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (!e0) { SDTF_CUDA(cudaEventCreate(&e0)); SDTF_CUDA(cudaEventCreate(&e1)); }
    SDTF_CUDA(cudaEventRecord(e0, st));
    f();
    SDTF_CUDA(cudaEventRecord(e1, st));
    SDTF_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    SDTF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    TraceTotals& tt = trace_totals();
    if (tt.collecting) {
      const int k = kind[0] == 'c' ? 0 : kind[0] == 'a' ? 1 : kind[0] == 'g' ? 2 : 3;
      ++tt.launches[k]; tt.us[k] += ms * 1e3; tt.flop[k] += flop; tt.bytes[k] += bytes;
    }
    if (!tt.quiet)
      fprintf(stderr, "[trace] %-6s %-44s %9.2f us %8.1f TFLOP/s %8.1f GB/s\n", kind, shape.c_str(), ms * 1e3, flop / (ms * 1e9),
              bytes / (ms * 1e6));
  }
  // fingerprint of the last sample of an NHWC view (SDTF_TRACE=2, eager launches only)
  void fingerprint(const char* kind, const bf16* p, long long ld, int C, long long pixels_per_sample, int B) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (trace_level() < 2 || dry || (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone)) return;
    static unsigned long long* d = nullptr;
    if (!d) SDTF_CUDA(cudaMalloc((void**)&d, 16));
    SDTF_CUDA(cudaMemsetAsync(d, 0, 16, st));
    // token-shaped views fold the batch into the pixel count: SDTF_TRACE_BATCH tells how many samples the call carries
    static const int tb = getenv("SDTF_TRACE_BATCH") ? atoi(getenv("SDTF_TRACE_BATCH")) : 0;
    const long long total = pixels_per_sample * B;
    if (tb > 0) { pixels_per_sample = total / tb; B = tb; }
    fingerprint_kernel<<<64, 256, 0, st>>>(p + (long long)(B - 1) * pixels_per_sample * ld, ld, C, pixels_per_sample, d);
    unsigned long long h[2];
    SDTF_CUDA(cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, st));
    SDTF_CUDA(cudaStreamSynchronize(st));
    fprintf(stderr, "[fp] %-6s C%-5d px%-6lld %016llx %016llx\n", kind, C, pixels_per_sample, h[0], h[1]);
  }

  View alloc_view(int B, int H, int W, int C) {
    View v;
    v.p = ws->alloc_n<bf16>((size_t)B * H * W * C);
    v.B = B; v.H = H; v.W = W; v.C = C; v.ld = C;
    return v;
  }

  void conv(const ConvArgs& a0) {
    ++launches;
    // split-K scratch (fp32 partial sums) lives in the arena for the duration of the launch pair
    ConvArgs a = a0;
    if (!a.batch_class) a.batch_class = batch_class;
    const size_t sk_floats = a.splitk_ws ? 0 : conv_splitk_floats(a);
    const size_t sk_mark = ws->mark();
    if (sk_floats) {
      a.splitk_ws = ws->alloc_n<float>(sk_floats);
      ++launches;
    }
    struct Release { Arena* w; size_t m; ~Release() { w->release(m); } } rel{ws, sk_mark};
    if (dry || (skip_mask() & SKIP_CONV)) return;
    const PackedWeight& w = *a.w;
    const int K = a.a0.C + (a.a1.p ? a.a1.C : 0);
    const double M = (double)a.a0.B * a.outH * a.outW;
    const int Nout = a.act == ACT_GEGLU ? w.N / 2 : w.N;
    char buf[96];
    snprintf(buf, sizeof buf, "%dx%dx%d %d%s->%d k%d s%d%s%s", a.a0.B, a.outH, a.outW, K, a.a1.p ? "(cat)" : "", w.N, w.kh, a.stride,
             a.res ? " +res" : "", a.act == ACT_GEGLU ? " geglu" : "");
    traced("conv", buf, 2.0 * M * w.N * K * w.kh * w.kw,
           2.0 * ((double)a.a0.B * a.a0.H * a.a0.W * K + M * Nout * (a.res ? 2 : 1) + (double)w.N * K * w.kh * w.kw),
           [&] { launch_conv(st, a); });
    if (!a.out_fp32 && a.out_step == 1 && a.out_head_stride == 0) fingerprint("conv", reinterpret_cast<const bf16*>(a.out), a.out_ld, Nout, (long long)a.outH * a.outW, a.a0.B);
  }
  // y = conv(x) (+bias) (+temb) (+res) ; out view may be a channel slice
  // linear whose output columns are attention heads stored `head_stride` columns apart (d real columns each): the
  // padding columns are neither computed-and-written nor read (clipped TMA store / loads)
  static ConvArgs args_heads(const View& x, const PackedWeight& w, const View& out, int d, int head_stride) {
    ConvArgs a;
    a.a0 = x; a.w = &w;
    a.outH = out.H; a.outW = out.W;
    a.out = out.p; a.out_ld = out.ld;
    static const int clip = getenv("SDTF_HEAD_CLIP") ? atoi(getenv("SDTF_HEAD_CLIP")) : 1;  // A/B: 0 = write the (zero) padding columns too
    if (head_stride != d && clip) { a.out_head_d = d; a.out_head_stride = head_stride; }
    return a;
  }
  void conv_heads(const View& x, const PackedWeight& w, const View& out, int d, int head_stride) { conv(args_heads(x, w, out, d, head_stride)); }
  void conv(const View& x, const PackedWeight& w, const View& out, int stride = 1, int pad = -1, const View* res = nullptr,
            const float* temb = nullptr, int temb_ld = 0, int act = ACT_NONE, const View* x2 = nullptr) {
    conv(args(x, w, out, stride, pad, res, temb, temb_ld, act, x2));
  }
  static ConvArgs args(const View& x, const PackedWeight& w, const View& out, int stride = 1, int pad = -1, const View* res = nullptr,
                       const float* temb = nullptr, int temb_ld = 0, int act = ACT_NONE, const View* x2 = nullptr) {
    ConvArgs a;
    a.a0 = x;
    if (x2) a.a1 = *x2;
    a.w = &w;
    a.stride = stride;
    const int p = pad >= 0 ? pad : (w.kh == 3 ? 1 : 0);
    a.pad_t = a.pad_l = p;
    a.outH = out.H; a.outW = out.W;
    a.temb = temb; a.temb_ld = temb_ld;
    if (res) { a.res = res->p; a.res_ld = res->ld; }
    a.out = out.p; a.out_ld = out.ld;
    a.act = act;
    return a;
  }
  // y = conv3x3(upsample2x(x)) without the upsampled tensor: one 2x2 conv per output parity, each writing every other
  // pixel of y through a strided TMA store map
  void upconv(const View& x, const UpConvW& w, const View& y) {
    for (int cls = 0; cls < 4; ++cls) {
      const int py = cls >> 1, px = cls & 1;
      ConvArgs a;
      a.a0 = x;
      a.w = &w.cls[cls];
      a.pad_t = 1 - py; a.pad_l = 1 - px;
      a.outH = x.H; a.outW = x.W;
      a.out = y.p + ((long long)py * y.W + px) * y.ld;
      a.out_ld = y.ld;
      a.out_step = 2;
      conv(a);
    }
  }
  void groupnorm(const View& x, const NormW& n, bool silu, const View& y) {
    ++launches;
    if (dry || (skip_mask() & SKIP_GN)) return;
    char buf[64];
    snprintf(buf, sizeof buf, "%dx%dx%dx%d%s", x.B, x.H, x.W, x.C, silu ? " silu" : "");
    traced("gn", buf, 0.0, 4.0 * x.pixels() * x.C, [&] { launch_groupnorm(st, x, n.gamma, n.beta, silu, y.p, y.ld, gn, batch_class); });
    fingerprint("gn", y.p, y.ld, x.C, (long long)x.H * x.W, x.B);
  }
  void layernorm(const View& x, const NormW& n, const View& y) {
    ++launches;
    if (dry || (skip_mask() & SKIP_LN)) return;
    char buf[64];
    snprintf(buf, sizeof buf, "%lldx%d", x.pixels(), x.C);
    traced("ln", buf, 0.0, 4.0 * x.pixels() * x.C,
           [&] { launch_layernorm(st, x.p, x.ld, x.C, x.pixels(), n.gamma, n.beta, y.p, y.ld); });
  }
  void attention(const AttnArgs& a) {
    ++launches;
    if (dry || (skip_mask() & SKIP_ATTN)) return;
    char buf[64];
    snprintf(buf, sizeof buf, "B%d h%d Nq%d Nk%d d%d", a.B, a.heads, a.Nq, a.Nk, a.d);
    traced("attn", buf, 4.0 * a.B * a.heads * (double)a.Nq * a.Nk * a.d,
           2.0 * a.B * a.heads * a.d * (2.0 * a.Nq + 2.0 * a.Nk), [&] { launch_attn(st, a); });
  }
  // VAE mid-block attention: qkv (B, N, 1536) -> o (B, N, 512)
  void vattn(const View& qkv, int N, const float* v_bias, const View& o) {
    ++launches;
    if (dry || (skip_mask() & SKIP_ATTN)) return;
    char buf[64];
    snprintf(buf, sizeof buf, "B%d h1 Nq%d Nk%d d512", qkv.B, N, N);
    traced("attn", buf, 4.0 * qkv.B * (double)N * N * 512, 2.0 * qkv.B * N * 512 * 4.0,
           [&] { launch_vattn(st, qkv.p, qkv.ld, qkv.B, N, v_bias, o.p, o.ld); });
  }
  void add_inplace(const View& y, const bf16* c) {
    ++launches;
    if (dry) return;
    long long total = y.pixels() * (y.C / 8);
    long long blocks = ceil_div_ll(total, 256 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    add_inplace_kernel<<<(unsigned)blocks, 256, 0, st>>>(y.p, y.ld, y.C, y.pixels(), c);
    SDTF_CUDA(cudaGetLastError());
  }
  void skinny(const float* x, int M, int K, const bf16* W, const float* bias, int N, bool silu, float* out, int ldo) {
    ++launches;
    if (!dry) launch_skinny_linear(st, x, M, K, W, bias, N, silu, out, ldo);
  }
  void cast_pad(const float* x, long long pixels, int Cs, int Cd, float scale, bf16* y, bool dup) {
    ++launches;
    if (!dry) launch_cast_pad(st, x, pixels, Cs, Cd, scale, y, dup);
  }
  void cast_out(const bf16* x, long long ld, long long pixels, int Cs, float* y) {
    ++launches;
    if (!dry) launch_cast_out(st, x, ld, pixels, Cs, y);
  }
  void memset0(void* p, size_t bytes) {
    if (!dry) SDTF_CUDA(cudaMemsetAsync(p, 0, bytes, st));
  }
};

// ----------------------------------------------------------------------------------------------------------
// Weight store: staged raw tensors by checkpoint key, and helpers that pack them for the kernels
// ----------------------------------------------------------------------------------------------------------
struct WeightStore {
  std::unordered_map<std::string, RawTensor> raw;
  // packed weights live in one pool per component, so that re-finalising a component (another checkpoint, LoRA-merged
  // weights) frees what the previous build of that component allocated
  std::map<std::string, std::unique_ptr<DevicePool>> pools;
  struct PoolRef {
    DevicePool* p = nullptr;
    void* alloc(size_t n) {
      SDTF_CHECK(p != nullptr, "WeightStore::begin_component() was not called");
      return p->alloc(n);
    }
  } pool;
  void begin_component(const std::string& name) {
    pools[name].reset(new DevicePool());
    pool.p = pools[name].get();
  }
  cudaStream_t st = nullptr;
  std::vector<std::string> missing;

  const RawTensor* find(const std::string& key) {
    auto it = raw.find(key);
    if (it == raw.end()) {
      missing.push_back(key);
      return nullptr;
    }
    return &it->second;
  }
  int* upload_map(const std::vector<int>& m) {
    int* d = nullptr;
    SDTF_CUDA(cudaMallocAsync(&d, m.size() * sizeof(int), st));
    SDTF_CUDA(cudaMemcpyAsync(d, m.data(), m.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    SDTF_CUDA(cudaStreamSynchronize(st));  // the host vector may die right after this call
    return d;
  }
  void free_map(int* d) { SDTF_CUDA(cudaFreeAsync(d, st)); }

  // pack rows of `key` ([O][I][kh][kw] or [O][I]) into dst [taps][Ntot][Kp] at row n0, column k0
  // `ln` (optional): a LayerNorm over the I input columns is folded in — gamma scales the columns here, and the per-row
  // constants c1 / c0 (ops.cuh ln_fold_consts_kernel) are written to ln_c1 / accumulated into ln_c0 at rows n0 + [0, nrows)
  void pack_into(bf16* dst, int Ntot, int Kp, int n0, int k0, const RawTensor& t, const std::vector<int>* row_map, float scale,
                 const NormW* ln = nullptr, float* ln_c1 = nullptr, float* ln_c0 = nullptr) {
    const int O = (int)t.shape[0], I = (int)t.shape[1];
    const int taps = t.shape.size() == 4 ? (int)(t.shape[2] * t.shape[3]) : 1;
    const int nrows = row_map ? (int)row_map->size() : O;
    int* dm = row_map ? upload_map(*row_map) : nullptr;
    const long long total = (long long)taps * nrows * I;
    long long blocks = ceil_div_ll(total, 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (blocks < 1) blocks = 1;
    pack_weight_kernel<<<(unsigned)blocks, 256, 0, st>>>(t.p, I, taps, dm, nrows, n0, Ntot, Kp, k0, scale, dst, ln ? ln->gamma : nullptr);
    SDTF_CUDA(cudaGetLastError());
    if (ln) {
      SDTF_CHECK(taps == 1 && k0 == 0 && scale == 1.f && ln->C == I && ln_c1 && ln_c0, "LayerNorm folding: plain linear over the normalised axis only");
      ln_fold_consts_kernel<<<ceil_div(nrows, 8), 256, 0, st>>>(t.p, I, dm, nrows, n0, ln->gamma, ln->beta, ln_c1, ln_c0);
      SDTF_CUDA(cudaGetLastError());
    }
    if (dm) free_map(dm);
  }
  float* zeros_f32(int n) {
    float* p = (float*)pool.alloc((size_t)n * 4);
    SDTF_CUDA(cudaMemsetAsync(p, 0, (size_t)n * 4, st));
    return p;
  }
  float* pack_vec(const RawTensor& t, const std::vector<int>* row_map, float scale, float* dst = nullptr, int n0 = 0) {
    const int n = row_map ? (int)row_map->size() : (int)t.numel();
    if (!dst) dst = (float*)pool.alloc((size_t)n * 4);
    int* dm = row_map ? upload_map(*row_map) : nullptr;
    gather_f32_kernel<<<ceil_div(n, 256), 256, 0, st>>>(t.p, dm, n, scale, dst + n0);
    SDTF_CUDA(cudaGetLastError());
    if (dm) free_map(dm);
    return dst;
  }

  // plain conv / linear: key.weight [+ key.bias]
  // `ln`: fold the LayerNorm that precedes this linear into it (gemm.cuh GemmParams::ln_in)
  PackedWeight conv(const std::string& key, bool bias = true, float scale = 1.f, const std::vector<int>* row_map = nullptr,
                    const NormW* ln = nullptr) {
    PackedWeight pw;
    const RawTensor* w = find(key + ".weight");
    const RawTensor* b = bias ? find(key + ".bias") : nullptr;
    if (!w || (bias && !b)) return pw;
    const int I = (int)w->shape[1];
    pw.kh = w->shape.size() == 4 ? (int)w->shape[2] : 1;
    pw.kw = w->shape.size() == 4 ? (int)w->shape[3] : 1;
    pw.N = row_map ? (int)row_map->size() : (int)w->shape[0];
    pw.K = (I + 7) / 8 * 8;
    const size_t n = (size_t)pw.kh * pw.kw * pw.N * pw.K;
    pw.w = (bf16*)pool.alloc(n * 2);
    SDTF_CUDA(cudaMemsetAsync(pw.w, 0, n * 2, st));
    if (b) pw.bias = pack_vec(*b, row_map, scale);
    if (ln && ln->gamma) {
      if (!pw.bias) pw.bias = zeros_f32(pw.N);
      pw.ln_c1 = zeros_f32(pw.N);
      pw.ln_channels = I;
      pack_into(pw.w, pw.N, pw.K, 0, 0, *w, row_map, scale, ln, pw.ln_c1, pw.bias);
    } else {
      pack_into(pw.w, pw.N, pw.K, 0, 0, *w, row_map, scale);
    }
    return pw;
  }
  UpConvW upconv(const std::string& key) {
    UpConvW u;
    const RawTensor* w = find(key + ".weight");
    const RawTensor* b = find(key + ".bias");
    if (!w || !b) return u;
    SDTF_CHECK(w->shape.size() == 4 && w->shape[2] == 3 && w->shape[3] == 3, key + ": the upsampler convolution must be 3x3");
    const int O = (int)w->shape[0], I = (int)w->shape[1], Kp = (I + 7) / 8 * 8;
    bf16* d = (bf16*)pool.alloc((size_t)16 * O * Kp * 2);
    SDTF_CUDA(cudaMemsetAsync(d, 0, (size_t)16 * O * Kp * 2, st));
    pack_upconv_kernel<<<148 * 8, 256, 0, st>>>(w->p, O, I, Kp, d);
    SDTF_CUDA(cudaGetLastError());
    float* bias = pack_vec(*b, nullptr, 1.f);
    for (int c = 0; c < 4; ++c) {
      u.cls[c].w = d + (size_t)c * 4 * O * Kp;
      u.cls[c].bias = bias;
      u.cls[c].N = O; u.cls[c].K = Kp; u.cls[c].kh = u.cls[c].kw = 2;
    }
    return u;
  }
  NormW norm(const std::string& key) {
    NormW n;
    const RawTensor* g = find(key + ".weight");
    const RawTensor* b = find(key + ".bias");
    if (!g || !b) return n;
    n.C = (int)g->numel();
    n.gamma = pack_vec(*g, nullptr, 1.f);
    n.beta = pack_vec(*b, nullptr, 1.f);
    return n;
  }
  // several [O_i][I] matrices stacked along N (q|k|v, k|v); heads optionally zero-padded from d to dpad rows
  PackedWeight stack(const std::vector<std::string>& keys, int heads, int d, int dpad, const NormW* ln = nullptr) {
    PackedWeight pw;
    std::vector<const RawTensor*> ts;
    for (auto& k : keys) ts.push_back(find(k + ".weight"));
    for (auto t : ts)
      if (!t) return pw;
    const int I = (int)ts[0]->shape[1];
    std::vector<int> map;
    for (int h = 0; h < heads; ++h)
      for (int j = 0; j < dpad; ++j) map.push_back(j < d ? h * d + j : -1);
    const int per = heads * dpad;
    pw.N = per * (int)ts.size();
    pw.K = (I + 7) / 8 * 8;
    pw.w = (bf16*)pool.alloc((size_t)pw.N * pw.K * 2);
    SDTF_CUDA(cudaMemsetAsync(pw.w, 0, (size_t)pw.N * pw.K * 2, st));
    if (ln && ln->gamma) {
      pw.bias = zeros_f32(pw.N);
      pw.ln_c1 = zeros_f32(pw.N);
      pw.ln_channels = I;
    } else {
      ln = nullptr;
    }
    for (size_t i = 0; i < ts.size(); ++i) pack_into(pw.w, pw.N, pw.K, (int)i * per, 0, *ts[i], &map, 1.f, ln, pw.ln_c1, pw.bias);
    return pw;
  }
  // GEGLU projection [8C][C]: rows interleaved per 256-wide N tile as [128 value | 128 gate]
  PackedWeight geglu(const std::string& key, const NormW* ln = nullptr) {
    const RawTensor* w = find(key + ".weight");
    PackedWeight pw;
    if (!w) return pw;
    const int N = (int)w->shape[0], half = N / 2;
    std::vector<int> map(N);
    for (int t = 0; t < N / 256; ++t)
      for (int j = 0; j < 128; ++j) {
        map[t * 256 + j] = t * 128 + j;
        map[t * 256 + 128 + j] = half + t * 128 + j;
      }
    pw = conv(key, true, 1.f, &map, ln);
    pw.geglu_half = 128;
    return pw;
  }
  void drop_raw(const std::string& prefix) {
    for (auto it = raw.begin(); it != raw.end();) {
      if (it->first.compare(0, prefix.size(), prefix) == 0) {
        cudaFree(it->second.p);
        it = raw.erase(it);
      } else {
        ++it;
      }
    }
  }
};

}  // namespace sdtf
