// Standalone GPU check + micro-benchmark of conv_gemm_kernel against a naive CUDA reference.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I minsdtf_b200/csrc \
//              tests/cuda/test_gemm.cu -o build/test_gemm
// Run:    build/test_gemm [bench]
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "gemm_host.cuh"
#include "ops.cuh"

using namespace sdtf;

struct Case {
  const char* name;
  int B, H, W, C0, C1, N, kh, kw, stride, pad_t, pad_l, outH, outW;
  bool bias, temb, res, out_fp32;
  int act;
  int force_bn;
  int ld_extra;  // extra channels of pixel stride (tests strided / sliced views)
};

__global__ void ref_conv(const bf16* a0, long long ld0, int C0, const bf16* a1, long long ld1, int C1, const bf16* w,
                         int K, int N, int kh, int kw, int stride, int pad_t, int pad_l, int B, int H, int W, int outH,
                         int outW, const float* bias, const float* temb, int temb_ld, const bf16* res, long long res_ld,
                         float* out /* [pix][Nout] */, int act, int geglu_half) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Nout = act == ACT_GEGLU ? N / 2 : N;
  long long total = (long long)B * outH * outW * Nout;
  if (idx >= total) return;
  int n = idx % Nout;
  long long pix = idx / Nout;
  int ox = pix % outW;
  int oy = (pix / outW) % outH;
  int b = pix / ((long long)outW * outH);
  auto dot = [&](int row) {
    float acc = 0.f;
    for (int r = 0; r < kh; ++r)
      for (int s = 0; s < kw; ++s) {
        int iy = oy * stride + r - pad_t, ix = ox * stride + s - pad_l;
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
        long long ip = ((long long)b * H + iy) * W + ix;
        const bf16* wr = w + ((long long)(r * kw + s) * N + row) * K;
        for (int k = 0; k < C0; ++k) acc += __bfloat162float(a0[ip * ld0 + k]) * __bfloat162float(wr[k]);
        for (int k = 0; k < C1; ++k) acc += __bfloat162float(a1[ip * ld1 + k]) * __bfloat162float(wr[C0 + k]);
      }
    return acc;
  };
  float v;
  if (act == ACT_GEGLU) {
    int tile = n / geglu_half, j = n % geglu_half;
    int rv = tile * 2 * geglu_half + j, rg = rv + geglu_half;
    float val = dot(rv) + (bias ? bias[rv] : 0.f);
    float g = dot(rg) + (bias ? bias[rg] : 0.f);
    v = val * 0.5f * g * (1.f + tanhf(g * 0.7978845608f * (1.f + 0.044715f * g * g)));
  } else {
    v = dot(n);
    if (bias) v += bias[n];
    if (temb) v += temb[(long long)b * temb_ld + n];
    if (res) v += __bfloat162float(res[pix * res_ld + n]);
    if (act == ACT_SILU) v = v / (1.f + expf(-v));
  }
  out[pix * Nout + n] = v;
}

static uint32_t rng_state = 12345;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 32768.f - 1.f;
}

template <class T>
T* dalloc(size_t n) {
  T* p;
  SDTF_CUDA(cudaMalloc(&p, n * sizeof(T)));
  return p;
}
static bf16* upload_bf16(const std::vector<float>& h) {
  std::vector<bf16> t(h.size());
  for (size_t i = 0; i < h.size(); ++i) t[i] = __float2bfloat16(h[i]);
  bf16* d = dalloc<bf16>(h.size());
  SDTF_CUDA(cudaMemcpy(d, t.data(), h.size() * 2, cudaMemcpyHostToDevice));
  return d;
}
static float* upload_f32(const std::vector<float>& h) {
  float* d = dalloc<float>(h.size());
  SDTF_CUDA(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  return d;
}

static bool run_case(const Case& c, bool bench) {
  const int K = c.C0 + c.C1;
  const int Kp = (K + 7) / 8 * 8;
  const long long ipix = (long long)c.B * c.H * c.W, opix = (long long)c.B * c.outH * c.outW;
  const long long ld0 = c.C0 + c.ld_extra, ld1 = c.C1 + c.ld_extra;
  std::vector<float> h;
  h.resize(ipix * ld0);
  for (auto& x : h) x = frand();
  bf16* a0 = upload_bf16(h);
  bf16* a1 = nullptr;
  if (c.C1) {
    h.resize(ipix * ld1);
    for (auto& x : h) x = frand();
    a1 = upload_bf16(h);
  }
  const int taps = c.kh * c.kw;
  h.assign((size_t)taps * c.N * Kp, 0.f);
  const float ws = 1.f / sqrtf((float)K * taps);
  for (int t = 0; t < taps; ++t)
    for (int n = 0; n < c.N; ++n)
      for (int k = 0; k < K; ++k) h[((size_t)t * c.N + n) * Kp + k] = frand() * ws * 1.7f;
  bf16* w = upload_bf16(h);
  float* bias = nullptr;
  if (c.bias) {
    h.resize(c.N);
    for (auto& x : h) x = frand();
    bias = upload_f32(h);
  }
  const int Nout = c.act == ACT_GEGLU ? c.N / 2 : c.N;
  float* temb = nullptr;
  if (c.temb) {
    h.resize((size_t)c.B * Nout);
    for (auto& x : h) x = frand();
    temb = upload_f32(h);
  }
  const long long out_ld = Nout + c.ld_extra;
  bf16* res = nullptr;
  if (c.res) {
    h.resize(opix * out_ld);
    for (auto& x : h) x = frand();
    res = upload_bf16(h);
  }
  void* out = c.out_fp32 ? (void*)dalloc<float>(opix * out_ld) : (void*)dalloc<bf16>(opix * out_ld);
  SDTF_CUDA(cudaMemset(out, 0, opix * out_ld * (c.out_fp32 ? 4 : 2)));
  float* ref = dalloc<float>(opix * Nout);

  PackedWeight pw;
  pw.w = w; pw.bias = bias; pw.K = Kp; pw.N = c.N; pw.kh = c.kh; pw.kw = c.kw;
  pw.geglu_half = c.act == ACT_GEGLU ? 128 : 0;
  ConvArgs a;
  a.a0 = View{a0, c.B, c.H, c.W, c.C0, ld0};
  if (c.C1) a.a1 = View{a1, c.B, c.H, c.W, c.C1, ld1};
  a.w = &pw; a.stride = c.stride; a.pad_t = c.pad_t; a.pad_l = c.pad_l; a.outH = c.outH; a.outW = c.outW;
  a.temb = temb; a.temb_ld = Nout; a.res = res; a.res_ld = out_ld; a.out = out; a.out_ld = out_ld;
  a.out_fp32 = c.out_fp32; a.act = c.act; a.force_bn = c.force_bn;
  // split-K scratch, exactly as the engine's launch context provides it (0 floats: the layer is not split)
  const size_t skf = conv_splitk_floats(a);
  float* skws = skf ? dalloc<float>(skf) : nullptr;
  a.splitk_ws = skws;

  launch_conv(0, a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("CASE %-28s  CUDA ERROR: %s\n", c.name, cudaGetErrorString(e));
    exit(3);
  }
  bool ok = true;
  if (!bench || opix * Nout * (long long)K * taps < 4e11) {
    long long total = opix * Nout;
    ref_conv<<<(unsigned)((total + 255) / 256), 256>>>(a0, ld0, c.C0, a1, ld1, c.C1, w, Kp, c.N, c.kh, c.kw, c.stride,
                                                       c.pad_t, c.pad_l, c.B, c.H, c.W, c.outH, c.outW, bias, temb,
                                                       Nout, res, out_ld, ref, c.act, pw.geglu_half);
    SDTF_CUDA(cudaDeviceSynchronize());
    std::vector<float> hr(opix * Nout);
    SDTF_CUDA(cudaMemcpy(hr.data(), ref, hr.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<float> ho(opix * out_ld);
    if (c.out_fp32) {
      SDTF_CUDA(cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost));
    } else {
      std::vector<bf16> t(ho.size());
      SDTF_CUDA(cudaMemcpy(t.data(), out, t.size() * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < t.size(); ++i) ho[i] = __bfloat162float(t[i]);
    }
    double max_err = 0, max_ref = 0;
    long long bad = 0, first_bad = -1;
    for (long long p = 0; p < opix; ++p)
      for (int n = 0; n < Nout; ++n) {
        float r = hr[p * Nout + n], o = ho[p * out_ld + n];
        double err = fabs((double)r - o);
        double tol = (c.out_fp32 ? 2e-3 : 1e-2) * (1.0 + fabs(r));
        if (!(err <= tol)) {
          if (first_bad < 0) first_bad = p * Nout + n;
          ++bad;
        }
        if (err > max_err) max_err = err;
        if (fabs(r) > max_ref) max_ref = fabs(r);
      }
    // untouched padding columns must stay zero
    long long pad_bad = 0;
    for (long long p = 0; p < opix; ++p)
      for (int n = Nout; n < out_ld; ++n)
        if (ho[p * out_ld + n] != 0.f) ++pad_bad;
    ok = bad == 0 && pad_bad == 0;
    printf("CASE %-28s  %s  max_err %.4g (max_ref %.3g) bad %lld pad_bad %lld", c.name, ok ? "OK  " : "FAIL", max_err,
           max_ref, bad, pad_bad);
    if (first_bad >= 0)
      printf(" first_bad pix %lld n %lld ref %.4f got %.4f", first_bad / Nout, first_bad % Nout, hr[first_bad],
             ho[(first_bad / Nout) * out_ld + first_bad % Nout]);
    printf("\n");
  }
  if (bench) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch_conv(0, a);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) launch_conv(0, a);
    cudaEventRecord(e1);
    SDTF_CUDA(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    double flop = 2.0 * opix * c.N * (double)K * taps;
    printf("BENCH %-27s  %.3f ms  %.1f TFLOP/s\n", c.name, ms, flop / ms * 1e-9);
  }
  cudaFree(skws); cudaFree(a0); cudaFree(a1); cudaFree(w); cudaFree(bias); cudaFree(temb); cudaFree(res); cudaFree(out); cudaFree(ref);
  fflush(stdout);
  return ok;
}

// A sample's conv result must not depend on what it is batched with: run a split-K layer on B samples, then on the
// last sample alone (same device data), and compare bit for bit.
static bool check_batch_invariance(int B, int HW, int C, int N, bool temb, bool res) {
  const long long pix = (long long)B * HW * HW;
  std::vector<float> h(pix * C);
  for (auto& x : h) x = frand();
  bf16* a0 = upload_bf16(h);
  const int taps = 9;
  h.assign((size_t)taps * N * C, 0.f);
  for (auto& x : h) x = frand() * 0.01f;
  bf16* w = upload_bf16(h);
  h.resize(N);
  for (auto& x : h) x = frand();
  float* bias = upload_f32(h);
  h.resize((size_t)B * N);
  for (auto& x : h) x = frand();
  float* te = upload_f32(h);
  h.resize(pix * N);
  for (auto& x : h) x = frand();
  bf16* rs = upload_bf16(h);
  bf16* out_all = dalloc<bf16>(pix * N);
  bf16* out_one = dalloc<bf16>((long long)HW * HW * N);
  PackedWeight pw;
  pw.w = w; pw.bias = bias; pw.K = C; pw.N = N; pw.kh = 3; pw.kw = 3;
  auto run = [&](int b0, int nb, bf16* out) {
    ConvArgs a;
    a.a0 = View{a0 + (long long)b0 * HW * HW * C, nb, HW, HW, C, C};
    a.w = &pw; a.stride = 1; a.pad_t = 1; a.pad_l = 1; a.outH = HW; a.outW = HW;
    if (temb) { a.temb = te + (long long)b0 * N; a.temb_ld = N; }
    if (res) { a.res = rs + (long long)b0 * HW * HW * N; a.res_ld = N; }
    a.out = out; a.out_ld = N;
    const size_t skf = conv_splitk_floats(a);
    float* ws = skf ? dalloc<float>(skf) : nullptr;
    a.splitk_ws = ws;
    launch_conv(0, a);
    SDTF_CUDA(cudaDeviceSynchronize());
    cudaFree(ws);
    return skf;
  };
  const size_t f_all = run(0, B, out_all);
  const size_t f_one = run(B - 1, 1, out_one);
  const size_t n = (size_t)HW * HW * N;
  std::vector<bf16> ha(n), ho(n);
  SDTF_CUDA(cudaMemcpy(ha.data(), out_all + (long long)(B - 1) * n, n * 2, cudaMemcpyDeviceToHost));
  SDTF_CUDA(cudaMemcpy(ho.data(), out_one, n * 2, cudaMemcpyDeviceToHost));
  long long diff = 0;
  double maxd = 0;
  for (size_t i = 0; i < n; ++i) {
    const double d = fabs((double)__bfloat162float(ha[i]) - (double)__bfloat162float(ho[i]));
    if (d != 0) ++diff;
    if (d > maxd) maxd = d;
  }
  printf("CASE batch_invariance B%d %dx%d %d->%d%s%s  %s  split scratch %zu / %zu floats, %lld elements differ (max %.4g)\n", B, HW, HW, C, N,
         temb ? " temb" : "", res ? " res" : "", diff == 0 ? "OK  " : "FAIL", f_all, f_one, diff, maxd);
  cudaFree(a0); cudaFree(w); cudaFree(bias); cudaFree(te); cudaFree(rs); cudaFree(out_all); cudaFree(out_one);
  return diff == 0;
}

// Nearest-2x upsample + 3x3 conv folded into four 2x2 parity convolutions (engine.cuh Ctx::upconv, ops.cuh
// pack_upconv_kernel) against the naive definition: upsample, zero-pad, 3x3 conv with the fp32 weights.
__global__ void ref_upconv(const bf16* x, int B, int H, int W, int C, const float* w /* [N][C][3][3] */, const float* bias, int N,
                           float* out /* [B][2H][2W][N] */) {
  const long long total = (long long)B * 4 * H * W * N;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int n = (int)(idx % N);
  long long pix = idx / N;
  const int ox = (int)(pix % (2 * W)), oy = (int)((pix / (2 * W)) % (2 * H)), b = (int)(pix / (4LL * H * W));
  float acc = bias[n];
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      const int uy = oy + r - 1, ux = ox + s - 1;  // coordinates on the upsampled grid
      if (uy < 0 || uy >= 2 * H || ux < 0 || ux >= 2 * W) continue;
      const bf16* px = x + (((long long)b * H + (uy >> 1)) * W + (ux >> 1)) * C;
      for (int k = 0; k < C; ++k) acc += __bfloat162float(px[k]) * w[((long long)n * C + k) * 9 + r * 3 + s];
    }
  out[idx] = acc;
}

static bool check_upconv(int B, int H, int W, int C, int N, int ld_extra) {
  std::vector<float> h((size_t)B * H * W * C);
  for (auto& v : h) v = frand();
  bf16* x = upload_bf16(h);
  h.resize((size_t)N * C * 9);
  const float ws = 1.f / sqrtf((float)C * 9);
  for (auto& v : h) v = frand() * ws * 1.7f;
  float* w = upload_f32(h);
  h.resize(N);
  for (auto& v : h) v = frand();
  float* bias = upload_f32(h);
  bf16* packed = dalloc<bf16>((size_t)16 * N * C);
  pack_upconv_kernel<<<148 * 4, 256>>>(w, N, C, C, packed);
  const long long opix = (long long)B * 4 * H * W, ld = N + ld_extra;
  bf16* out = dalloc<bf16>(opix * ld);
  SDTF_CUDA(cudaMemset(out, 0, opix * ld * 2));
  PackedWeight pw[4];
  for (int cls = 0; cls < 4; ++cls) {
    pw[cls].w = packed + (size_t)cls * 4 * N * C; pw[cls].bias = bias; pw[cls].N = N; pw[cls].K = C; pw[cls].kh = pw[cls].kw = 2;
    ConvArgs a;
    a.a0 = View{x, B, H, W, C, C};
    a.w = &pw[cls];
    a.pad_t = 1 - (cls >> 1); a.pad_l = 1 - (cls & 1);
    a.outH = H; a.outW = W;
    a.out = out + ((long long)(cls >> 1) * 2 * W + (cls & 1)) * ld; a.out_ld = ld; a.out_step = 2;
    launch_conv(0, a);
  }
  float* ref = dalloc<float>(opix * N);
  ref_upconv<<<(unsigned)((opix * N + 255) / 256), 256>>>(x, B, H, W, C, w, bias, N, ref);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CASE upconv  CUDA ERROR: %s\n", cudaGetErrorString(e)); exit(3); }
  std::vector<float> hr(opix * N);
  std::vector<bf16> ho(opix * ld);
  SDTF_CUDA(cudaMemcpy(hr.data(), ref, hr.size() * 4, cudaMemcpyDeviceToHost));
  SDTF_CUDA(cudaMemcpy(ho.data(), out, ho.size() * 2, cudaMemcpyDeviceToHost));
  long long bad = 0, pad_bad = 0;
  double max_err = 0;
  for (long long p = 0; p < opix; ++p) {
    for (int n = 0; n < N; ++n) {
      const double r = hr[p * N + n], o = __bfloat162float(ho[p * ld + n]), err = fabs(r - o);
      if (!(err <= 1e-2 * (1.0 + fabs(r)))) ++bad;
      if (err > max_err) max_err = err;
    }
    for (int n = N; n < ld; ++n)
      if (__bfloat162float(ho[p * ld + n]) != 0.f) ++pad_bad;
  }
  printf("CASE upconv B%d %dx%d->%dx%d %d->%d ldx%d  %s  max_err %.4g bad %lld pad_bad %lld\n", B, H, W, 2 * H, 2 * W, C, N, ld_extra,
         bad == 0 && pad_bad == 0 ? "OK  " : "FAIL", max_err, bad, pad_bad);
  cudaFree(x); cudaFree(w); cudaFree(bias); cudaFree(packed); cudaFree(out); cudaFree(ref);
  return bad == 0 && pad_bad == 0;
}

int main(int argc, char** argv) {
  bool bench = argc > 1 && !strcmp(argv[1], "bench");
  const bool pick = argc > 2 && !strcmp(argv[1], "case");  // case <substr>: only the correctness cases whose name contains substr
  const char* only = argc > 2 ? argv[2] : nullptr;  // bench <substr>: run only the benchmark cases whose name contains substr
  init_gemm_kernels();
  std::vector<Case> cases = {
      // name                      B  H   W   C0   C1   N   kh kw s pt pl oH oW  bias temb res  f32  act  bn ldx
      {"linear_64x64x16",          1, 1, 128,  64,   0,  16, 1, 1, 1, 0, 0, 1, 128, false, false, false, true, ACT_NONE, 16, 0},
      {"linear_k128_n64",          1, 1, 128, 128,   0,  64, 1, 1, 1, 0, 0, 1, 128, false, false, false, true, ACT_NONE, 0, 0},
      {"linear_ragged_m300",       1, 1, 300, 320,   0, 320, 1, 1, 1, 0, 0, 1, 300, true, false, false, false, ACT_NONE, 0, 0},
      {"linear_bn80",              1, 1, 300, 320,   0, 320, 1, 1, 1, 0, 0, 1, 300, true, false, true, false, ACT_NONE, 80, 0},
      {"linear_n1280_res",         1, 1, 512, 640,   0, 1280, 1, 1, 1, 0, 0, 1, 512, true, false, true, false, ACT_NONE, 0, 0},
      {"linear_geglu",             1, 1, 256, 320,   0, 2560, 1, 1, 1, 0, 0, 1, 256, true, false, false, false, ACT_GEGLU, 0, 0},
      {"conv1x1_concat_slice",     2, 8, 8, 128,  64, 128, 1, 1, 1, 0, 0, 8, 8, true, false, false, false, ACT_NONE, 0, 64},
      {"conv3x3_16x16",            2, 16, 16, 64,   0,  64, 3, 3, 1, 1, 1, 16, 16, true, false, false, false, ACT_NONE, 0, 0},
      {"conv3x3_temb_res",         2, 32, 32, 320,  0, 320, 3, 3, 1, 1, 1, 32, 32, true, true, true, false, ACT_NONE, 0, 0},
      {"conv3x3_concat",           2, 16, 16, 128, 64, 160, 3, 3, 1, 1, 1, 16, 16, true, true, false, false, ACT_NONE, 0, 0},
      {"conv3x3_in4_pad8",         2, 64, 64,   8,  0, 320, 3, 3, 1, 1, 1, 64, 64, true, false, false, false, ACT_NONE, 0, 0},
      {"conv3x3_out4_f32",         2, 64, 64, 320,  0,   4, 3, 3, 1, 1, 1, 64, 64, true, false, false, true, ACT_NONE, 0, 0},
      {"conv3x3_s2_pad1",          2, 32, 32, 64,   0,  64, 3, 3, 2, 1, 1, 16, 16, true, false, false, false, ACT_NONE, 0, 0},
      {"conv3x3_s2_pad0(asym)",    1, 32, 32, 128,  0, 128, 3, 3, 2, 0, 0, 16, 16, true, false, false, false, ACT_NONE, 0, 0},
      {"conv3x3_s2_8x8",           2, 16, 16, 64,   0,  96, 3, 3, 2, 1, 1, 8, 8, true, false, false, false, ACT_SILU, 0, 0},
      {"conv3x3_96x96",            1, 96, 96, 64,   0,  64, 3, 3, 1, 1, 1, 96, 96, true, false, false, false, ACT_NONE, 0, 0},
      {"conv3x3_12x12_b3",         3, 12, 12, 64,   0,  64, 3, 3, 1, 1, 1, 12, 12, true, false, false, false, ACT_NONE, 0, 0},
      {"conv3x3_c16_silu",         1, 64, 64,  16,  0,  32, 3, 3, 1, 1, 1, 64, 64, true, false, false, false, ACT_SILU, 0, 0},
      {"conv3x3_n16_hint",         1, 64, 64,   8,  0,  16, 3, 3, 1, 1, 1, 64, 64, true, false, false, false, ACT_SILU, 0, 0},
      {"conv3x3_24x24_b2_res",     2, 24, 24, 128,  0, 320, 3, 3, 1, 1, 1, 24, 24, true, true, true, false, ACT_NONE, 0, 0},
      {"linear_n512_k320",         1, 1, 1000, 320,   0, 512, 1, 1, 1, 0, 0, 1, 1000, false, false, false, false, ACT_NONE, 0, 0},
      {"linear_n1280_bn80_res",    1, 1, 1024, 1280,  0, 1280, 1, 1, 1, 0, 0, 1, 1024, true, false, true, false, ACT_NONE, 80, 0},
      {"linear_geglu_m1000",       1, 1, 1000, 640,   0, 5120, 1, 1, 1, 0, 0, 1, 1000, true, false, false, false, ACT_GEGLU, 0, 0},
      {"conv3x3_96ch_slice_res",   2, 32, 32,  96,  0,  96, 3, 3, 1, 1, 1, 32, 32, true, false, true, false, ACT_NONE, 0, 32},
      // K = 320 linears with many M units per N tile (a one-tile-deep ring: stages == k-iterations)
      {"linear_k320_n960_res",     1, 1, 16384, 320,  0, 960, 1, 1, 1, 0, 0, 1, 16384, true, false, true, false, ACT_NONE, 0, 0},
      {"linear_k320_m16000_n320",  1, 1, 16000, 320,  0, 320, 1, 1, 1, 0, 0, 1, 16000, true, false, false, false, ACT_SILU, 0, 0},
      {"linear_k320_geglu",        1, 1, 16384, 320,  0, 2560, 1, 1, 1, 0, 0, 1, 16384, true, false, false, false, ACT_GEGLU, 0, 0},
      // split-K (few tiles, long K): CTA pairs with BN 320, one-CTA units, concat K loop, residual, SiLU, strided view, stride 2
      {"splitk_8x8_b16_1280_temb", 16, 8, 8, 1280,  0, 1280, 3, 3, 1, 1, 1, 8, 8, true, true, false, false, ACT_NONE, 0, 0},
      {"splitk_8x8_b16_cat_res",   16, 8, 8, 1280, 1280, 1280, 3, 3, 1, 1, 1, 8, 8, true, false, true, false, ACT_NONE, 0, 0},
      {"splitk_8x8_b2_silu",        2, 8, 8, 640,   0, 320, 3, 3, 1, 1, 1, 8, 8, true, false, false, false, ACT_SILU, 0, 0},
      {"splitk_4x4_b5_res_ldx",     5, 4, 4, 1280, 0, 1280, 3, 3, 1, 1, 1, 4, 4, true, true, true, false, ACT_NONE, 0, 32},
      {"splitk_s2_16to8_b16",      16, 16, 16, 1280, 0, 1280, 3, 3, 2, 1, 1, 8, 8, true, false, false, false, ACT_NONE, 0, 0},
      {"splitk_3x5_b3_n640",        3, 3, 5, 1280,  0, 640, 3, 3, 1, 1, 1, 3, 5, true, false, true, false, ACT_NONE, 0, 0},
  };
  std::vector<Case> bench_cases = {
      {"b16_conv3x3_64_320",      16, 64, 64, 320,  0, 320, 3, 3, 1, 1, 1, 64, 64, true, true, false, false, ACT_NONE, 0, 0},
      {"b16_conv3x3_32_640",      16, 32, 32, 640,  0, 640, 3, 3, 1, 1, 1, 32, 32, true, true, false, false, ACT_NONE, 0, 0},
      {"b16_conv3x3_16_1280",     16, 16, 16, 1280, 0, 1280, 3, 3, 1, 1, 1, 16, 16, true, true, false, false, ACT_NONE, 0, 0},
      {"b16_conv3x3_8_1280",      16, 8, 8, 1280,   0, 1280, 3, 3, 1, 1, 1, 8, 8, true, true, false, false, ACT_NONE, 0, 0},
      {"b16_conv3x3_8_2560cat",   16, 8, 8, 1280, 1280, 1280, 3, 3, 1, 1, 1, 8, 8, true, true, false, false, ACT_NONE, 0, 0},
      {"b16_conv3x3_64_960cat",   16, 64, 64, 640, 320, 320, 3, 3, 1, 1, 1, 64, 64, true, true, false, false, ACT_NONE, 0, 0},
      {"b16_geglu_4096x320",       1, 1, 65536, 320, 0, 2560, 1, 1, 1, 0, 0, 1, 65536, true, false, false, false, ACT_GEGLU, 0, 0},
      {"b16_ff2_4096x1280",        1, 1, 65536, 1280, 0, 320, 1, 1, 1, 0, 0, 1, 65536, true, false, true, false, ACT_NONE, 0, 0},
      {"b16_linear_1024x640",      1, 1, 16384, 640, 0, 640, 1, 1, 1, 0, 0, 1, 16384, true, false, true, false, ACT_NONE, 0, 0},
      {"b16_linear_4096x320_res",  1, 1, 65536, 320, 0, 320, 1, 1, 1, 0, 0, 1, 65536, true, false, true, false, ACT_NONE, 0, 0},
      {"b16_qkv_4096x320",         1, 1, 65536, 320, 0, 1536, 1, 1, 1, 0, 0, 1, 65536, false, false, false, false, ACT_NONE, 0, 0},
      {"b16_qkv_1024x640",         1, 1, 16384, 640, 0, 1920, 1, 1, 1, 0, 0, 1, 16384, false, false, false, false, ACT_NONE, 0, 0},
      {"b16_geglu_1024x640",       1, 1, 16384, 640, 0, 5120, 1, 1, 1, 0, 0, 1, 16384, true, false, false, false, ACT_GEGLU, 0, 0},
      {"b2_conv3x3_64_320",        2, 64, 64, 320,  0, 320, 3, 3, 1, 1, 1, 64, 64, true, true, false, false, ACT_NONE, 0, 0},
      {"b2_conv3x3_16_1280",       2, 16, 16, 1280, 0, 1280, 3, 3, 1, 1, 1, 16, 16, true, true, false, false, ACT_NONE, 0, 0},
      {"vae_b4_conv3x3_512_128",   4, 512, 512, 128, 0, 128, 3, 3, 1, 1, 1, 512, 512, true, false, false, false, ACT_NONE, 0, 0},
      {"vae_b4_conv3x3_256_256",   4, 256, 256, 256, 0, 256, 3, 3, 1, 1, 1, 256, 256, true, false, false, false, ACT_NONE, 0, 0},
  };
  int fails = 0;
  try {
    if (pick) {
      for (auto& c : cases)
        if (strstr(c.name, only)) fails += run_case(c, false) ? 0 : 1;
      if (strstr("upconv", only)) fails += check_upconv(2, 8, 8, 128, 128, 0) ? 0 : 1;
    } else if (!only) {
      for (auto& c : cases) fails += run_case(c, false) ? 0 : 1;
      fails += check_batch_invariance(2, 8, 1280, 1280, true, false) ? 0 : 1;
      fails += check_batch_invariance(16, 8, 1280, 1280, false, true) ? 0 : 1;
      fails += check_batch_invariance(3, 4, 1280, 640, true, true) ? 0 : 1;
      fails += check_batch_invariance(2, 16, 640, 640, false, false) ? 0 : 1;  // not split: the plain schedule
      fails += check_upconv(2, 8, 8, 128, 128, 0) ? 0 : 1;
      fails += check_upconv(1, 16, 16, 320, 320, 64) ? 0 : 1;   // output is a channel slice of a wider buffer
      fails += check_upconv(3, 12, 20, 64, 96, 0) ? 0 : 1;      // ragged tiles, rectangular
      fails += check_upconv(1, 64, 64, 256, 256, 0) ? 0 : 1;    // VAE-sized grid
    }
    if (bench && !pick)
      for (auto& c : bench_cases)
        if (!only || strstr(c.name, only)) fails += run_case(c, true) ? 0 : 1;
  } catch (const std::exception& e) {
    printf("EXCEPTION: %s\n", e.what());
    return 2;
  }
  printf("%s (%d failing)\n", fails ? "SOME CASES FAILED" : "ALL CASES PASSED", fails);
  return fails ? 1 : 0;
}
