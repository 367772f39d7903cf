// sdtf.cu — C ABI (include/sdtf.h) over the engine: DLPack marshalling, weight ingest, workspace sizing by dry
// run, the per-model entry points and the whole denoising loop (optionally replayed from one captured CUDA graph).
#include "../../include/sdtf.h"

#include <cuda_fp16.h>

#include <algorithm>
#include <cstring>
#include <functional>

#include "comm.cuh"
#include "models.cuh"

using namespace sdtf;

namespace {

thread_local std::string g_create_error;

enum DType { F32, F16, BF16, U8, I32 };

struct TRef {
  void* data = nullptr;
  bool cuda = false;
  DType dt = F32;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
  size_t bytes() const { return (size_t)numel() * (dt == F32 || dt == I32 ? 4 : dt == U8 ? 1 : 2); }
};

TRef parse(const DLManagedTensor* m, const char* name, int device) {
  SDTF_CHECK(m != nullptr, std::string(name) + ": tensor is NULL");
  const DLTensor& t = m->dl_tensor;
  TRef r;
  r.data = (uint8_t*)t.data + t.byte_offset;
  if (t.device.device_type == kDLCUDA) {
    r.cuda = true;
    SDTF_CHECK(t.device.device_id == device, std::string(name) + ": tensor lives on another GPU than the engine");
  } else {
    SDTF_CHECK(t.device.device_type == kDLCPU || t.device.device_type == kDLCUDAHost,
               std::string(name) + ": unsupported DLPack device type");
  }
  SDTF_CHECK(t.dtype.lanes == 1, std::string(name) + ": vector dtypes unsupported");
  if (t.dtype.code == kDLFloat && t.dtype.bits == 32) r.dt = F32;
  else if (t.dtype.code == kDLFloat && t.dtype.bits == 16) r.dt = F16;
  else if (t.dtype.code == kDLBfloat && t.dtype.bits == 16) r.dt = BF16;
  else if (t.dtype.code == kDLUInt && t.dtype.bits == 8) r.dt = U8;
  else if (t.dtype.code == kDLInt && t.dtype.bits == 32) r.dt = I32;
  else throw Error(std::string(name) + ": unsupported dtype");
  r.shape.assign(t.shape, t.shape + t.ndim);
  if (t.strides) {
    int64_t expect = 1;
    for (int i = t.ndim - 1; i >= 0; --i) {
      SDTF_CHECK(t.shape[i] == 1 || t.strides[i] == expect, std::string(name) + ": tensor must be compact row-major");
      expect *= t.shape[i];
    }
  }
  return r;
}

void expect_shape(const TRef& t, std::initializer_list<int64_t> s, const char* name) {
  bool ok = t.shape.size() == s.size();
  if (ok) {
    size_t i = 0;
    for (auto v : s) {
      if (v >= 0 && t.shape[i] != v) ok = false;
      ++i;
    }
  }
  if (!ok) {
    std::string m = std::string(name) + ": bad shape (";
    for (auto v : t.shape) m += std::to_string(v) + ",";
    m += ") expected (";
    for (auto v : s) m += (v < 0 ? std::string("*") : std::to_string(v)) + ",";
    throw Error(m + ")");
  }
}

__global__ void half_to_float_kernel(const __half* x, float* y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __half2float(x[i]);
}
__global__ void bf16_to_float_kernel(const bf16* x, float* y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __bfloat162float(x[i]);
}

}  // namespace

struct sdtf_engine {
  int device = 0;
  cudaStream_t st = nullptr;
  std::string err;
  WeightStore weights;
  UNetW unet;
  ControlNetW cnet;
  VaeDecW vdec;
  VaeEncW venc;
  TextW text;
  Arena ws;
  GnScratch gn;
  int* step_dev = nullptr;
  Comm comm;
  const char* nccl_path = nullptr;
  sdtf_timings timings{};
  cudaEvent_t ev[4]{};
  // captured step graph
  cudaGraphExec_t graph = nullptr;
  std::string graph_key;
  int graph_launches = 0;

  Ctx make_ctx(bool dry) {
    Ctx c;
    c.st = st; c.ws = &ws; c.dry = dry; c.gn = gn;
    ws.dry = dry;
    return c;
  }
  void drop_graph() {
    if (graph) cudaGraphExecDestroy(graph);
    graph = nullptr;
    graph_key.clear();
  }
  // Run `fn` once in dry mode to find the workspace high-water mark, (re)allocate if needed, then for real.
  void run_sized(const std::function<void(Ctx&)>& fn) {
    ws.off = 0; ws.peak = 0;
    Ctx d = make_ctx(true);
    fn(d);
    const size_t need = ws.peak + (1 << 20);
    ws.off = 0; ws.peak = 0; ws.dry = false;
    if (need > ws.cap) {
      SDTF_CUDA(cudaStreamSynchronize(st));
      drop_graph();
      if (ws.base) SDTF_CUDA(cudaFree(ws.base));
      ws.base = nullptr; ws.cap = 0;
      size_t cap = need + need / 8;
      SDTF_CUDA(cudaMalloc((void**)&ws.base, cap));
      ws.cap = cap;
    }
    Ctx c = make_ctx(false);
    fn(c);
    timings.kernel_launches = c.launches;
  }
  // device fp32 copy of a (host or device) f32 tensor inside the arena
  float* stage_f32(Ctx& c, const TRef& t) {
    SDTF_CHECK(t.dt == F32, "expected a float32 tensor");
    float* d = c.ws->alloc_n<float>((size_t)t.numel());
    if (!c.dry) SDTF_CUDA(cudaMemcpyAsync(d, t.data, t.bytes(), t.cuda ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    return d;
  }
  void emit(Ctx& c, const void* dev, const TRef& out) {
    if (!c.dry) SDTF_CUDA(cudaMemcpyAsync(out.data, dev, out.bytes(), out.cuda ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  }
};

#define SDTF_API_BEGIN                         \
  if (!e) return SDTF_ERR_INVALID;             \
  try {                                        \
    SDTF_CUDA(cudaSetDevice(e->device));
#define SDTF_API_END                                                  \
  }                                                                   \
  catch (const Error& ex) {                                           \
    e->err = ex.what();                                               \
    e->ws.dry = false;                                                \
    const bool cuda = e->err.find("failed:") != std::string::npos;    \
    return cuda ? SDTF_ERR_CUDA : SDTF_ERR_INVALID;                   \
  }                                                                   \
  catch (const std::exception& ex) {                                  \
    e->err = ex.what();                                               \
    e->ws.dry = false;                                                \
    return SDTF_ERR_INTERNAL;                                         \
  }                                                                   \
  return SDTF_OK;

extern "C" {

const char* sdtf_version(void) { return "minsdtf_b200 0.1 (sm_100a)"; }

const char* sdtf_last_error(const sdtf_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int sdtf_create(int32_t device, sdtf_engine** out) {
  if (!out) return SDTF_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  cudaError_t ce = cudaGetDeviceCount(&n);
  if (ce != cudaSuccess || n == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(ce) + " — this engine has no CPU fallback";
    return SDTF_ERR_DEVICE;
  }
  if (device < 0 || device >= n) {
    g_create_error = "device index out of range";
    return SDTF_ERR_INVALID;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    g_create_error = std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                     "; the kernels are sm_100a only (tcgen05/TMEM/TMA) and there is no fallback path";
    return SDTF_ERR_DEVICE;
  }
  sdtf_engine* e = new sdtf_engine();
  e->device = device;
  try {
    SDTF_CUDA(cudaSetDevice(device));
    SDTF_CUDA(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
    e->weights.st = e->st;
    SDTF_CUDA(cudaMalloc((void**)&e->gn.partial, sizeof(double) * 64 * kGnMaxBlk * kGnMaxBatch));
    SDTF_CUDA(cudaMalloc((void**)&e->gn.counters, sizeof(unsigned) * kGnMaxBatch));
    SDTF_CUDA(cudaMalloc((void**)&e->gn.stats, sizeof(float) * 64 * kGnMaxBatch));
    SDTF_CUDA(cudaMemset(e->gn.counters, 0, sizeof(unsigned) * kGnMaxBatch));
    SDTF_CUDA(cudaMalloc((void**)&e->gn.gens, sizeof(unsigned) * kGnMaxBatch));
    SDTF_CUDA(cudaMemset(e->gn.gens, 0, sizeof(unsigned) * kGnMaxBatch));
    SDTF_CUDA(cudaMalloc((void**)&e->step_dev, sizeof(int) * 4));
    for (auto& ev : e->ev) SDTF_CUDA(cudaEventCreate(&ev));
    // kernel attributes are set up-front so that nothing but launches happens under stream capture
    init_gemm_kernels();
    init_attn_kernels();
    init_norm_kernels();
    tensor_map_encoder();
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    delete e;
    return SDTF_ERR_CUDA;
  }
  *out = e;
  return SDTF_OK;
}

void sdtf_destroy(sdtf_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->st);
  e->drop_graph();
  if (e->comm.comm) {
    try { nccl_api(nullptr).CommDestroy(e->comm.comm); } catch (...) {}
    e->comm.comm = nullptr;
  }
  for (auto& kv : e->weights.raw) cudaFree(kv.second.p);
  if (e->ws.base) cudaFree(e->ws.base);
  cudaFree(e->gn.partial);
  cudaFree(e->gn.counters);
  cudaFree(e->gn.gens);
  cudaFree(e->gn.stats);
  cudaFree(e->step_dev);
  for (auto& ev : e->ev) cudaEventDestroy(ev);
  cudaStreamDestroy(e->st);
  delete e;
}

int sdtf_load_tensor(sdtf_engine* e, const char* key, const DLManagedTensor* t) {
  SDTF_API_BEGIN
  SDTF_CHECK(key != nullptr, "key is NULL");
  TRef r = parse(t, key, e->device);
  SDTF_CHECK(r.dt == F32 || r.dt == F16 || r.dt == BF16, std::string(key) + ": weights must be f32/f16/bf16");
  RawTensor raw;
  raw.shape = r.shape;
  const size_t n = (size_t)r.numel();
  SDTF_CUDA(cudaMalloc((void**)&raw.p, n * 4 + 16));
  if (r.dt == F32) {
    SDTF_CUDA(cudaMemcpyAsync(raw.p, r.data, n * 4, r.cuda ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->st));
  } else {
    void* tmp = nullptr;
    SDTF_CUDA(cudaMalloc(&tmp, n * 2 + 16));
    SDTF_CUDA(cudaMemcpyAsync(tmp, r.data, n * 2, r.cuda ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->st));
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 32);
    if (r.dt == F16) half_to_float_kernel<<<blocks, 256, 0, e->st>>>((const __half*)tmp, raw.p, (long long)n);
    else bf16_to_float_kernel<<<blocks, 256, 0, e->st>>>((const bf16*)tmp, raw.p, (long long)n);
    SDTF_CUDA(cudaGetLastError());
    SDTF_CUDA(cudaStreamSynchronize(e->st));
    cudaFree(tmp);
  }
  SDTF_CUDA(cudaStreamSynchronize(e->st));  // host buffer is only borrowed for this call
  auto it = e->weights.raw.find(key);
  if (it != e->weights.raw.end()) cudaFree(it->second.p);
  e->weights.raw[key] = raw;
  SDTF_API_END
}

int sdtf_finalize_weights(sdtf_engine* e, const char* component) {
  SDTF_API_BEGIN
  SDTF_CHECK(component != nullptr, "component is NULL");
  const std::string comp = component;
  WeightStore& w = e->weights;
  w.missing.clear();
  e->drop_graph();
  SDTF_CUDA(cudaStreamSynchronize(e->st));  // nothing may still be reading the weights this call replaces
  SDTF_CHECK(comp == "unet" || comp == "controlnet" || comp == "vae_decoder" || comp == "vae_encoder" || comp == "text_encoder",
             "unknown component '" + comp + "' (unet | controlnet | vae_decoder | vae_encoder | text_encoder)");
  w.begin_component(comp);
  std::vector<std::string> prefixes;
  if (comp == "unet") {
    e->unet = UNetW();
    build_unet(w, e->unet);
    e->unet.ready = w.missing.empty();
    prefixes = {"model.diffusion_model."};
  } else if (comp == "controlnet") {
    e->cnet = ControlNetW();
    build_controlnet(w, e->cnet);
    e->cnet.ready = w.missing.empty();
    prefixes = {"control_model."};
  } else if (comp == "vae_decoder") {
    e->vdec = VaeDecW();
    build_vae_decoder(w, e->vdec);
    e->vdec.ready = w.missing.empty();
    prefixes = {"decoder.", "post_quant_conv."};
  } else if (comp == "vae_encoder") {
    e->venc = VaeEncW();
    build_vae_encoder(w, e->venc);
    e->venc.ready = w.missing.empty();
    prefixes = {"encoder.", "quant_conv."};
  } else if (comp == "text_encoder") {
    e->text = TextW();
    build_text_encoder(w, e->text);
    e->text.ready = w.missing.empty();
    prefixes = {"text_model."};
  } else {
    throw Error("unknown component '" + comp + "' (unet | controlnet | vae_decoder | vae_encoder | text_encoder)");
  }
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  if (!w.missing.empty()) {
    std::string m = comp + ": " + std::to_string(w.missing.size()) + " checkpoint tensors missing, e.g.";
    for (size_t i = 0; i < w.missing.size() && i < 4; ++i) m += " " + w.missing[i];
    throw Error(m);
  }
  for (auto& p : prefixes) w.drop_raw(p);
  SDTF_API_END
}

// -------------------------------------------------------------------------------------------------------
// per-model entry points
// -------------------------------------------------------------------------------------------------------
static bf16* stage_context(sdtf_engine* e, Ctx& c, const TRef& ctx) {
  float* f = e->stage_f32(c, ctx);
  bf16* b = c.ws->alloc_n<bf16>((size_t)ctx.numel());
  c.cast_pad(f, ctx.numel() / kCtxDim, kCtxDim, kCtxDim, 1.f, b, false);
  return b;
}

int sdtf_unet_forward(sdtf_engine* e, const DLManagedTensor* latent, const DLManagedTensor* t_emb,
                      const DLManagedTensor* context, const DLManagedTensor* const* controls, DLManagedTensor* out_eps) {
  SDTF_API_BEGIN
  SDTF_CHECK(e->unet.ready, "unet weights not finalized");
  TRef lat = parse(latent, "latent", e->device), te = parse(t_emb, "t_emb", e->device);
  TRef ctx = parse(context, "context", e->device), out = parse(out_eps, "out_eps", e->device);
  expect_shape(lat, {-1, -1, -1, 4}, "latent");
  const int B = (int)lat.shape[0], h = (int)lat.shape[1], w = (int)lat.shape[2];
  SDTF_CHECK(h % 8 == 0 && w % 8 == 0, "latent height/width must be multiples of 8");
  expect_shape(te, {B, 320}, "t_emb");
  expect_shape(ctx, {B, -1, kCtxDim}, "context");
  expect_shape(out, {B, h, w, 4}, "out_eps");
  SDTF_CHECK(out.dt == F32, "out_eps must be float32");
  const int T = (int)ctx.shape[1];
  static const int lv[13] = {0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 3};
  static const int ch[13] = {320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280, 1280};
  std::vector<TRef> ctr;
  if (controls)
    for (int i = 0; i < 13; ++i) {
      ctr.push_back(parse(controls[i], "control", e->device));
      expect_shape(ctr.back(), {B, h >> lv[i], w >> lv[i], ch[i]}, "control");
    }
  e->run_sized([&](Ctx& c) {
    float* latf = e->stage_f32(c, lat);
    bf16* lat8 = c.ws->alloc_n<bf16>((size_t)B * h * w * 8);
    c.cast_pad(latf, (long long)B * h * w, 4, 8, 1.f, lat8, false);
    float* tef = e->stage_f32(c, te);
    bf16* ctxb = stage_context(e, c, ctx);
    CtxKV kv;
    project_context(c, ctxb, B, T, unet_attn_layers(e->unet), kv);
    std::vector<const bf16*> cptr;
    if (controls)
      for (int i = 0; i < 13; ++i) {
        float* f = e->stage_f32(c, ctr[i]);
        bf16* b = c.ws->alloc_n<bf16>((size_t)ctr[i].numel());
        c.cast_pad(f, ctr[i].numel() / ch[i], ch[i], ch[i], 1.f, b, false);
        cptr.push_back(b);
      }
    float* eps = c.ws->alloc_n<float>((size_t)B * h * w * 4);
    unet_forward(c, e->unet, lat8, B, h, w, tef, kv, controls ? cptr.data() : nullptr, eps);
    e->emit(c, eps, out);
  });
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  SDTF_API_END
}

int sdtf_hintnet_forward(sdtf_engine* e, const DLManagedTensor* image, DLManagedTensor* out_t) {
  SDTF_API_BEGIN
  SDTF_CHECK(e->cnet.ready, "controlnet weights not finalized");
  TRef img = parse(image, "image", e->device), out = parse(out_t, "out", e->device);
  expect_shape(img, {-1, -1, -1, 3}, "image");
  const int B = (int)img.shape[0], H = (int)img.shape[1], W = (int)img.shape[2];
  SDTF_CHECK(H % 64 == 0 && W % 64 == 0, "image height/width must be multiples of 64");
  expect_shape(out, {B, H / 8, W / 8, 320}, "out");
  SDTF_CHECK(out.dt == F32, "out must be float32");
  e->run_sized([&](Ctx& c) {
    float* f = e->stage_f32(c, img);
    bf16* i8 = c.ws->alloc_n<bf16>((size_t)B * H * W * 8);
    c.cast_pad(f, (long long)B * H * W, 3, 8, 1.f, i8, false);
    View hint = c.alloc_view(B, H / 8, W / 8, 320);
    hintnet_forward(c, e->cnet, i8, B, H, W, hint);
    float* o = c.ws->alloc_n<float>((size_t)out.numel());
    c.cast_out(hint.p, 320, hint.pixels(), 320, o);
    e->emit(c, o, out);
  });
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  SDTF_API_END
}

int sdtf_controlnet_forward(sdtf_engine* e, const DLManagedTensor* latent, const DLManagedTensor* t_emb,
                            const DLManagedTensor* context, const DLManagedTensor* hint_t, DLManagedTensor* const* outs) {
  SDTF_API_BEGIN
  SDTF_CHECK(e->cnet.ready, "controlnet weights not finalized");
  SDTF_CHECK(outs != nullptr, "outs is NULL");
  TRef lat = parse(latent, "latent", e->device), te = parse(t_emb, "t_emb", e->device);
  TRef ctx = parse(context, "context", e->device), hin = parse(hint_t, "hint", e->device);
  expect_shape(lat, {-1, -1, -1, 4}, "latent");
  const int B = (int)lat.shape[0], h = (int)lat.shape[1], w = (int)lat.shape[2];
  SDTF_CHECK(h % 8 == 0 && w % 8 == 0, "latent height/width must be multiples of 8");
  expect_shape(te, {B, 320}, "t_emb");
  expect_shape(ctx, {B, -1, kCtxDim}, "context");
  expect_shape(hin, {B, h, w, 320}, "hint");
  const int T = (int)ctx.shape[1];
  static const int lv[13] = {0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 3};
  static const int ch[13] = {320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280, 1280};
  std::vector<TRef> o;
  for (int i = 0; i < 13; ++i) {
    o.push_back(parse(outs[i], "out", e->device));
    expect_shape(o.back(), {B, h >> lv[i], w >> lv[i], ch[i]}, "out");
    SDTF_CHECK(o.back().dt == F32, "outputs must be float32");
  }
  e->run_sized([&](Ctx& c) {
    float* latf = e->stage_f32(c, lat);
    bf16* lat8 = c.ws->alloc_n<bf16>((size_t)B * h * w * 8);
    c.cast_pad(latf, (long long)B * h * w, 4, 8, 1.f, lat8, false);
    float* tef = e->stage_f32(c, te);
    bf16* ctxb = stage_context(e, c, ctx);
    float* hf = e->stage_f32(c, hin);
    View hint = c.alloc_view(B, h, w, 320);
    c.cast_pad(hf, hint.pixels(), 320, 320, 1.f, hint.p, false);
    CtxKV kv;
    project_context(c, ctxb, B, T, encoder_attn_layers(e->cnet.enc), kv);
    View res[13];
    for (int i = 0; i < 13; ++i) res[i] = c.alloc_view(B, h >> lv[i], w >> lv[i], ch[i]);
    controlnet_forward(c, e->cnet, lat8, B, h, w, tef, kv, hint, res);
    for (int i = 0; i < 13; ++i) {
      float* f = c.ws->alloc_n<float>((size_t)o[i].numel());
      c.cast_out(res[i].p, ch[i], res[i].pixels(), ch[i], f);
      e->emit(c, f, o[i]);
    }
  });
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  SDTF_API_END
}

int sdtf_vae_decode(sdtf_engine* e, const DLManagedTensor* latent, DLManagedTensor* out_image) {
  SDTF_API_BEGIN
  SDTF_CHECK(e->vdec.ready, "vae_decoder weights not finalized");
  TRef lat = parse(latent, "latent", e->device), out = parse(out_image, "out_image", e->device);
  expect_shape(lat, {-1, -1, -1, 4}, "latent");
  const int B = (int)lat.shape[0], h = (int)lat.shape[1], w = (int)lat.shape[2];
  expect_shape(out, {B, 8 * h, 8 * w, 3}, "out_image");
  SDTF_CHECK(out.dt == F32, "out_image must be float32");
  e->run_sized([&](Ctx& c) {
    float* latf = e->stage_f32(c, lat);
    float* img = c.ws->alloc_n<float>((size_t)out.numel());
    vae_decode(c, e->vdec, latf, B, h, w, img);
    e->emit(c, img, out);
  });
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  SDTF_API_END
}

int sdtf_text_encode(sdtf_engine* e, const DLManagedTensor* tokens, int clip_skip, DLManagedTensor* out_context) {
  SDTF_API_BEGIN
  SDTF_CHECK(e->text.ready, "text_encoder weights not finalized");
  TRef tok = parse(tokens, "tokens", e->device), out = parse(out_context, "out_context", e->device);
  SDTF_CHECK(tok.dt == I32, "tokens must be int32");
  expect_shape(tok, {-1, -1}, "tokens");
  const int B = (int)tok.shape[0], T = (int)tok.shape[1];
  SDTF_CHECK(T >= 1 && T <= e->text.max_len, "tokens: sequence longer than the position table");
  SDTF_CHECK(clip_skip <= -1 && clip_skip >= -kClipLayers, "clip_skip must be in [-12, -1]");
  expect_shape(out, {B, T, kCtxDim}, "out_context");
  SDTF_CHECK(out.dt == F32, "out_context must be float32");
  e->run_sized([&](Ctx& c) {
    int* d_tok = c.ws->alloc_n<int>((size_t)B * T);
    if (!c.dry) SDTF_CUDA(cudaMemcpyAsync(d_tok, tok.data, tok.bytes(), tok.cuda ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->st));
    float* x = c.ws->alloc_n<float>((size_t)B * T * kCtxDim);
    float* ctx = c.ws->alloc_n<float>((size_t)out.numel());
    text_embed(c, e->text, d_tok, nullptr, 0, B, T, x);
    text_encode(c, e->text, x, B, T, clip_skip, ctx);
    e->emit(c, ctx, out);
  });
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  SDTF_API_END
}

int sdtf_text_embed(sdtf_engine* e, const DLManagedTensor* tokens, const DLManagedTensor* positions, DLManagedTensor* out_embedding) {
  SDTF_API_BEGIN
  SDTF_CHECK(e->text.ready, "text_encoder weights not finalized");
  TRef tok = parse(tokens, "tokens", e->device), out = parse(out_embedding, "out_embedding", e->device);
  SDTF_CHECK(tok.dt == I32, "tokens must be int32");
  expect_shape(tok, {-1, -1}, "tokens");
  const int B = (int)tok.shape[0], T = (int)tok.shape[1];
  SDTF_CHECK(T >= 1 && T <= e->text.max_len, "tokens: sequence longer than the position table");
  TRef pos;
  int pos_rows = 0;
  if (positions) {
    pos = parse(positions, "positions", e->device);
    SDTF_CHECK(pos.dt == I32, "positions must be int32");
    expect_shape(pos, {-1, T}, "positions");
    pos_rows = (int)pos.shape[0];
    SDTF_CHECK(pos_rows == 1 || pos_rows == B, "positions must be (1,T) or (B,T)");
  }
  expect_shape(out, {B, T, kCtxDim}, "out_embedding");
  SDTF_CHECK(out.dt == F32, "out_embedding must be float32");
  e->run_sized([&](Ctx& c) {
    int* d_tok = c.ws->alloc_n<int>((size_t)B * T);
    int* d_pos = positions ? c.ws->alloc_n<int>((size_t)pos_rows * T) : nullptr;
    if (!c.dry) {
      SDTF_CUDA(cudaMemcpyAsync(d_tok, tok.data, tok.bytes(), tok.cuda ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->st));
      if (d_pos) SDTF_CUDA(cudaMemcpyAsync(d_pos, pos.data, pos.bytes(), pos.cuda ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->st));
    }
    float* x = c.ws->alloc_n<float>((size_t)B * T * kCtxDim);
    text_embed(c, e->text, d_tok, d_pos, pos_rows, B, T, x);
    e->emit(c, x, out);
  });
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  SDTF_API_END
}

int sdtf_text_encode_embedded(sdtf_engine* e, const DLManagedTensor* embedding, int clip_skip, DLManagedTensor* out_context) {
  SDTF_API_BEGIN
  SDTF_CHECK(e->text.ready, "text_encoder weights not finalized");
  TRef emb = parse(embedding, "embedding", e->device), out = parse(out_context, "out_context", e->device);
  expect_shape(emb, {-1, -1, kCtxDim}, "embedding");
  const int B = (int)emb.shape[0], T = (int)emb.shape[1];
  SDTF_CHECK(T >= 1 && T <= 128, "embedding: at most 128 tokens per window (causal attention kernel)");
  SDTF_CHECK(clip_skip <= -1 && clip_skip >= -kClipLayers, "clip_skip must be in [-12, -1]");
  expect_shape(out, {B, T, kCtxDim}, "out_context");
  SDTF_CHECK(out.dt == F32, "out_context must be float32");
  e->run_sized([&](Ctx& c) {
    float* x = e->stage_f32(c, emb);
    float* ctx = c.ws->alloc_n<float>((size_t)out.numel());
    text_encode(c, e->text, x, B, T, clip_skip, ctx);
    e->emit(c, ctx, out);
  });
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  SDTF_API_END
}

int sdtf_vae_encode(sdtf_engine* e, const DLManagedTensor* image, DLManagedTensor* out_latent) {
  SDTF_API_BEGIN
  SDTF_CHECK(e->venc.ready, "vae_encoder weights not finalized");
  TRef img = parse(image, "image", e->device), out = parse(out_latent, "out_latent", e->device);
  expect_shape(img, {-1, -1, -1, 3}, "image");
  const int B = (int)img.shape[0], H = (int)img.shape[1], W = (int)img.shape[2];
  SDTF_CHECK(H % 8 == 0 && W % 8 == 0, "image height/width must be multiples of 8");
  expect_shape(out, {B, H / 8, W / 8, 4}, "out_latent");
  SDTF_CHECK(out.dt == F32, "out_latent must be float32");
  e->run_sized([&](Ctx& c) {
    float* f = e->stage_f32(c, img);
    float* lat = c.ws->alloc_n<float>((size_t)out.numel());
    vae_encode(c, e->venc, f, B, H, W, lat);
    e->emit(c, lat, out);
  });
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  SDTF_API_END
}

static StepCoef to_coef(const sdtf_step_coef& s) {
  StepCoef c;
  c.guidance = s.guidance; c.rescale = s.rescale; c.ca = s.ca; c.cb = s.cb; c.cn = s.cn; c.sig_t = s.sig_t; c.noi_t = s.noi_t;
  return c;
}

int sdtf_cfg_sched_step(sdtf_engine* e, const DLManagedTensor* eps_u, const DLManagedTensor* eps_c,
                        const DLManagedTensor* latent_prev, const sdtf_step_coef* coef, const DLManagedTensor* noise,
                        const DLManagedTensor* mask, const DLManagedTensor* init_latent, const DLManagedTensor* init_noise,
                        DLManagedTensor* out_latent) {
  SDTF_API_BEGIN
  SDTF_CHECK(coef != nullptr, "coef is NULL");
  TRef ec = parse(eps_c, "eps_c", e->device), lp = parse(latent_prev, "latent_prev", e->device);
  TRef out = parse(out_latent, "out_latent", e->device);
  expect_shape(ec, {-1, -1, -1, 4}, "eps_c");
  const int B = (int)ec.shape[0], h = (int)ec.shape[1], w = (int)ec.shape[2];
  expect_shape(lp, {B, h, w, 4}, "latent_prev");
  expect_shape(out, {B, h, w, 4}, "out_latent");
  SDTF_CHECK(out.dt == F32, "out_latent must be float32");
  TRef eu, nz, mk, il, in_;
  if (eps_u) { eu = parse(eps_u, "eps_u", e->device); expect_shape(eu, {B, h, w, 4}, "eps_u"); }
  if (noise) { nz = parse(noise, "noise", e->device); expect_shape(nz, {B, h, w, 4}, "noise"); }
  if (mask) {
    SDTF_CHECK(init_latent && init_noise, "mask needs init_latent and init_noise");
    mk = parse(mask, "mask", e->device); expect_shape(mk, {h, w}, "mask");
    il = parse(init_latent, "init_latent", e->device); expect_shape(il, {h, w, 4}, "init_latent");
    in_ = parse(init_noise, "init_noise", e->device); expect_shape(in_, {B, h, w, 4}, "init_noise");
  }
  const StepCoef sc = to_coef(*coef);
  e->run_sized([&](Ctx& c) {
    float* d_ec = e->stage_f32(c, ec);
    float* d_lp = e->stage_f32(c, lp);
    float* d_eu = eps_u ? e->stage_f32(c, eu) : nullptr;
    float* d_nz = noise ? e->stage_f32(c, nz) : nullptr;
    float* d_mk = mask ? e->stage_f32(c, mk) : nullptr;
    float* d_il = mask ? e->stage_f32(c, il) : nullptr;
    float* d_in = mask ? e->stage_f32(c, in_) : nullptr;
    StepCoef* d_sc = c.ws->alloc_n<StepCoef>(1);
    float* d_out = c.ws->alloc_n<float>((size_t)out.numel());
    ++c.launches;
    if (!c.dry) {
      SDTF_CUDA(cudaMemcpyAsync(d_sc, &sc, sizeof(sc), cudaMemcpyHostToDevice, e->st));
      SDTF_CUDA(cudaStreamSynchronize(e->st));  // &sc is a stack variable
      cfg_sched_kernel<<<B, 512, 0, e->st>>>(d_eu, d_ec, d_lp, d_sc, nullptr, d_nz, d_mk, d_il, d_in, h * w * 4, d_out, nullptr, B, 0);
      SDTF_CUDA(cudaGetLastError());
    }
    e->emit(c, d_out, out);
  });
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  SDTF_API_END
}

int sdtf_to_uint8(sdtf_engine* e, const DLManagedTensor* decoded, const DLManagedTensor* blend_image,
                  const DLManagedTensor* blend_mask, DLManagedTensor* out_u8) {
  SDTF_API_BEGIN
  TRef d = parse(decoded, "decoded", e->device), out = parse(out_u8, "out_u8", e->device);
  expect_shape(d, {-1, -1, -1, 3}, "decoded");
  const int B = (int)d.shape[0], H = (int)d.shape[1], W = (int)d.shape[2];
  expect_shape(out, {B, H, W, 3}, "out_u8");
  SDTF_CHECK(out.dt == U8, "out_u8 must be uint8");
  TRef bi, bm;
  if (blend_image) {
    SDTF_CHECK(blend_mask != nullptr, "blend_image needs blend_mask");
    bi = parse(blend_image, "blend_image", e->device); expect_shape(bi, {H, W, 3}, "blend_image");
    bm = parse(blend_mask, "blend_mask", e->device); expect_shape(bm, {H, W}, "blend_mask");
  }
  e->run_sized([&](Ctx& c) {
    float* dd = e->stage_f32(c, d);
    float* dbi = blend_image ? e->stage_f32(c, bi) : nullptr;
    float* dbm = blend_image ? e->stage_f32(c, bm) : nullptr;
    uint8_t* o = c.ws->alloc_n<uint8_t>((size_t)out.numel());
    ++c.launches;
    if (!c.dry) {
      to_uint8_kernel<<<148 * 8, 256, 0, e->st>>>(dd, d.numel(), dbi, dbm, (long long)H * W * 3, o);
      SDTF_CUDA(cudaGetLastError());
    }
    e->emit(c, o, out);
  });
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  SDTF_API_END
}

// -------------------------------------------------------------------------------------------------------
// the whole loop: stable_diffusion.py:442-486 on the device
// -------------------------------------------------------------------------------------------------------
int sdtf_denoise(sdtf_engine* e, const sdtf_denoise_desc* d) {
  SDTF_API_BEGIN
  SDTF_CHECK(d != nullptr, "desc is NULL");
  SDTF_CHECK(e->unet.ready, "unet weights not finalized");
  SDTF_CHECK(d->n_steps >= 1 && d->coefs != nullptr, "n_steps >= 1 and coefs required");
  // every float input is staged with a plain byte copy into an fp32 buffer: anything but float32 is refused here
  auto parse_f32 = [&](const DLManagedTensor* t, const char* name) {
    TRef r = parse(t, name, e->device);
    SDTF_CHECK(r.dt == F32, std::string(name) + ": must be float32");
    return r;
  };
  TRef lat0 = parse_f32(d->latent0, "latent0"), ctx = parse_f32(d->context, "context");
  TRef temb = parse_f32(d->t_emb, "t_emb");
  expect_shape(lat0, {-1, -1, -1, 4}, "latent0");
  const int B = (int)lat0.shape[0], h = (int)lat0.shape[1], w = (int)lat0.shape[2], S = d->n_steps;
  SDTF_CHECK(h % 8 == 0 && w % 8 == 0, "latent height/width must be multiples of 8");
  expect_shape(ctx, {B, -1, kCtxDim}, "context");
  const int T = (int)ctx.shape[1];
  expect_shape(temb, {S, 320}, "t_emb");
  const bool cfg = d->uncond_context != nullptr && d->coefs[0].guidance > 0.f;
  const bool split = d->cfg_split != 0;
  if (split) SDTF_CHECK(cfg && e->comm.active() && e->comm.world == 2, "cfg_split needs guidance > 0 and a 2-rank communicator (sdtf_comm_init)");
  const int srank = split ? e->comm.rank : 0;
  TRef uctx, snoise, mask, initl, initn, hint_img, bimg, bmask, oimg, olat;
  // The reference evaluates the two CFG branches as separate model calls (stable_diffusion.py:454-457), so the prompt
  // and the negative prompt may span different numbers of 77-token windows.
  if (cfg) { uctx = parse_f32(d->uncond_context, "uncond_context"); expect_shape(uctx, {B, -1, kCtxDim}, "uncond_context"); }
  const int Tu = cfg ? (int)uctx.shape[1] : T;
  if (d->step_noise) { snoise = parse_f32(d->step_noise, "step_noise"); expect_shape(snoise, {S, B, h, w, 4}, "step_noise"); }
  const bool inpaint = d->mask != nullptr;
  if (inpaint) {
    SDTF_CHECK(d->init_latent && d->init_noise, "mask needs init_latent and init_noise");
    mask = parse_f32(d->mask, "mask"); expect_shape(mask, {h, w}, "mask");
    initl = parse_f32(d->init_latent, "init_latent"); expect_shape(initl, {h, w, 4}, "init_latent");
    initn = parse_f32(d->init_noise, "init_noise"); expect_shape(initn, {B, h, w, 4}, "init_noise");
  }
  const bool control = d->hint_image != nullptr;
  if (control) {
    SDTF_CHECK(e->cnet.ready, "controlnet weights not finalized");
    hint_img = parse_f32(d->hint_image, "hint_image"); expect_shape(hint_img, {B, 8 * h, 8 * w, 3}, "hint_image");
  }
  if (d->decode) {
    SDTF_CHECK(e->vdec.ready, "vae_decoder weights not finalized");
    SDTF_CHECK(d->out_images != nullptr, "decode needs out_images");
    oimg = parse(d->out_images, "out_images", e->device); expect_shape(oimg, {B, 8 * h, 8 * w, 3}, "out_images");
    SDTF_CHECK(oimg.dt == U8, "out_images must be uint8");
    if (d->blend_image) {
      SDTF_CHECK(d->blend_mask != nullptr, "blend_image needs blend_mask");
      bimg = parse_f32(d->blend_image, "blend_image"); expect_shape(bimg, {8 * h, 8 * w, 3}, "blend_image");
      bmask = parse_f32(d->blend_mask, "blend_mask"); expect_shape(bmask, {8 * h, 8 * w}, "blend_mask");
    }
  }
  if (d->out_latent) { olat = parse_f32(d->out_latent, "out_latent"); expect_shape(olat, {B, h, w, 4}, "out_latent"); }
  std::vector<StepCoef> coefs(S);
  for (int i = 0; i < S; ++i) coefs[i] = to_coef(d->coefs[i]);
  // How the CFG pair is evaluated on THIS rank:
  //   dup       one UNet pass over [uncond B | cond B] (equal context lengths; the default)
  //   two_pass  contexts of different length: one B-sample pass per branch, as the reference's two model calls
  //   split     2-GPU CFG split: this rank evaluates one branch, the epsilons are all-gathered
  const bool two_pass = cfg && !split && Tu != T;
  const bool dup = cfg && !split && !two_pass;
  const int Bt = dup ? 2 * B : B;  // samples per UNet pass
  const int n = h * w * 4;
  const std::string key = std::to_string(B) + "x" + std::to_string(h) + "x" + std::to_string(w) + "x" + std::to_string(T) + "x" +
                          std::to_string(Tu) + (cfg ? (split ? (srank ? "S" : "s") : (two_pass ? "t" : "c")) : "-") +
                          (control ? "n" : "-") + (inpaint ? "m" : "-") + (d->step_noise ? "z" : "-") + "s" + std::to_string(S);
  // SDTF_TRACE printing needs eager launches (a host read-back per operator); the quiet accounting of sdtf_trace_begin
  // rides inside the captured graph as external event-record nodes
  const bool accounting = trace_totals().collecting && trace_totals().quiet;
  const bool use_graph = d->use_cuda_graph != 0 && (!trace_on() || accounting);

  e->run_sized([&](Ctx& c) {
    trace_totals().paused = true;  // operator accounting (sdtf_trace_begin) covers the steps and the decode, not the once-per-job prologue
    // ---- job-resident buffers (same addresses for the same configuration => the captured graph stays valid) ----
    float* latent = c.ws->alloc_n<float>((size_t)B * n);
    bf16* lat8 = c.ws->alloc_n<bf16>((size_t)Bt * h * w * 8);
    float* eps = c.ws->alloc_n<float>((size_t)(cfg ? 2 * B : B) * n);  // [uncond B | cond B] (split: gathered from both ranks)
    float* temb_tab = c.ws->alloc_n<float>((size_t)S * 320);
    StepCoef* d_coefs = c.ws->alloc_n<StepCoef>(S);
    float* d_snoise = d->step_noise ? c.ws->alloc_n<float>((size_t)S * B * n) : nullptr;
    float* d_mask = inpaint ? c.ws->alloc_n<float>((size_t)h * w) : nullptr;
    float* d_initl = inpaint ? c.ws->alloc_n<float>((size_t)n) : nullptr;
    float* d_initn = inpaint ? c.ws->alloc_n<float>((size_t)B * n) : nullptr;
    View hint;
    if (control) hint = c.alloc_view(Bt, h, w, 320);
    auto copy_in = [&](void* dst, const TRef& t) {
      if (!c.dry) SDTF_CUDA(cudaMemcpyAsync(dst, t.data, t.bytes(), t.cuda ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->st));
    };
    if (!c.dry) SDTF_CUDA(cudaEventRecord(e->ev[0], e->st));
    copy_in(latent, lat0);
    copy_in(temb_tab, temb);
    if (!c.dry) SDTF_CUDA(cudaMemcpyAsync(d_coefs, coefs.data(), sizeof(StepCoef) * S, cudaMemcpyHostToDevice, e->st));
    if (d_snoise) copy_in(d_snoise, snoise);
    if (inpaint) { copy_in(d_mask, mask); copy_in(d_initl, initl); copy_in(d_initn, initn); }
    c.memset0(e->step_dev, sizeof(int));

    // ---- UNet passes of one step: contexts -> bf16, K/V projections hoisted out of the step loop ----
    struct Pass {
      int T = 0;
      size_t eps_off = 0;
      CtxKV kv, kv_cn;
    };
    std::vector<Pass> passes;
    auto add_pass = [&](const TRef* first, const TRef* second, int Tp, size_t eps_off) {
      Pass ps;
      ps.T = Tp; ps.eps_off = eps_off;
      bf16* ctxb = c.ws->alloc_n<bf16>((size_t)Bt * Tp * kCtxDim);
      const size_t m = c.ws->mark();
      float* f = c.ws->alloc_n<float>((size_t)B * Tp * kCtxDim);
      copy_in(f, *first);
      c.cast_pad(f, (long long)B * Tp, kCtxDim, kCtxDim, 1.f, ctxb, false);
      if (second) {
        copy_in(f, *second);
        c.cast_pad(f, (long long)B * Tp, kCtxDim, kCtxDim, 1.f, ctxb + (size_t)B * Tp * kCtxDim, false);
      }
      c.ws->release(m);
      project_context(c, ctxb, Bt, Tp, unet_attn_layers(e->unet), ps.kv);
      if (control) project_context(c, ctxb, Bt, Tp, encoder_attn_layers(e->cnet.enc), ps.kv_cn);
      passes.push_back(std::move(ps));
    };
    if (dup) add_pass(&uctx, &ctx, T, 0);
    else if (two_pass) { add_pass(&uctx, nullptr, Tu, 0); add_pass(&ctx, nullptr, T, (size_t)B * n); }
    else if (split) { if (srank == 0) add_pass(&uctx, nullptr, Tu, 0); else add_pass(&ctx, nullptr, T, (size_t)B * n); }
    else add_pass(&ctx, nullptr, T, 0);
    if (control) {
      const size_t m = c.ws->mark();
      const int H = 8 * h, W = 8 * w;
      float* f = c.ws->alloc_n<float>((size_t)B * H * W * 3);
      copy_in(f, hint_img);
      bf16* i8 = c.ws->alloc_n<bf16>((size_t)Bt * H * W * 8);
      c.cast_pad(f, (long long)B * H * W, 3, 8, 1.f, i8, dup);
      hintnet_forward(c, e->cnet, i8, Bt, H, W, hint);
      c.ws->release(m);
    }
    c.cast_pad(latent, (long long)B * h * w, 4, 8, 1.f, lat8, dup);
    // the time-embedding MLP and the 22 per-ResBlock projections depend on the timestep only: evaluated for all S
    // steps here, once (52 MB of weights streamed once instead of every step); a step copies its row for each sample
    const int ncat_u = e->unet.enc.time.ncat, ncat_c = control ? e->cnet.enc.time.ncat : 0;
    const float* tab_all_u = time_table(c, e->unet.enc.time, temb_tab, S);
    const float* tab_all_c = control ? time_table(c, e->cnet.enc.time, temb_tab, S) : nullptr;
    float* tab_u = c.ws->alloc_n<float>((size_t)Bt * ncat_u);
    float* tab_c = control ? c.ws->alloc_n<float>((size_t)Bt * ncat_c) : nullptr;

    // ---- one denoising step ----
    // (Two half-batch passes in flight on two streams were measured and rejected in round 1: 18.45 ms per step against
    // 17.80 ms for the single 2B pass, profiles/r01_i_dual_stream_ab.log.)
    auto step = [&](Ctx& sc) {
      // kernel schedules that depend on the batch class must see the step's whole sample count, however the CFG pair is
      // evaluated (one 2B pass, two B passes, or one branch per GPU): split / two-pass runs then equal the batched one
      sc.batch_class = cfg ? 2 * B : B;
      ++sc.launches;
      if (!sc.dry) {
        temb_select_kernel<<<ceil_div(Bt * ncat_u, 256), 256, 0, e->st>>>(tab_all_u, e->step_dev, ncat_u, Bt, tab_u);
        SDTF_CUDA(cudaGetLastError());
      }
      if (control) {
        ++sc.launches;
        if (!sc.dry) {
          temb_select_kernel<<<ceil_div(Bt * ncat_c, 256), 256, 0, e->st>>>(tab_all_c, e->step_dev, ncat_c, Bt, tab_c);
          SDTF_CUDA(cudaGetLastError());
        }
      }
      for (Pass& ps : passes) {
        // UNet encoder half, then ControlNet (its zero-convolutions add into the UNet's skip tensors from their own
        // epilogue, diffusion_model.py:230-234), then the UNet's up path
        UNetState us;
        unet_encode(sc, e->unet, lat8, Bt, h, w, nullptr, ps.kv, tab_u, us);
        if (control) controlnet_forward(sc, e->cnet, lat8, Bt, h, w, nullptr, ps.kv_cn, hint, nullptr, tab_c, &us);
        unet_decode(sc, e->unet, Bt, h, w, ps.kv, us, eps + ps.eps_off);
      }
      if (split) {  // C1: in-place all-gather of this rank's branch; both ranks then run the update redundantly
        ++sc.launches;
        if (!sc.dry) {
          NcclApi& api = nccl_api(nullptr);
          SDTF_NCCL(api, api.AllGather(eps + (size_t)srank * B * n, eps, (size_t)B * n, kNcclFloat32, e->comm.comm, e->st));
        }
      }
      sc.launches += 2;
      if (!sc.dry) {
        cfg_sched_kernel<<<B, 512, 0, e->st>>>(cfg ? eps : nullptr, cfg ? eps + (size_t)B * n : eps, latent, d_coefs, e->step_dev,
                                                d_snoise, d_mask, d_initl, d_initn, n, latent, lat8, B, dup ? 1 : 0);
        SDTF_CUDA(cudaGetLastError());
        step_advance_kernel<<<1, 1, 0, e->st>>>(e->step_dev);
        SDTF_CUDA(cudaGetLastError());
      }
    };
    // stable_diffusion.py:476-478: `callback(iteration)` after every iteration.  The callback runs on the calling
    // thread once the step has finished on the device (one stream synchronisation per step, only when requested).
    auto after_step = [&](int s) {
      if (!d->on_step) return;
      SDTF_CUDA(cudaStreamSynchronize(e->st));
      d->on_step(s + 1, d->on_step_user);
    };

    trace_totals().paused = false;
    if (!c.dry) SDTF_CUDA(cudaEventRecord(e->ev[1], e->st));
    const size_t m_loop = c.ws->mark();
    if (c.dry) {
      step(c);
    } else if (!use_graph) {
      for (int s = 0; s < S; ++s) { step(c); after_step(s); }
    } else {
      const std::string gkey = key + (accounting ? "#acct" : "") + "@" + std::to_string((uintptr_t)e->ws.base);
      if (!e->graph || e->graph_key != gkey) {
        e->drop_graph();
        Ctx gc = c;
        gc.launches = 0;
        cudaGraph_t g = nullptr;
        SDTF_CUDA(cudaStreamBeginCapture(e->st, cudaStreamCaptureModeThreadLocal));
        try {
          step(gc);
        } catch (...) {
          cudaStreamEndCapture(e->st, &g);
          if (g) cudaGraphDestroy(g);
          throw;
        }
        SDTF_CUDA(cudaStreamEndCapture(e->st, &g));
        SDTF_CUDA(cudaGraphInstantiate(&e->graph, g, 0));
        SDTF_CUDA(cudaGraphDestroy(g));
        e->graph_key = gkey;
        e->graph_launches = gc.launches;
      }
      for (int s = 0; s < S; ++s) { SDTF_CUDA(cudaGraphLaunch(e->graph, e->st)); after_step(s); }
      c.launches += e->graph_launches * S;
    }
    c.ws->release(m_loop);
    c.batch_class = 0;  // the decode below is a call of its own (B images)
    if (!c.dry) SDTF_CUDA(cudaEventRecord(e->ev[2], e->st));

    if (d->out_latent) e->emit(c, latent, olat);
    if (d->decode) {
      const int H = 8 * h, W = 8 * w;
      float* img = c.ws->alloc_n<float>((size_t)B * H * W * 3);
      vae_decode(c, e->vdec, latent, B, h, w, img);
      float* dbi = nullptr;
      float* dbm = nullptr;
      if (d->blend_image) {
        dbi = c.ws->alloc_n<float>((size_t)H * W * 3);
        dbm = c.ws->alloc_n<float>((size_t)H * W);
        copy_in(dbi, bimg);
        copy_in(dbm, bmask);
      }
      uint8_t* u8 = c.ws->alloc_n<uint8_t>((size_t)B * H * W * 3);
      ++c.launches;
      if (!c.dry) {
        to_uint8_kernel<<<148 * 8, 256, 0, e->st>>>(img, (long long)B * H * W * 3, dbi, dbm, (long long)H * W * 3, u8);
        SDTF_CUDA(cudaGetLastError());
      }
      e->emit(c, u8, oimg);
    }
    if (!c.dry) SDTF_CUDA(cudaEventRecord(e->ev[3], e->st));
  });
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  float ms = 0;
  cudaEventElapsedTime(&ms, e->ev[1], e->ev[2]); e->timings.loop_ms = ms;
  cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]); e->timings.decode_ms = ms;
  cudaEventElapsedTime(&ms, e->ev[0], e->ev[3]); e->timings.total_ms = ms;
  SDTF_API_END
}

int sdtf_comm_unique_id(const char* nccl_lib, void* out_id128) {
  if (!out_id128) return SDTF_ERR_INVALID;
  try {
    NcclApi& api = nccl_api(nccl_lib);
    NcclUniqueId id;
    SDTF_NCCL(api, api.GetUniqueId(&id));
    memcpy(out_id128, &id, sizeof(id));
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return SDTF_ERR_INTERNAL;
  }
  return SDTF_OK;
}

int sdtf_comm_init(sdtf_engine* e, const char* nccl_lib, const void* id128, int32_t rank, int32_t world) {
  SDTF_API_BEGIN
  SDTF_CHECK(id128 != nullptr, "id is NULL");
  SDTF_CHECK(world == 2 && (rank == 0 || rank == 1), "the CFG split is 2-way: world must be 2, rank 0 or 1");
  NcclApi& api = nccl_api(nccl_lib);
  e->drop_graph();
  if (e->comm.comm) {
    api.CommDestroy(e->comm.comm);
    e->comm = Comm();
  }
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NcclComm c = nullptr;
  SDTF_NCCL(api, api.CommInitRank(&c, world, id, rank));
  e->comm.comm = c; e->comm.rank = rank; e->comm.world = world;
  // one eager collective so that NCCL's lazy channel / buffer setup never happens under stream capture
  float* tmp = nullptr;
  SDTF_CUDA(cudaMalloc((void**)&tmp, sizeof(float) * 2 * 256));
  SDTF_CUDA(cudaMemsetAsync(tmp, 0, sizeof(float) * 2 * 256, e->st));
  SDTF_NCCL(api, api.AllGather(tmp + rank * 256, tmp, 256, kNcclFloat32, c, e->st));
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  cudaFree(tmp);
  SDTF_API_END
}

int sdtf_comm_destroy(sdtf_engine* e) {
  SDTF_API_BEGIN
  e->drop_graph();
  if (e->comm.comm) {
    nccl_api(nullptr).CommDestroy(e->comm.comm);
    e->comm = Comm();
  }
  SDTF_API_END
}

int sdtf_get_timings(const sdtf_engine* e, sdtf_timings* out) {
  if (!e || !out) return SDTF_ERR_INVALID;
  *out = e->timings;
  return SDTF_OK;
}

int sdtf_trace_begin(sdtf_engine* e) {
  SDTF_API_BEGIN
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  TraceTotals& t = trace_totals();
  t.drain();
  for (int k = 0; k < 4; ++k) { t.launches[k] = 0; t.us[k] = t.flop[k] = t.bytes[k] = 0; }
  t.collecting = true;
  t.quiet = getenv("SDTF_TRACE") == nullptr;
  SDTF_API_END
}

int sdtf_trace_end(sdtf_engine* e, sdtf_trace_summary* out) {
  SDTF_API_BEGIN
  SDTF_CHECK(out != nullptr, "out is NULL");
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  TraceTotals& t = trace_totals();
  t.drain();
  for (int k = 0; k < 4; ++k) {
    out->launches[k] = t.launches[k]; out->us[k] = t.us[k]; out->flop[k] = t.flop[k]; out->bytes[k] = t.bytes[k];
  }
  t.collecting = false;
  SDTF_API_END
}

int sdtf_bench_conv(sdtf_engine* e, int32_t batch, int32_t hw, int32_t cin, int32_t cout, int32_t ksize, int32_t reps,
                    float* ms_per_launch) {
  SDTF_API_BEGIN
  SDTF_CHECK(ms_per_launch != nullptr && reps >= 1, "bad arguments");
  SDTF_CHECK(cin % 8 == 0 && (ksize == 1 || ksize == 3), "cin % 8 == 0, ksize in {1,3}");
  float result = 0.f;
  e->run_sized([&](Ctx& c) {
    // enough distinct input/output pairs (> 126 MB in total) that a launch does not find its operands in L2
    const size_t pair_bytes = ((size_t)batch * hw * hw * (cin + cout)) * 2;
    int nbuf = (int)((size_t)192 * 1024 * 1024 / (pair_bytes ? pair_bytes : 1)) + 1;
    if (nbuf > 16) nbuf = 16;
    std::vector<View> xs, ys;
    for (int i = 0; i < nbuf; ++i) {
      xs.push_back(c.alloc_view(batch, hw, hw, cin));
      ys.push_back(c.alloc_view(batch, hw, hw, cout));
    }
    PackedWeight pw;
    pw.K = cin; pw.N = cout; pw.kh = pw.kw = ksize;
    pw.w = c.ws->alloc_n<bf16>((size_t)ksize * ksize * cout * cin);
    pw.bias = c.ws->alloc_n<float>(cout);
    if (c.dry) return;
    for (int i = 0; i < nbuf; ++i) SDTF_CUDA(cudaMemsetAsync(xs[i].p, 0x3c, (size_t)xs[i].pixels() * cin * 2, e->st));
    SDTF_CUDA(cudaMemsetAsync(pw.w, 0x3c, (size_t)ksize * ksize * cout * cin * 2, e->st));
    SDTF_CUDA(cudaMemsetAsync(pw.bias, 0, (size_t)cout * 4, e->st));
    for (int i = 0; i < 3; ++i) c.conv(xs[i % nbuf], pw, ys[i % nbuf]);
    SDTF_CUDA(cudaEventRecord(e->ev[0], e->st));
    for (int i = 0; i < reps; ++i) c.conv(xs[i % nbuf], pw, ys[i % nbuf]);
    SDTF_CUDA(cudaEventRecord(e->ev[1], e->st));
    SDTF_CUDA(cudaStreamSynchronize(e->st));
    float ms = 0;
    SDTF_CUDA(cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]));
    result = ms / reps;
  });
  *ms_per_launch = result;
  SDTF_API_END
}


int sdtf_bench_attention(sdtf_engine* e, int32_t batch, int32_t heads, int32_t nq, int32_t nk, int32_t d, int32_t reps,
                         int32_t legacy, float* ms_per_launch) {
  SDTF_API_BEGIN
  SDTF_CHECK(ms_per_launch != nullptr && reps >= 1, "bad arguments");
  float result = 0.f;
  e->run_sized([&](Ctx& c) {
    const int dstride = d, hs = heads * dstride;
    // rotate over enough q/k/v/out sets that a launch does not find its operands in L2 (> 126 MB in total)
    const size_t set_bytes = ((size_t)batch * nq * (hs + heads * d) + (size_t)2 * batch * nk * hs) * 2;
    int nbuf = (int)((size_t)192 * 1024 * 1024 / (set_bytes ? set_bytes : 1)) + 1;
    if (nbuf > 8) nbuf = 8;
    std::vector<AttnArgs> sets;
    for (int i = 0; i < nbuf; ++i) {
      AttnArgs a;
      bf16* q = c.ws->alloc_n<bf16>((size_t)batch * nq * hs);
      bf16* k = c.ws->alloc_n<bf16>((size_t)batch * nk * hs);
      bf16* v = c.ws->alloc_n<bf16>((size_t)batch * nk * hs);
      a.q = q; a.k = k; a.v = v; a.ldq = a.ldk = a.ldv = hs;
      a.B = batch; a.heads = heads; a.Nq = nq; a.Nk = nk; a.d = d; a.dstride = dstride;
      a.out = c.ws->alloc_n<bf16>((size_t)batch * nq * heads * d); a.ldo = heads * d;
      a.legacy = legacy != 0;
      if (!c.dry) {
        SDTF_CUDA(cudaMemsetAsync(q, 0x3c, (size_t)batch * nq * hs * 2, e->st));  // bf16 0x3c3c = 0.0115
        SDTF_CUDA(cudaMemsetAsync(k, 0x3c, (size_t)batch * nk * hs * 2, e->st));
        SDTF_CUDA(cudaMemsetAsync(v, 0x3c, (size_t)batch * nk * hs * 2, e->st));
      }
      sets.push_back(a);
    }
    if (c.dry) return;
    for (int i = 0; i < 3; ++i) c.attention(sets[i % nbuf]);
    SDTF_CUDA(cudaEventRecord(e->ev[0], e->st));
    for (int i = 0; i < reps; ++i) c.attention(sets[i % nbuf]);
    SDTF_CUDA(cudaEventRecord(e->ev[1], e->st));
    SDTF_CUDA(cudaStreamSynchronize(e->st));
    float ms = 0;
    SDTF_CUDA(cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]));
    result = ms / reps;
  });
  *ms_per_launch = result;
  SDTF_API_END
}

// -------------------------------------------------------------------------------------------------------
// test hooks: single kernels behind the same marshalling, so tests/ can check them against torch one by one
// -------------------------------------------------------------------------------------------------------
__global__ void pad_heads_kernel(const float* __restrict__ x, long long rows, int heads, int d, int dstride, bf16* __restrict__ y) {
  const long long total = rows * heads * dstride;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % dstride);
    const long long rh = i / dstride;
    const int hd = (int)(rh % heads);
    const long long r = rh / heads;
    y[i] = __float2bfloat16(j < d ? x[(r * heads + hd) * d + j] : 0.f);
  }
}

int sdtf_test_attention(sdtf_engine* e, const DLManagedTensor* q_t, const DLManagedTensor* k_t, const DLManagedTensor* v_t,
                        int32_t heads, DLManagedTensor* out_t) {
  SDTF_API_BEGIN
  TRef q = parse(q_t, "q", e->device), k = parse(k_t, "k", e->device), v = parse(v_t, "v", e->device);
  TRef out = parse(out_t, "out", e->device);
  SDTF_CHECK(q.shape.size() == 3 && k.shape.size() == 3, "q/k/v must be (B,N,heads*d)");
  const bool legacy = heads < 0;  // negative head count selects the one-tile-per-CTA kernel
  if (legacy) heads = -heads;
  const int B = (int)q.shape[0], Nq = (int)q.shape[1], C = (int)q.shape[2], Nk = (int)k.shape[1];
  const int d = C / heads, dstride = d, hs = heads * dstride;
  expect_shape(k, {B, Nk, C}, "k");
  expect_shape(v, {B, Nk, C}, "v");
  expect_shape(out, {B, Nq, C}, "out");
  if (heads == 1 && d == 512) {  // the VAE mid-block attention kernel: one head of 512, q | k | v interleaved per token
    SDTF_CHECK(Nq == Nk, "the d = 512 kernel is self-attention");
    e->run_sized([&](Ctx& c) {
      View qkv = c.alloc_view(B, 1, Nq, 3 * C), o = c.alloc_view(B, 1, Nq, C);
      const TRef* src[3] = {&q, &k, &v};
      for (int i = 0; i < 3; ++i) {
        float* f = e->stage_f32(c, *src[i]);
        View part = c.alloc_view(B, 1, Nq, C);
        c.cast_pad(f, (long long)B * Nq, C, C, 1.f, part.p, false);
        if (!c.dry) SDTF_CUDA(cudaMemcpy2DAsync(qkv.p + (size_t)i * C, (size_t)3 * C * 2, part.p, (size_t)C * 2, (size_t)C * 2,
                                                 (size_t)B * Nq, cudaMemcpyDeviceToDevice, e->st));
      }
      c.vattn(qkv, Nq, nullptr, o);
      float* fo = c.ws->alloc_n<float>((size_t)B * Nq * C);
      c.cast_out(o.p, C, (long long)B * Nq, C, fo);
      e->emit(c, fo, out);
    });
    SDTF_CUDA(cudaStreamSynchronize(e->st));
    return SDTF_OK;
  }
  e->run_sized([&](Ctx& c) {
    float* fq = e->stage_f32(c, q);
    float* fk = e->stage_f32(c, k);
    float* fv = e->stage_f32(c, v);
    bf16* bq = c.ws->alloc_n<bf16>((size_t)B * Nq * hs);
    bf16* bk = c.ws->alloc_n<bf16>((size_t)B * Nk * hs);
    bf16* bv = c.ws->alloc_n<bf16>((size_t)B * Nk * hs);
    bf16* bo = c.ws->alloc_n<bf16>((size_t)B * Nq * C);
    float* fo = c.ws->alloc_n<float>((size_t)B * Nq * C);
    if (!c.dry) {
      pad_heads_kernel<<<148 * 4, 256, 0, e->st>>>(fq, (long long)B * Nq, heads, d, dstride, bq);
      pad_heads_kernel<<<148 * 4, 256, 0, e->st>>>(fk, (long long)B * Nk, heads, d, dstride, bk);
      pad_heads_kernel<<<148 * 4, 256, 0, e->st>>>(fv, (long long)B * Nk, heads, d, dstride, bv);
      SDTF_CUDA(cudaGetLastError());
    }
    AttnArgs a;
    a.q = bq; a.k = bk; a.v = bv; a.ldq = a.ldk = a.ldv = hs;
    a.B = B; a.heads = heads; a.Nq = Nq; a.Nk = Nk; a.d = d; a.dstride = dstride; a.out = bo; a.ldo = C;
    a.legacy = legacy;
    c.attention(a);
    c.cast_out(bo, C, (long long)B * Nq, C, fo);
    e->emit(c, fo, out);
  });
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  SDTF_API_END
}

// mode 0: GroupNorm(32), 1: GroupNorm + SiLU, 2: LayerNorm over the last axis.  x (B,H,W,C) f32.
int sdtf_test_norm(sdtf_engine* e, const DLManagedTensor* x_t, const DLManagedTensor* gamma_t, const DLManagedTensor* beta_t,
                   int32_t mode, DLManagedTensor* out_t) {
  SDTF_API_BEGIN
  TRef x = parse(x_t, "x", e->device), g = parse(gamma_t, "gamma", e->device), b = parse(beta_t, "beta", e->device);
  TRef out = parse(out_t, "out", e->device);
  SDTF_CHECK(x.shape.size() == 4, "x must be (B,H,W,C)");
  const int B = (int)x.shape[0], H = (int)x.shape[1], W = (int)x.shape[2], C = (int)x.shape[3];
  expect_shape(g, {C}, "gamma");
  expect_shape(b, {C}, "beta");
  expect_shape(out, {B, H, W, C}, "out");
  e->run_sized([&](Ctx& c) {
    float* fx = e->stage_f32(c, x);
    NormW n;
    n.gamma = e->stage_f32(c, g); n.beta = e->stage_f32(c, b); n.C = C;
    View xv = c.alloc_view(B, H, W, C), yv = c.alloc_view(B, H, W, C);
    c.cast_pad(fx, xv.pixels(), C, C, 1.f, xv.p, false);
    if (mode == 2) c.layernorm(xv, n, yv);
    else c.groupnorm(xv, n, mode == 1, yv);
    float* fo = c.ws->alloc_n<float>((size_t)out.numel());
    c.cast_out(yv.p, C, yv.pixels(), C, fo);
    e->emit(c, fo, out);
  });
  SDTF_CUDA(cudaStreamSynchronize(e->st));
  SDTF_API_END
}

}  // extern "C"
