// models.cuh — the reference's model graphs re-expressed as kernel sequences on NHWC bf16 views:
//   UNet       diffusion_model.py:163-283      ControlNet / HintNet   control_net.py:10-107
//   VAE decoder image_decoder.py:22-55         VAE encoder            image_encoder.py:21-48
// Channel concats are never materialised: the two producers write into channel slices of one buffer.
#pragma once
#include "engine.cuh"

namespace sdtf {

static constexpr int kTimeDim = 1280;
static constexpr int kCtxDim = 768;
static constexpr int kHeads = 8;

struct ResW {
  int cin = 0, cout = 0;
  NormW n1, n2;
  PackedWeight c1, c2, sc;
  bool has_sc = false;
  int temb_off = -1;  // column offset in the per-call time-embedding projection table (-1: VAE block)
};

struct AttnW {
  int C = 0, d = 0, dstride = 0, hs = 0;  // hs = heads * dstride
  NormW gn, ln1, ln2, ln3;
  PackedWeight proj_in, qkv, out1, q2, kv2, out2, ff1, ff2, proj_out;
  bool ln_fold = false;  // the LayerNorms are folded into qkv / q2 / ff1 (gamma in the weights, statistics from the producers)
  bool ln3_fold = false; // ... including the one in front of the GEGLU projection (SDTF_LN_FOLD=2 keeps that one a kernel)
};

struct TimeW {
  bf16 *w1 = nullptr, *w2 = nullptr, *wcat = nullptr;
  float *b1 = nullptr, *b2 = nullptr, *bcat = nullptr;
  int ncat = 0;
};

struct EncoderHalfW {  // time embedding + input blocks + middle block (shared shape of UNet and ControlNet)
  TimeW time;
  PackedWeight conv_in;
  ResW res[8];    // input_blocks 1,2,4,5,7,8,10,11
  AttnW attn[6];  // input_blocks 1,2,4,5,7,8
  PackedWeight down[3];
  ResW mid0, mid2;
  AttnW mid_attn;
};

struct UNetW {
  bool ready = false;
  EncoderHalfW enc;
  ResW up_res[12];
  AttnW up_attn[9];  // output_blocks 3..11
  UpConvW up_conv[3];
  NormW out_norm;
  PackedWeight conv_out;
};

struct ControlNetW {
  bool ready = false;
  EncoderHalfW enc;
  PackedWeight zero[13];
  PackedWeight hint[8];
};

struct VaeAttnW {
  NormW gn;
  PackedWeight qkv, proj;  // q | k | v stacked along N (bias on the q and k rows only)
  float* v_bias = nullptr; // added after the attention (rows of softmax sum to 1)
};
struct VaeDecW {
  bool ready = false;
  PackedWeight post_quant, conv_in, conv_out;
  UpConvW up_conv[3];
  ResW mid0, mid1, up[4][3];
  VaeAttnW attn;
  NormW out_norm;
};
struct VaeEncW {
  bool ready = false;
  PackedWeight conv_in, conv_out, quant, down_conv[3];
  ResW mid0, mid1, down[4][2];
  VaeAttnW attn;
  NormW out_norm;
};

// ----------------------------------------------------------------------------------------------------------
// weight builders
// ----------------------------------------------------------------------------------------------------------
inline ResW build_res(WeightStore& ws, const std::string& p, int cin, int cout, bool ldm) {
  ResW r;
  r.cin = cin; r.cout = cout;
  if (ldm) {
    r.n1 = ws.norm(p + ".in_layers.0");
    r.c1 = ws.conv(p + ".in_layers.2");
    r.n2 = ws.norm(p + ".out_layers.0");
    r.c2 = ws.conv(p + ".out_layers.3");
    r.has_sc = cin != cout;
    if (r.has_sc) r.sc = ws.conv(p + ".skip_connection");
  } else {
    r.n1 = ws.norm(p + ".norm1");
    r.c1 = ws.conv(p + ".conv1");
    r.n2 = ws.norm(p + ".norm2");
    r.c2 = ws.conv(p + ".conv2");
    r.has_sc = cin != cout;
    if (r.has_sc) r.sc = ws.conv(p + ".conv_shortcut");
  }
  return r;
}

inline AttnW build_attn(WeightStore& ws, const std::string& p, int C) {
  AttnW a;
  a.C = C;
  a.d = C / kHeads;
  // Heads are stored densely (d = 40: 80-byte rows); the attention kernels' 4-D TMA maps clip a 64-column box at the head
  // size and zero-fill the rest in shared memory, so nothing is padded in HBM: the q | k | v GEMM at the 64x64 level writes
  // 126 MB instead of 201 MB.  SDTF_HEAD_DENSE=0 (A/B) puts d = 40 heads 64 columns apart (one aligned 128-byte row per
  // head and token) with the padding neither written (clipped 5-D TMA store, Ctx::conv_heads; SDTF_HEAD_CLIP=0 writes
  // it) nor read.  Same box, UNet step at batch 16: dense 17.65 ms, padded + clipped 17.95 ms, padded + written 17.91 ms
  // (profiles/r02_e_ab.log).
  static const int dense = getenv("SDTF_HEAD_DENSE") ? atoi(getenv("SDTF_HEAD_DENSE")) : 1;
  a.dstride = (a.d == 40 && !dense) ? 64 : a.d;
  a.hs = kHeads * a.dstride;
  const std::string t = p + ".transformer_blocks.0";
  a.gn = ws.norm(p + ".norm");
  a.proj_in = ws.conv(p + ".proj_in");
  // LayerNorm (diffusion_model.py:84,86,88) is not a kernel of its own: gamma is folded into the linear that follows, the
  // row statistics come out of the epilogue of the GEMM that produces the LayerNorm's input, and the consumer's epilogue
  // applies y = rstd (acc - mean c1) + c0.  Default (2): norm1 -> q|k|v and norm2 -> q are folded, norm3 in front of the
  // GEGLU projection stays a kernel — that GEMM is epilogue-bound already and the extra FMAs cost more than the kernel
  // saves.  Same box, UNet step at batch 16 (profiles/r02_f_ab.log): 2 -> 17.66 ms, 1 (all three folded) -> 17.83 ms,
  // 0 (round 1: three LayerNorm kernels per block) with the other round-2 switches off too -> 18.00 ms.
  static const int fold_env = getenv("SDTF_LN_FOLD") ? atoi(getenv("SDTF_LN_FOLD")) : 2;
  a.ln_fold = fold_env != 0;
  a.ln3_fold = fold_env == 1;
  a.ln1 = ws.norm(t + ".norm1");
  a.qkv = ws.stack({t + ".attn1.to_q", t + ".attn1.to_k", t + ".attn1.to_v"}, kHeads, a.d, a.dstride, a.ln_fold ? &a.ln1 : nullptr);
  a.out1 = ws.conv(t + ".attn1.to_out.0");
  a.ln2 = ws.norm(t + ".norm2");
  a.q2 = ws.stack({t + ".attn2.to_q"}, kHeads, a.d, a.dstride, a.ln_fold ? &a.ln2 : nullptr);
  a.kv2 = ws.stack({t + ".attn2.to_k", t + ".attn2.to_v"}, kHeads, a.d, a.dstride);
  a.out2 = ws.conv(t + ".attn2.to_out.0");
  a.ln3 = ws.norm(t + ".norm3");
  a.ff1 = ws.geglu(t + ".ff.net.0.proj", a.ln3_fold ? &a.ln3 : nullptr);
  a.ff2 = ws.conv(t + ".ff.net.2");
  a.proj_out = ws.conv(p + ".proj_out");
  return a;
}

inline bf16* pack_linear(WeightStore& ws, const std::string& key, float** bias) {
  const RawTensor* w = ws.find(key + ".weight");
  const RawTensor* b = ws.find(key + ".bias");
  if (!w || !b) return nullptr;
  const int N = (int)w->shape[0], K = (int)w->shape[1];
  bf16* d = (bf16*)ws.pool.alloc((size_t)N * K * 2);
  ws.pack_into(d, N, K, 0, 0, *w, nullptr, 1.f);
  *bias = ws.pack_vec(*b, nullptr, 1.f);
  return d;
}

// concatenate every ResBlock's time_emb_proj into one [sum Cout][1280] matrix: one skinny GEMV per call
inline void build_time(WeightStore& ws, const std::string& p, TimeW& t, std::vector<std::pair<std::string, ResW*>>& blocks) {
  t.w1 = pack_linear(ws, p + ".time_embed.0", &t.b1);
  t.w2 = pack_linear(ws, p + ".time_embed.2", &t.b2);
  int total = 0;
  for (auto& b : blocks) total += b.second->cout;
  t.ncat = total;
  t.wcat = (bf16*)ws.pool.alloc((size_t)total * kTimeDim * 2);
  t.bcat = (float*)ws.pool.alloc((size_t)total * 4);
  int off = 0;
  for (auto& b : blocks) {
    const RawTensor* w = ws.find(b.first + ".emb_layers.1.weight");
    const RawTensor* bi = ws.find(b.first + ".emb_layers.1.bias");
    if (w && bi) {
      ws.pack_into(t.wcat, total, kTimeDim, off, 0, *w, nullptr, 1.f);
      ws.pack_vec(*bi, nullptr, 1.f, t.bcat, off);
      // fold the bias of the conv that consumes this projection (in_layers.2, diffusion_model.py:29,48) into the table:
      // its epilogue then adds ONE per-(sample, channel) vector instead of two
      ResW* r = b.second;
      if (r->c1.bias) {
        vec_add_kernel<<<ceil_div(r->cout, 256), 256, 0, ws.st>>>(t.bcat + off, r->c1.bias, r->cout);
        SDTF_CUDA(cudaGetLastError());
        r->c1.bias = nullptr;
      }
    }
    b.second->temb_off = off;
    off += b.second->cout;
  }
}

static const int kEncIdx[8] = {1, 2, 4, 5, 7, 8, 10, 11};
static const int kEncCin[8] = {320, 320, 320, 640, 640, 1280, 1280, 1280};
static const int kEncCout[8] = {320, 320, 640, 640, 1280, 1280, 1280, 1280};
static const int kDownIdx[3] = {3, 6, 9};
static const int kDecCin[12] = {2560, 2560, 2560, 2560, 2560, 1920, 1920, 1280, 960, 960, 640, 640};
static const int kDecCout[12] = {1280, 1280, 1280, 1280, 1280, 1280, 640, 640, 640, 320, 320, 320};

inline void build_encoder_half(WeightStore& ws, const std::string& p, EncoderHalfW& e,
                               std::vector<std::pair<std::string, ResW*>>& tblocks) {
  e.conv_in = ws.conv(p + ".input_blocks.0.0");
  for (int i = 0; i < 8; ++i) {
    const std::string b = p + ".input_blocks." + std::to_string(kEncIdx[i]);
    e.res[i] = build_res(ws, b + ".0", kEncCin[i], kEncCout[i], true);
    tblocks.push_back({b + ".0", &e.res[i]});
    if (i < 6) e.attn[i] = build_attn(ws, b + ".1", kEncCout[i]);
  }
  for (int i = 0; i < 3; ++i) e.down[i] = ws.conv(p + ".input_blocks." + std::to_string(kDownIdx[i]) + ".0.op");
  e.mid0 = build_res(ws, p + ".middle_block.0", 1280, 1280, true);
  tblocks.push_back({p + ".middle_block.0", &e.mid0});
  e.mid_attn = build_attn(ws, p + ".middle_block.1", 1280);
  e.mid2 = build_res(ws, p + ".middle_block.2", 1280, 1280, true);
  tblocks.push_back({p + ".middle_block.2", &e.mid2});
}

inline void build_unet(WeightStore& ws, UNetW& u) {
  const std::string p = "model.diffusion_model";
  std::vector<std::pair<std::string, ResW*>> tb;
  build_encoder_half(ws, p, u.enc, tb);
  int ai = 0, ui = 0;
  for (int i = 0; i < 12; ++i) {
    const std::string b = p + ".output_blocks." + std::to_string(i);
    u.up_res[i] = build_res(ws, b + ".0", kDecCin[i], kDecCout[i], true);
    tb.push_back({b + ".0", &u.up_res[i]});
    int j = 1;
    if (i >= 3) {
      u.up_attn[ai++] = build_attn(ws, b + ".1", kDecCout[i]);
      j = 2;
    }
    if (i == 2 || i == 5 || i == 8) u.up_conv[ui++] = ws.upconv(b + "." + std::to_string(j) + ".conv");
  }
  u.out_norm = ws.norm(p + ".out.0");
  u.conv_out = ws.conv(p + ".out.2");
  build_time(ws, p, u.enc.time, tb);
}

inline void build_controlnet(WeightStore& ws, ControlNetW& c) {
  const std::string p = "control_model";
  std::vector<std::pair<std::string, ResW*>> tb;
  build_encoder_half(ws, p, c.enc, tb);
  for (int i = 0; i < 12; ++i) c.zero[i] = ws.conv(p + ".zero_convs." + std::to_string(i) + ".0");
  c.zero[12] = ws.conv(p + ".middle_block_out.0");
  for (int i = 0; i < 8; ++i) c.hint[i] = ws.conv(p + ".input_hint_block." + std::to_string(2 * i));
  build_time(ws, p, c.enc.time, tb);
}

inline VaeAttnW build_vae_attn(WeightStore& ws, const std::string& p) {
  VaeAttnW a;
  a.gn = ws.norm(p + ".group_norm");
  a.qkv = ws.stack({p + ".query", p + ".key", p + ".value"}, 1, 512, 512);
  const RawTensor* qb = ws.find(p + ".query.bias");
  const RawTensor* kb = ws.find(p + ".key.bias");
  const RawTensor* vb = ws.find(p + ".value.bias");
  if (qb && kb && vb) {
    a.qkv.bias = (float*)ws.pool.alloc(1536 * 4);
    SDTF_CUDA(cudaMemsetAsync(a.qkv.bias, 0, 1536 * 4, ws.st));
    ws.pack_vec(*qb, nullptr, 1.f, a.qkv.bias, 0);
    ws.pack_vec(*kb, nullptr, 1.f, a.qkv.bias, 512);
    a.v_bias = ws.pack_vec(*vb, nullptr, 1.f);
  }
  a.proj = ws.conv(p + ".proj_attn");
  return a;
}

static const int kVaeDecCin[4] = {512, 512, 512, 256};
static const int kVaeDecCout[4] = {512, 512, 256, 128};
static const int kVaeEncCin[4] = {128, 128, 256, 512};
static const int kVaeEncCout[4] = {128, 256, 512, 512};

inline void build_vae_decoder(WeightStore& ws, VaeDecW& d) {
  d.post_quant = ws.conv("post_quant_conv");
  d.conv_in = ws.conv("decoder.conv_in");
  d.mid0 = build_res(ws, "decoder.mid_block.resnets.0", 512, 512, false);
  d.attn = build_vae_attn(ws, "decoder.mid_block.attentions.0");
  d.mid1 = build_res(ws, "decoder.mid_block.resnets.1", 512, 512, false);
  for (int b = 0; b < 4; ++b) {
    for (int r = 0; r < 3; ++r)
      d.up[b][r] = build_res(ws, "decoder.up_blocks." + std::to_string(b) + ".resnets." + std::to_string(r),
                             r == 0 ? kVaeDecCin[b] : kVaeDecCout[b], kVaeDecCout[b], false);
    if (b < 3) d.up_conv[b] = ws.upconv("decoder.up_blocks." + std::to_string(b) + ".upsamplers.0.conv");
  }
  d.out_norm = ws.norm("decoder.conv_norm_out");
  d.conv_out = ws.conv("decoder.conv_out");
}

inline void build_vae_encoder(WeightStore& ws, VaeEncW& e) {
  e.conv_in = ws.conv("encoder.conv_in");
  for (int b = 0; b < 4; ++b) {
    for (int r = 0; r < 2; ++r)
      e.down[b][r] = build_res(ws, "encoder.down_blocks." + std::to_string(b) + ".resnets." + std::to_string(r),
                               r == 0 ? kVaeEncCin[b] : kVaeEncCout[b], kVaeEncCout[b], false);
    if (b < 3) e.down_conv[b] = ws.conv("encoder.down_blocks." + std::to_string(b) + ".downsamplers.0.conv");
  }
  e.mid0 = build_res(ws, "encoder.mid_block.resnets.0", 512, 512, false);
  e.attn = build_vae_attn(ws, "encoder.mid_block.attentions.0");
  e.mid1 = build_res(ws, "encoder.mid_block.resnets.1", 512, 512, false);
  e.out_norm = ws.norm("encoder.conv_norm_out");
  e.conv_out = ws.conv("encoder.conv_out");
  // quant_conv (8->8, 1x1) followed by `split(2)[0] * 0.18215` (image_encoder.py:47): only the mean rows, pre-scaled
  std::vector<int> first4 = {0, 1, 2, 3};
  e.quant = ws.conv("quant_conv", true, 0.18215f, &first4);
}

// ----------------------------------------------------------------------------------------------------------
// graph pieces
// ----------------------------------------------------------------------------------------------------------
// ResBlock (diffusion_model.py:22-51) / VAE ResnetBlock (layers.py:62-80): x -> out (out may be a slice)
inline void res_block(Ctx& c, const ResW& w, const View& x, const View& out, const float* temb_table, int temb_ld) {
  const size_t m = c.ws->mark();
  View t1 = c.alloc_view(x.B, x.H, x.W, w.cin);
  c.groupnorm(x, w.n1, true, t1);
  View h = c.alloc_view(x.B, x.H, x.W, w.cout);
  c.conv(t1, w.c1, h, 1, -1, nullptr, w.temb_off >= 0 ? temb_table + w.temb_off : nullptr, temb_ld);
  View t2 = c.alloc_view(x.B, x.H, x.W, w.cout);
  c.groupnorm(h, w.n2, true, t2);
  if (w.has_sc) {
    View s = c.alloc_view(x.B, x.H, x.W, w.cout);
    c.conv(x, w.sc, s);
    c.conv(t2, w.c2, out, 1, -1, &s);
  } else {
    c.conv(t2, w.c2, out, 1, -1, &x);
  }
  c.ws->release(m);
}

// Attentions (diffusion_model.py:54-78) with TransformerBlock (:81-96), CrossAttention (:99-129), GEGLU (:142-153).
// ctx_kv: [B*T][2*hs] K|V projections of the context for this layer.
inline void attentions(Ctx& c, const AttnW& w, const View& x, const View& out, const bf16* ctx_kv, int T) {
  const size_t m = c.ws->mark();
  const int B = x.B, HW = x.H * x.W, C = w.C;
  const bool fold = w.ln_fold;
  View t = c.alloc_view(B, x.H, x.W, C);
  c.groupnorm(x, w.gn, false, t);
  View h0 = c.alloc_view(B, x.H, x.W, C);
  // LayerNorm statistics: per row, (sum, sum of squares) partials written by the producing GEMM's epilogue
  const long long rows = (long long)B * HW;
  int slots = C / 32;  // one (sum, sum of squares) pair per row and 32 columns, written by the producing GEMM's epilogue
  float2* part = fold ? c.ws->alloc_n<float2>((size_t)slots * rows) : nullptr;
  auto produce = [&](ConvArgs a) {  // a GEMM whose output feeds a LayerNorm
    if (fold) { a.ln_out = part; a.ln_slots_out = &slots; }
    c.conv(a);
  };
  auto consume = [&](ConvArgs a) {  // a GEMM whose A operand is a LayerNorm input
    if (fold) { a.ln_in = part; a.ln_slots = slots; }
    c.conv(a);
  };
  produce(Ctx::args(t, w.proj_in, h0));
  const bool fold3 = w.ln3_fold;
  View n = (fold && fold3) ? View() : c.alloc_view(B, x.H, x.W, C);
  View a = c.alloc_view(B, x.H, x.W, C);
  // --- self attention ---
  if (!fold) c.layernorm(h0, w.ln1, n);
  View qkv = c.alloc_view(B, x.H, x.W, 3 * w.hs);
  consume(Ctx::args_heads((fold ? h0 : n).tokens(), w.qkv, qkv.tokens(), w.d, w.dstride));
  AttnArgs aa;
  aa.q = qkv.p; aa.k = qkv.p + w.hs; aa.v = qkv.p + 2 * w.hs;
  aa.ldq = aa.ldk = aa.ldv = 3 * w.hs;
  aa.B = B; aa.heads = kHeads; aa.Nq = HW; aa.Nk = HW; aa.d = w.d; aa.dstride = w.dstride;
  aa.out = a.p; aa.ldo = C;
  c.attention(aa);
  View h1 = c.alloc_view(B, x.H, x.W, C);
  View h0t = h0.tokens();
  produce(Ctx::args(a.tokens(), w.out1, h1.tokens(), 1, -1, &h0t));
  // --- cross attention ---
  if (!fold) c.layernorm(h1, w.ln2, n);
  View q2 = qkv;  // reuse
  q2.C = w.hs; q2.ld = w.hs;
  consume(Ctx::args_heads((fold ? h1 : n).tokens(), w.q2, q2.tokens(), w.d, w.dstride));
  aa.q = q2.p; aa.ldq = w.hs;
  aa.k = ctx_kv; aa.v = ctx_kv + w.hs; aa.ldk = aa.ldv = 2 * w.hs;
  aa.Nk = T;
  c.attention(aa);
  View h2 = h0;  // h0 is dead after out1's residual read
  View h1t = h1.tokens();
  if (fold3) produce(Ctx::args(a.tokens(), w.out2, h2.tokens(), 1, -1, &h1t));
  else c.conv(a.tokens(), w.out2, h2.tokens(), 1, -1, &h1t);
  // --- GEGLU feed-forward ---
  if (!fold3) c.layernorm(h2, w.ln3, n);
  View g = c.alloc_view(B, x.H, x.W, 4 * C);
  if (fold3) consume(Ctx::args(h2.tokens(), w.ff1, g.tokens(), 1, -1, nullptr, nullptr, 0, ACT_GEGLU));
  else c.conv(n.tokens(), w.ff1, g.tokens(), 1, -1, nullptr, nullptr, 0, ACT_GEGLU);
  View h3 = h1;
  View h2t = h2.tokens();
  c.conv(g.tokens(), w.ff2, h3.tokens(), 1, -1, &h2t);
  c.conv(h3, w.proj_out, out, 1, -1, &x);
  c.ws->release(m);
}

// time-embedding MLP + all per-ResBlock projections -> table [B][ncat] fp32 (diffusion_model.py:184-188,30,47)
inline float* time_table(Ctx& c, const TimeW& t, const float* t_emb, int B) {
  float* h1 = c.ws->alloc_n<float>((size_t)B * kTimeDim);
  float* h2 = c.ws->alloc_n<float>((size_t)B * kTimeDim);
  float* tab = c.ws->alloc_n<float>((size_t)B * t.ncat);
  c.skinny(t_emb, B, 320, t.w1, t.b1, kTimeDim, true, h1, kTimeDim);
  c.skinny(h1, B, kTimeDim, t.w2, t.b2, kTimeDim, true, h2, kTimeDim);
  c.skinny(h2, B, kTimeDim, t.wcat, t.bcat, t.ncat, false, tab, t.ncat);
  return tab;
}

struct CtxKV {  // per-layer K|V projections of the text context
  std::vector<bf16*> kv;
  int T = 0;
};

// context (B,T,768) bf16 -> K|V for the listed attention layers (loop-invariant: hoisted out of the step loop)
inline void project_context(Ctx& c, const bf16* ctx, int B, int T, const std::vector<const AttnW*>& layers, CtxKV& out) {
  out.kv.clear();
  out.T = T;
  View cv;
  cv.p = const_cast<bf16*>(ctx); cv.B = 1; cv.H = 1; cv.W = B * T; cv.C = kCtxDim; cv.ld = kCtxDim;
  for (const AttnW* w : layers) {
    View o;
    o.p = c.ws->alloc_n<bf16>((size_t)B * T * 2 * w->hs);
    o.B = 1; o.H = 1; o.W = B * T; o.C = 2 * w->hs; o.ld = 2 * w->hs;
    c.conv(cv, w->kv2, o);
    out.kv.push_back(o.p);
  }
}

inline std::vector<const AttnW*> encoder_attn_layers(const EncoderHalfW& e) {
  std::vector<const AttnW*> v;
  for (int i = 0; i < 6; ++i) v.push_back(&e.attn[i]);
  v.push_back(&e.mid_attn);
  return v;
}
inline std::vector<const AttnW*> unet_attn_layers(const UNetW& u) {
  std::vector<const AttnW*> v = encoder_attn_layers(u.enc);
  for (int i = 0; i < 9; ++i) v.push_back(&u.up_attn[i]);
  return v;
}

// Encoder half: conv_in output x0 already in outs[0]; fills outs[1..11] (views chosen by the caller) and x_mid.
inline void encoder_half(Ctx& c, const EncoderHalfW& e, View* outs /*[12]*/, const View& x_mid, const float* tab, int tab_ld,
                         const CtxKV& kv) {
  const int B = outs[0].B;
  int ai = 0;
  View x = outs[0];
  int oi = 1;
  for (int i = 0; i < 8; ++i) {
    const bool has_attn = i < 6;
    const size_t m = c.ws->mark();
    if (has_attn) {
      View r = c.alloc_view(B, x.H, x.W, e.res[i].cout);
      res_block(c, e.res[i], x, r, tab, tab_ld);
      attentions(c, e.attn[ai], r, outs[oi], kv.kv[ai], kv.T);
      ++ai;
    } else {
      res_block(c, e.res[i], x, outs[oi], tab, tab_ld);
    }
    c.ws->release(m);
    x = outs[oi++];
    if (i == 1 || i == 3 || i == 5) {  // stride-2 downsample conv, symmetric pad 1 (diffusion_model.py:200)
      c.conv(x, e.down[i / 2], outs[oi], 2, 1);
      x = outs[oi++];
    }
  }
  const size_t m = c.ws->mark();
  View a = c.alloc_view(B, x.H, x.W, 1280), b2 = c.alloc_view(B, x.H, x.W, 1280);
  res_block(c, e.mid0, x, a, tab, tab_ld);
  attentions(c, e.mid_attn, a, b2, kv.kv[6], kv.T);
  res_block(c, e.mid2, b2, x_mid, tab, tab_ld);
  c.ws->release(m);
}

// UNet forward in two halves, so that ControlNet's zero-convolutions can add their residuals to the skip tensors between
// them (diffusion_model.py:230-234) from their own epilogue instead of through 13 element-wise kernels.
struct UNetState {
  View cat[12];   // concat buffers [x | skip] of the up path
  View outs[12];  // the 12 skip tensors (channel slices of cat[])
  View xmid;      // middle-block output (slice of cat[0])
  const float* tab = nullptr;
  size_t mark = 0;
};

// latent8: (B,h,w,8) bf16 (4 real channels), t_emb (B,320) f32 device, kv: projected context.  time_tab: optional
// precomputed [B][ncat] table of the time-embedding MLP + per-ResBlock projections (the denoise loop computes it for every
// step before the loop: it depends on the timestep only), in which case t_emb is not read.
inline void unet_encode(Ctx& c, const UNetW& u, const bf16* latent8, int B, int h, int w, const float* t_emb, const CtxKV& kv,
                        const float* time_tab, UNetState& s) {
  s.mark = c.ws->mark();
  s.tab = time_tab ? time_tab : time_table(c, u.enc.time, t_emb, B);
  // concat buffers cat[i] = [x (cx) | skip (cs)] at the resolution of up block i
  static const int cx[12] = {1280, 1280, 1280, 1280, 1280, 1280, 1280, 640, 640, 640, 320, 320};
  static const int cs[12] = {1280, 1280, 1280, 1280, 1280, 640, 640, 640, 320, 320, 320, 320};
  static const int lvl[12] = {3, 3, 3, 2, 2, 2, 1, 1, 1, 0, 0, 0};
  for (int i = 0; i < 12; ++i) s.cat[i] = c.alloc_view(B, h >> lvl[i], w >> lvl[i], cx[i] + cs[i]);
  for (int k = 0; k < 12; ++k) s.outs[k] = s.cat[11 - k].slice(cx[11 - k], cs[11 - k]);
  View lat;
  lat.p = const_cast<bf16*>(latent8); lat.B = B; lat.H = h; lat.W = w; lat.C = 8; lat.ld = 8;
  c.conv(lat, u.enc.conv_in, s.outs[0]);
  s.xmid = s.cat[0].slice(0, 1280);
  encoder_half(c, u.enc, s.outs, s.xmid, s.tab, u.enc.time.ncat, kv);
}

// up path + output convolution; eps_out: (B,h,w,4) f32.  Releases everything unet_encode allocated.
inline void unet_decode(Ctx& c, const UNetW& u, int B, int h, int w, const CtxKV& kv, UNetState& s, float* eps_out) {
  static const int cx[12] = {1280, 1280, 1280, 1280, 1280, 1280, 1280, 640, 640, 640, 320, 320};
  const int tl = u.enc.time.ncat;
  const float* tab = s.tab;
  View* cat = s.cat;
  int ai = 0, ui = 0;
  View final_x = c.alloc_view(B, h, w, 320);
  for (int i = 0; i < 12; ++i) {
    const size_t m = c.ws->mark();
    const bool has_attn = i >= 3, up = (i == 2 || i == 5 || i == 8);
    View dst = (i == 11) ? final_x : cat[i + 1].slice(0, cx[i + 1]);  // where this stage's x goes
    const ResW& rw = u.up_res[i];
    View stage_out = up ? c.alloc_view(B, cat[i].H, cat[i].W, rw.cout) : dst;
    if (has_attn) {
      View r = c.alloc_view(B, cat[i].H, cat[i].W, rw.cout);
      res_block(c, rw, cat[i], r, tab, tl);
      attentions(c, u.up_attn[ai], r, stage_out, kv.kv[7 + ai], kv.T);
      ++ai;
    } else {
      res_block(c, rw, cat[i], stage_out, tab, tl);
    }
    if (up) c.upconv(stage_out, u.up_conv[ui++], dst);  // Upsamplers (diffusion_model.py:132-139), upsample folded into the conv
    c.ws->release(m);
  }
  View t = c.alloc_view(B, h, w, 320);
  c.groupnorm(final_x, u.out_norm, true, t);
  ConvArgs a;
  a.a0 = t; a.w = &u.conv_out; a.pad_t = a.pad_l = 1; a.outH = h; a.outW = w;
  a.out = eps_out; a.out_ld = 4; a.out_fp32 = true;
  c.conv(a);
  c.ws->release(s.mark);
}

// DiffusionModel.predict_on_batch (diffusion_model.py:163-283).  controls: null or 13 dense bf16 tensors.
inline void unet_forward(Ctx& c, const UNetW& u, const bf16* latent8, int B, int h, int w, const float* t_emb, const CtxKV& kv,
                         const bf16* const* controls, float* eps_out, const float* time_tab = nullptr) {
  UNetState s;
  unet_encode(c, u, latent8, B, h, w, t_emb, kv, time_tab, s);
  if (controls) {  // diffusion_model.py:230-234 with caller-supplied residual tensors
    c.add_inplace(s.xmid, controls[12]);
    for (int k = 0; k < 12; ++k) c.add_inplace(s.outs[k], controls[k]);
  }
  unet_decode(c, u, B, h, w, kv, s, eps_out);
}

// HintNet (control_net.py:10-31): image8 (B,H,W,8) bf16 (3 real channels, [0,1]) -> hint (B,H/8,W/8,320) bf16
inline void hintnet_forward(Ctx& c, const ControlNetW& cn, const bf16* image8, int B, int H, int W, const View& hint) {
  static const int stride[8] = {1, 1, 2, 1, 2, 1, 2, 1};
  const size_t m = c.ws->mark();
  View x;
  x.p = const_cast<bf16*>(image8); x.B = B; x.H = H; x.W = W; x.C = 8; x.ld = 8;
  for (int i = 0; i < 8; ++i) {
    const int oh = x.H / stride[i], ow = x.W / stride[i];
    View y = (i == 7) ? hint : c.alloc_view(B, oh, ow, cn.hint[i].N);
    c.conv(x, cn.hint[i], y, stride[i], 1, nullptr, nullptr, 0, i < 7 ? ACT_SILU : ACT_NONE);
    x = y;
  }
  c.ws->release(m);
}

// ControlNet (control_net.py:45-107): writes 13 dense bf16 residual tensors into res[i] — or, with `into` (the UNet's 12
// skip tensors and its middle-block output, from unet_encode), ADDS them there: the zero-convolution's epilogue reads its
// output tile as the residual and stores the sum in place (diffusion_model.py:230-234 without a separate add).
inline void controlnet_forward(Ctx& c, const ControlNetW& cn, const bf16* latent8, int B, int h, int w, const float* t_emb,
                               const CtxKV& kv, const View& hint, View* res /*[13]*/, const float* time_tab = nullptr,
                               const UNetState* into = nullptr) {
  const size_t m0 = c.ws->mark();
  const float* tab = time_tab ? time_tab : time_table(c, cn.enc.time, t_emb, B);
  const int tl = cn.enc.time.ncat;
  static const int ch[12] = {320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280};
  static const int lv[12] = {0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3};
  View outs[12];
  for (int k = 0; k < 12; ++k) outs[k] = c.alloc_view(B, h >> lv[k], w >> lv[k], ch[k]);
  View xmid = c.alloc_view(B, h >> 3, w >> 3, 1280);
  View lat;
  lat.p = const_cast<bf16*>(latent8); lat.B = B; lat.H = h; lat.W = w; lat.C = 8; lat.ld = 8;
  c.conv(lat, cn.enc.conv_in, outs[0], 1, 1, &hint);  // conv_in(latent) + hint (control_net.py:56)
  encoder_half(c, cn.enc, outs, xmid, tab, tl, kv);
  for (int k = 0; k < 13; ++k) {
    const View& src = k < 12 ? outs[k] : xmid;
    if (into) {
      const View& dst = k < 12 ? into->outs[k] : into->xmid;
      c.conv(src, cn.zero[k], dst, 1, -1, &dst);
    } else {
      c.conv(src, cn.zero[k], res[k]);
    }
  }
  c.ws->release(m0);
}

// VAE AttentionBlock (layers.py:28-59): GroupNorm, one q | k | v GEMM, the single-head d = 512 flash kernel (attn.cuh
// vattn_kernel: no N x N tensor), output projection + residual.
inline void vae_attention(Ctx& c, const VaeAttnW& w, const View& x, const View& out) {
  const size_t m = c.ws->mark();
  const int B = x.B, N = x.H * x.W, C = x.C;
  SDTF_CHECK(C == 512, "VAE attention block: 512 channels expected");
  View t = c.alloc_view(B, x.H, x.W, C);
  c.groupnorm(x, w.gn, false, t);
  View qkv = c.alloc_view(B, x.H, x.W, 3 * C);
  c.conv(t.tokens(), w.qkv, qkv.tokens());
  View o = t;  // the normalised input is dead once q | k | v exist
  c.vattn(qkv, N, w.v_bias, o);
  View xt = x.tokens();
  c.conv(o.tokens(), w.proj, out.tokens(), 1, -1, &xt);
  c.ws->release(m);
}

// VAE decoder (image_decoder.py:22-55): latent (B,h,w,4) f32 -> image (B,8h,8w,3) f32
inline void vae_decode(Ctx& c, const VaeDecW& d, const float* latent, int B, int h, int w, float* image_out) {
  const size_t m0 = c.ws->mark();
  View z = c.alloc_view(B, h, w, 8);
  c.cast_pad(latent, (long long)B * h * w, 4, 8, 1.0f / 0.18215f, z.p, false);  // Rescaling (image_decoder.py:27)
  View z2 = c.alloc_view(B, h, w, 8);
  c.memset0(z2.p, (size_t)B * h * w * 8 * 2);
  c.conv(z, d.post_quant, z2);
  View x = c.alloc_view(B, h, w, 512);
  c.conv(z2, d.conv_in, x);
  View y = c.alloc_view(B, h, w, 512);
  res_block(c, d.mid0, x, y, nullptr, 0);
  vae_attention(c, d.attn, y, x);
  res_block(c, d.mid1, x, y, nullptr, 0);
  x = y;
  int H = h, W = w;
  for (int b = 0; b < 4; ++b) {
    const int co = kVaeDecCout[b];
    View a = c.alloc_view(B, H, W, co), b2 = c.alloc_view(B, H, W, co);
    res_block(c, d.up[b][0], x, a, nullptr, 0);
    res_block(c, d.up[b][1], a, b2, nullptr, 0);
    res_block(c, d.up[b][2], b2, a, nullptr, 0);
    x = a;
    if (b < 3) {
      H *= 2; W *= 2;
      View n = c.alloc_view(B, H, W, co);
      c.upconv(x, d.up_conv[b], n);  // UpSampling2D(2) + PaddedConv2D(3) (image_decoder.py:36-47), upsample folded into the conv
      x = n;
    }
  }
  View t = c.alloc_view(B, H, W, 128);
  c.groupnorm(x, d.out_norm, true, t);
  ConvArgs a;
  a.a0 = t; a.w = &d.conv_out; a.pad_t = a.pad_l = 1; a.outH = H; a.outW = W;
  a.out = image_out; a.out_ld = 3; a.out_fp32 = true;
  c.conv(a);
  c.ws->release(m0);
}

// VAE encoder (image_encoder.py:21-48): image (B,H,W,3) f32 in [-1,1] -> latent mean * 0.18215 (B,H/8,W/8,4) f32
inline void vae_encode(Ctx& c, const VaeEncW& e, const float* image, int B, int H, int W, float* latent_out) {
  const size_t m0 = c.ws->mark();
  View img = c.alloc_view(B, H, W, 8);
  c.cast_pad(image, (long long)B * H * W, 3, 8, 1.f, img.p, false);
  View x = c.alloc_view(B, H, W, 128);
  c.conv(img, e.conv_in, x);
  for (int b = 0; b < 4; ++b) {
    const int co = kVaeEncCout[b];
    View a = c.alloc_view(B, H, W, co), b2 = c.alloc_view(B, H, W, co);
    res_block(c, e.down[b][0], x, a, nullptr, 0);
    res_block(c, e.down[b][1], a, b2, nullptr, 0);
    x = b2;
    if (b < 3) {  // stride 2, zero pad bottom/right only (image_encoder.py:31,34,37)
      H /= 2; W /= 2;
      View n = c.alloc_view(B, H, W, co);
      c.conv(x, e.down_conv[b], n, 2, 0);
      x = n;
    }
  }
  View y = c.alloc_view(B, H, W, 512);
  res_block(c, e.mid0, x, y, nullptr, 0);
  vae_attention(c, e.attn, y, x);
  res_block(c, e.mid1, x, y, nullptr, 0);
  View t = c.alloc_view(B, H, W, 512);
  c.groupnorm(y, e.out_norm, true, t);
  View m8 = c.alloc_view(B, H, W, 8);
  c.conv(t, e.conv_out, m8);
  ConvArgs a;
  a.a0 = m8; a.w = &e.quant; a.outH = H; a.outW = W; a.out = latent_out; a.out_ld = 4; a.out_fp32 = true;
  c.conv(a);
  c.ws->release(m0);
}

// ----------------------------------------------------------------------------------------------------------
// CLIP text tower (reference: text_encoder.py:22-33 embeddings, :36-55 encoder layer,
// :58-99 attention, :102-103 quick-GELU, :125-135 TextEncoder with clip_skip).  12 layers x (LN, q|k|v, causal
// attention, out + residual, LN, fc1 + quick-GELU, fc2 + residual), final LayerNorm of layer output `clip_skip`.
// The linears run on the same persistent tcgen05 GEMM as the UNet (M = B * 77 token rows).
// ----------------------------------------------------------------------------------------------------------
static constexpr int kClipLayers = 12, kClipHeads = 12, kClipTokens = 77;

struct TextLayerW {
  NormW ln1, ln2;
  PackedWeight qkv, out, fc1, fc2;
};
struct TextW {
  float* tok = nullptr;  // [vocab][768] fp32
  float* pos = nullptr;  // [77][768]
  int vocab = 0, max_len = 0;
  TextLayerW layer[kClipLayers];
  NormW final_ln;
  bool ready = false;
};

inline void build_text_encoder(WeightStore& ws, TextW& t) {
  const RawTensor* te = ws.find("text_model.embeddings.token_embedding.weight");
  const RawTensor* pe = ws.find("text_model.embeddings.position_embedding.weight");
  if (te && pe) {
    t.vocab = (int)te->shape[0];
    t.max_len = (int)pe->shape[0];
    t.tok = ws.pack_vec(*te, nullptr, 1.f);
    t.pos = ws.pack_vec(*pe, nullptr, 1.f);
  }
  const float qscale = 0.125f;  // head_dim^-1/2 = 64^-1/2, applied to q AFTER its bias in the reference (:84): folded into both
  for (int l = 0; l < kClipLayers; ++l) {
    const std::string p = "text_model.encoder.layers." + std::to_string(l);
    TextLayerW& L = t.layer[l];
    L.ln1 = ws.norm(p + ".layer_norm1");
    L.ln2 = ws.norm(p + ".layer_norm2");
    const RawTensor* w[3] = {ws.find(p + ".self_attn.q_proj.weight"), ws.find(p + ".self_attn.k_proj.weight"),
                             ws.find(p + ".self_attn.v_proj.weight")};
    const RawTensor* b[3] = {ws.find(p + ".self_attn.q_proj.bias"), ws.find(p + ".self_attn.k_proj.bias"),
                             ws.find(p + ".self_attn.v_proj.bias")};
    if (w[0] && w[1] && w[2] && b[0] && b[1] && b[2]) {
      const int C = (int)w[0]->shape[1];
      L.qkv.N = 3 * C; L.qkv.K = C; L.qkv.kh = L.qkv.kw = 1;
      L.qkv.w = (bf16*)ws.pool.alloc((size_t)3 * C * C * 2);
      L.qkv.bias = (float*)ws.pool.alloc((size_t)3 * C * 4);
      for (int i = 0; i < 3; ++i) {
        ws.pack_into(L.qkv.w, 3 * C, C, i * C, 0, *w[i], nullptr, i == 0 ? qscale : 1.f);
        ws.pack_vec(*b[i], nullptr, i == 0 ? qscale : 1.f, L.qkv.bias, i * C);
      }
    }
    L.out = ws.conv(p + ".self_attn.out_proj");
    L.fc1 = ws.conv(p + ".mlp.fc1");
    L.fc2 = ws.conv(p + ".mlp.fc2");
  }
  t.final_ln = ws.norm("text_model.final_layer_norm");
}

// TextClipEmbedding (text_encoder.py:22-33): x[b][t] = token_embedding[tokens[b][t]] + position_embedding[pos], fp32
inline void text_embed(Ctx& c, const TextW& t, const int* tokens, const int* positions, int pos_rows, int B, int T, float* x) {
  const long long rows = (long long)B * T;
  ++c.launches;
  if (c.dry) return;
  long long blocks = ceil_div_ll(rows * (kCtxDim / 4), 256);
  clip_embed_kernel<<<(unsigned)blocks, 256, 0, c.st>>>(tokens, positions, pos_rows, t.tok, t.pos, t.vocab, t.max_len, T, kCtxDim, rows, x);
  SDTF_CUDA(cudaGetLastError());
}

// TextEncoder (text_encoder.py:125-135) on x: [B][T][768] fp32 embeddings in the arena (overwritten: it is the residual
// stream); out: [B][T][768] fp32.  clip_skip = -1: all 12 layers, -2: 11, ... (reference indexing out[clip_skip], :133)
inline void text_encode(Ctx& c, const TextW& t, float* x, int B, int T, int clip_skip, float* out) {
  const int C = kCtxDim;
  const int n_layers = kClipLayers + clip_skip + 1;
  const size_t m0 = c.ws->mark();
  const long long rows = (long long)B * T;
  // residual stream x in fp32; GEMM operands (h, a, f) and GEMM outputs (qkv, d) in bf16
  View h = c.alloc_view(1, 1, (int)rows, C), a = c.alloc_view(1, 1, (int)rows, C), d = c.alloc_view(1, 1, (int)rows, C);
  View qkv = c.alloc_view(1, 1, (int)rows, 3 * C), f = c.alloc_view(1, 1, (int)rows, 4 * C);
  SDTF_CHECK(C == 768, "text tower: embed_dim must be 768");
  auto add_ln = [&](const bf16* delta, const NormW& n, bf16* out_bf16, float* out_f32) {
    ++c.launches;
    if (c.dry) return;
    clip_add_ln_kernel<6><<<(unsigned)ceil_div_ll(rows, 8), 256, 0, c.st>>>(x, delta, C, rows, n.gamma, n.beta, out_bf16, out_f32);
    SDTF_CUDA(cudaGetLastError());
  };
  const bf16* pending = nullptr;  // GEMM output not yet added to the stream
  for (int l = 0; l < n_layers; ++l) {
    const TextLayerW& L = t.layer[l];
    add_ln(pending, L.ln1, h.p, nullptr);
    c.conv(h, L.qkv, qkv);
    ++c.launches;
    if (!c.dry) {
      const size_t smem = ((size_t)2 * T * 65 + 4 * 128) * sizeof(float);
      clip_causal_attn_kernel<<<dim3(kClipHeads, B), 128, smem, c.st>>>(qkv.p, T, C, a.p);
      SDTF_CUDA(cudaGetLastError());
    }
    c.conv(a, L.out, d);
    add_ln(d.p, L.ln2, h.p, nullptr);
    c.conv(h, L.fc1, f, 1, -1, nullptr, nullptr, 0, ACT_QGELU);
    c.conv(f, L.fc2, d);
    pending = d.p;
  }
  add_ln(pending, t.final_ln, nullptr, out);
  c.ws->release(m0);
}

}  // namespace sdtf
