/* sdtf.h — C ABI of the B200-native SD1.5 denoising engine (libsdtf.so).
 *
 * The reference (cpuimage/minSDTF) has no FFI: its only seam is Python duck typing — the pipeline touches its
 * models exclusively through `.predict_on_batch(list_of_ndarrays)` (stable_diffusion/stable_diffusion.py:415,
 * 439,447-457,482) and the scheduler through `set_timesteps / timesteps / signal_rates / noise_rates / step`
 * (:399-400,468,566-567).  Each entry point below is what a ctypes binding for one of those calls binds to;
 * the reference call it replaces is cited per function.  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - Tensors cross as `const DLManagedTensor*` (DLPack v0.x struct, extracted from a "dltensor" PyCapsule).
 *     They are BORROWED for the duration of the call: the engine never calls `deleter`.  Host (kDLCPU) and
 *     device (kDLCUDA) tensors are both accepted; host tensors are copied by the engine (H2D / D2H on the
 *     engine's stream), device tensors must live on the engine's GPU.  Must be compact row-major.
 *   - Activations are NHWC, exactly as the reference's Keras graphs take them.  Weights are in the reference's
 *     checkpoint convention: PyTorch layout (conv OIHW, linear (out,in)) under the reference's key names
 *     (ckpt_loader.py:708-2133 CKPT_MAPPING).
 *   - Every function returns 0 on success, a negative code on failure; the message is in sdtf_last_error().
 *     No C++ exception crosses this boundary.  One engine per GPU; calls on one engine must be serialised.
 *   - There is no CPU fallback: sdtf_create fails if the device is not compute capability 10.x.
 */
#ifndef SDTF_H_
#define SDTF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- minimal DLPack (v0.8 ABI) ------------------------------------------------------------------------ */
#ifndef DLPACK_DLPACK_H_
typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3 } DLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef enum { kDLInt = 0, kDLUInt = 1, kDLFloat = 2, kDLBfloat = 4 } DLDataTypeCode;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;
typedef struct {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides; /* in elements; NULL = compact row-major */
  uint64_t byte_offset;
} DLTensor;
typedef struct DLManagedTensor {
  DLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;
#endif

typedef struct sdtf_engine sdtf_engine;

enum {
  SDTF_OK = 0,
  SDTF_ERR_INVALID = -1,  /* bad argument / shape / dtype / missing weight */
  SDTF_ERR_CUDA = -2,     /* CUDA runtime or driver error */
  SDTF_ERR_DEVICE = -3,   /* no sm_100 device: there is no fallback path */
  SDTF_ERR_INTERNAL = -4
};

/* Per-step scalars of the fused CFG + scheduler kernel, computed on the host in fp64 from the reference's
 * schedule (scheduler.py:52-55, 285-312) and passed as fp32:
 *   eps     = guidance > 0 ? eps_u + guidance * (eps_c - eps_u) : eps_c          (stable_diffusion.py:458)
 *   eps    *= rescale * std(eps_c)/(std(eps)+1e-5) + (1 - rescale)   if rescale > 0   (:304-315)
 *   latent' = ca * latent + cb * eps + cn * noise                                 (scheduler.py:285-312)
 *   latent' = (sig_t*init + noi_t*init_noise) * (1-mask) + latent' * mask   if mask (stable_diffusion.py:469-475) */
typedef struct {
  float guidance, rescale;
  float ca, cb, cn;
  float sig_t, noi_t;
} sdtf_step_coef;

/* Whole denoising job = the loop of StableDiffusionBase.generate_image (stable_diffusion.py:442-486). */
typedef struct {
  int32_t n_steps;
  int32_t use_cuda_graph;                  /* 1: capture one step, replay it n_steps times */
  int32_t decode;                          /* 1: run the VAE decoder and produce uint8 images */
  int32_t cfg_split;                       /* 1: 2-way CFG split over the engine's communicator (sdtf_comm_init): this
                                              rank evaluates one branch (rank 0 uncond, rank 1 cond) and the epsilons
                                              are exchanged with one ncclAllGather per step */
  const DLManagedTensor* latent0;          /* (B,h,w,4) f32 start latent (noise, or noised init for img2img) */
  const DLManagedTensor* context;          /* (B,T,768) f32 */
  const DLManagedTensor* uncond_context;   /* (B,Tu,768) f32, NULL when guidance == 0.  Tu may differ from T (a long
                                              prompt against a short negative prompt): the two branches are then
                                              evaluated as two passes, like the reference's two model calls */
  const DLManagedTensor* t_emb;            /* (n_steps,320) f32 sinusoidal embeddings in execution order */
  const sdtf_step_coef* coefs;             /* [n_steps] host array, execution order */
  const DLManagedTensor* step_noise;       /* (n_steps,B,h,w,4) f32 TCD noise, or NULL */
  const DLManagedTensor* mask;             /* (h,w) f32 latent mask for inpaint, or NULL */
  const DLManagedTensor* init_latent;      /* (h,w,4) f32 encoded source image (inpaint), or NULL */
  const DLManagedTensor* init_noise;       /* (B,h,w,4) f32 noise used to re-noise init_latent, or NULL */
  const DLManagedTensor* hint_image;       /* (B,H,W,3) f32 in [0,1] ControlNet image, or NULL */
  const DLManagedTensor* blend_image;      /* (H,W,3) f32 in [0,1] source image for the final pixel blend, or NULL */
  const DLManagedTensor* blend_mask;       /* (H,W) f32 */
  DLManagedTensor* out_images;             /* (B,H,W,3) u8 (decode=1) */
  DLManagedTensor* out_latent;             /* (B,h,w,4) f32 final latent, or NULL */
  /* `callback(iteration)` of generate_image (stable_diffusion.py:476-478), or NULL: called on the calling thread
   * after every denoising step has completed on the device, iteration = 1..n_steps */
  void (*on_step)(int32_t iteration, void* user);
  void* on_step_user;
} sdtf_denoise_desc;

typedef struct {
  float loop_ms;     /* device time of the n_steps denoising steps of the last sdtf_denoise */
  float decode_ms;   /* device time of the VAE decode */
  float total_ms;    /* whole call incl. copies */
  int32_t kernel_launches; /* kernels launched (or replayed inside graphs) by the last call */
} sdtf_timings;

/* lifecycle ------------------------------------------------------------------------------------------ */
int sdtf_create(int32_t device, sdtf_engine** out);
void sdtf_destroy(sdtf_engine* e);
const char* sdtf_last_error(const sdtf_engine* e); /* e may be NULL: error of the last failed sdtf_create */
const char* sdtf_version(void);

/* weights: replaces load_weights_from_file (ckpt_loader.py:2136-2193).  `key` is a reference checkpoint key
 * (LDM `model.diffusion_model.*`, `control_model.*`, diffusers-legacy VAE names); tensor f32 / f16 / bf16 in
 * PyTorch layout.  sdtf_finalize_weights packs a component ("unet", "controlnet" (incl. hint block),
 * "vae_decoder", "vae_encoder") into bf16 device arenas and fails listing any missing key. */
int sdtf_load_tensor(sdtf_engine* e, const char* key, const DLManagedTensor* t);
int sdtf_finalize_weights(sdtf_engine* e, const char* component);

/* DiffusionModel.predict_on_batch([latent, t_emb, context] (+13 controls)) — diffusion_model.py:163-283.
 * latent (B,h,w,4) f32, t_emb (B,320) f32, context (B,T,768) f32, controls: NULL or 13 tensors (f32 NHWC),
 * out_eps (B,h,w,4) f32. */
int sdtf_unet_forward(sdtf_engine* e, const DLManagedTensor* latent, const DLManagedTensor* t_emb,
                      const DLManagedTensor* context, const DLManagedTensor* const* controls, DLManagedTensor* out_eps);

/* ControlNet.predict_on_batch([latent, t_emb, context, hint]) -> 13 residuals — control_net.py:45-107 */
int sdtf_controlnet_forward(sdtf_engine* e, const DLManagedTensor* latent, const DLManagedTensor* t_emb,
                            const DLManagedTensor* context, const DLManagedTensor* hint, DLManagedTensor* const* outs);

/* HintNet.predict_on_batch(image (B,H,W,3) in [0,1]) -> (B,H/8,W/8,320) — control_net.py:10-31 */
int sdtf_hintnet_forward(sdtf_engine* e, const DLManagedTensor* image, DLManagedTensor* out);

/* ImageDecoder.predict_on_batch(latent) -> (B,8h,8w,3) f32 — image_decoder.py:22-55 */
int sdtf_vae_decode(sdtf_engine* e, const DLManagedTensor* latent, DLManagedTensor* out_image);

/* ImageEncoder.predict_on_batch(image in [-1,1]) -> (B,H/8,W/8,4) f32 — image_encoder.py:21-48 */
int sdtf_vae_encode(sdtf_engine* e, const DLManagedTensor* image, DLManagedTensor* out_latent);

/* TextEncoder.predict_on_batch(TextClipEmbedding.predict_on_batch([tokens, positions])) — text_encoder.py:106-135:
 * tokens (B,T<=77) int32 -> context (B,T,768) f32; clip_skip in [-12,-1] selects the encoder layer whose output is
 * normalised (text_encoder.py:133).  Component "text_encoder", keys text_model.* (text_encoder.py:110-111,137-157). */
int sdtf_text_encode(sdtf_engine* e, const DLManagedTensor* tokens, int clip_skip, DLManagedTensor* out_context);

/* The reference's two text models separately, for textual inversion, which splices learned vectors between them
 * (long_prompt_weighting.py:202-209, 231-235):
 *   sdtf_text_embed            TextClipEmbedding.predict_on_batch([tokens, positions]) — text_encoder.py:22-33,106-122:
 *                              tokens (B,T) int32, positions (1,T) or (B,T) int32 or NULL (0..T-1) -> (B,T,768) f32
 *   sdtf_text_encode_embedded  TextEncoder.predict_on_batch(clip_embedding) — text_encoder.py:125-135:
 *                              embedding (B,T,768) f32 -> context (B,T,768) f32 */
int sdtf_text_embed(sdtf_engine* e, const DLManagedTensor* tokens, const DLManagedTensor* positions,
                    DLManagedTensor* out_embedding);
int sdtf_text_encode_embedded(sdtf_engine* e, const DLManagedTensor* embedding, int clip_skip, DLManagedTensor* out_context);

/* CFG combine + rescale + Scheduler.step (+ inpaint blend) as ONE kernel — stable_diffusion.py:458-475,
 * scheduler.py:246-315.  eps_u may be NULL (no guidance).  All (B,h,w,4) f32 except mask (h,w), init_latent (h,w,4). */
int sdtf_cfg_sched_step(sdtf_engine* e, const DLManagedTensor* eps_u, const DLManagedTensor* eps_c,
                        const DLManagedTensor* latent_prev, const sdtf_step_coef* coef, const DLManagedTensor* noise,
                        const DLManagedTensor* mask, const DLManagedTensor* init_latent,
                        const DLManagedTensor* init_noise, DLManagedTensor* out_latent);

/* decoded f32 (B,H,W,3) -> uint8 with the reference's arithmetic (stable_diffusion.py:483-486) */
int sdtf_to_uint8(sdtf_engine* e, const DLManagedTensor* decoded, const DLManagedTensor* blend_image,
                  const DLManagedTensor* blend_mask, DLManagedTensor* out_u8);

/* the whole generate_image loop on the device */
int sdtf_denoise(sdtf_engine* e, const sdtf_denoise_desc* d);
int sdtf_get_timings(const sdtf_engine* e, sdtf_timings* out);

/* 2-way CFG split (the reference has no multi-device path).  libnccl is dlopen'ed from `nccl_lib`
 * (NULL: $SDTF_NCCL_LIB, then the default soname).  sdtf_comm_unique_id fills 128 bytes on one rank; the caller ships
 * them to its partner (any transport) and both call sdtf_comm_init(engine, lib, id, rank, 2). */
int sdtf_comm_unique_id(const char* nccl_lib, void* out_id128);
int sdtf_comm_init(sdtf_engine* e, const char* nccl_lib, const void* id128, int32_t rank, int32_t world);
int sdtf_comm_destroy(sdtf_engine* e);

/* Operator-class accounting for bench.py's roofline line: between sdtf_trace_begin and sdtf_trace_end every conv /
 * linear, attention, GroupNorm and LayerNorm launch is issued eagerly (no CUDA graph) and bracketed by CUDA events on
 * the engine's stream; sdtf_trace_end returns, per class, the launch count, the summed device time and the summed
 * ALGORITHMIC work (FLOP = 2 MAC of the contraction; bytes = minimal reads + writes) of those launches. */
typedef struct {
  int64_t launches[4];  /* conv/linear GEMM, attention, GroupNorm, LayerNorm */
  double us[4], flop[4], bytes[4];
} sdtf_trace_summary;
int sdtf_trace_begin(sdtf_engine* e);
int sdtf_trace_end(sdtf_engine* e, sdtf_trace_summary* out);

/* kernel micro-benchmarks used by bench.py for the roofline line: runs `reps` launches of one representative
 * contraction of the UNet (3x3 conv, batch*hw x cin -> cout) on device-resident synthetic data and returns the
 * average CUDA-event time per launch in ms. */
int sdtf_bench_conv(sdtf_engine* e, int32_t batch, int32_t hw, int32_t cin, int32_t cout, int32_t ksize,
                    int32_t reps, float* ms_per_launch);

/* same for the attention kernel: q (batch, nq, heads*d), k/v (batch, nk, heads*d) synthetic; legacy != 0 selects the
 * one-query-tile-per-CTA kernel (kept for d > 64 and for A/B measurements). */
int sdtf_bench_attention(sdtf_engine* e, int32_t batch, int32_t heads, int32_t nq, int32_t nk, int32_t d, int32_t reps,
                         int32_t legacy, float* ms_per_launch);

/* test hooks (used by tests/ only): one kernel each behind the same marshalling.
 * attention: q (B,Nq,heads*d), k/v (B,Nk,heads*d) f32 -> softmax(q k^T d^-1/2) v per head (diffusion_model.py:118-128);
 *   heads < 0 runs |heads| heads through the legacy one-tile-per-CTA kernel
 * norm: x (B,H,W,C) f32; mode 0 GroupNorm(32), 1 GroupNorm+SiLU, 2 LayerNorm(C); eps 1e-5 */
int sdtf_test_attention(sdtf_engine* e, const DLManagedTensor* q, const DLManagedTensor* k, const DLManagedTensor* v,
                        int32_t heads, DLManagedTensor* out);
int sdtf_test_norm(sdtf_engine* e, const DLManagedTensor* x, const DLManagedTensor* gamma, const DLManagedTensor* beta,
                   int32_t mode, DLManagedTensor* out);

#ifdef __cplusplus
}
#endif
#endif /* SDTF_H_ */
