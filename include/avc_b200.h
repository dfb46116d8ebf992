/* avc_b200.h -- C ABI of libavc_b200.so: B200-native (sm_100a) kernels for the AvatarCLIP
 * appearance-optimisation hot path.
 *
 * The reference (hongfz16/AvatarCLIP, AvatarGen/AppearanceGen) is pure Python/PyTorch and has
 * no FFI layer of its own; the seam it offers is the Python object protocol used by
 * `Runner` (main.py:147-151, 418-420).  Each entry point below replaces the torch-eager
 * implementation of one reference function; the citation names it (paths relative to
 * AvatarGen/AppearanceGen).  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (a torch tensor), fp32 unless noted;
 *   - every function enqueues work on `stream` and returns immediately (no host sync);
 *   - nothing allocates: scratch is sized by the *_workspace_bytes queries and passed in;
 *   - return value: 0 ok, <0 AVC_E_* (invalid argument), >0 a cudaError_t;
 *   - no results or device state are kept between calls outside the caller's buffers.  Host-side conveniences are
 *     thread-local (a cache of encoded TMA tensor maps, the SM count, "attribute already set" flags): safe to call
 *     from one host thread per device (the supported deployment is one process per GPU, torchrun style).  Tuning
 *     knobs are environment variables (AVC_*) read on the host.
 */
#ifndef AVC_B200_H
#define AVC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* avc_stream_t; /* cudaStream_t */

#define AVC_OK 0
#define AVC_E_BADCFG (-1)   /* unsupported network / renderer configuration            */
#define AVC_E_NULL (-2)     /* a required pointer is NULL                               */
#define AVC_E_SIZE (-3)     /* workspace too small / size mismatch                      */
#define AVC_E_ALIGN (-4)    /* pointer not 16-byte aligned                              */
#define AVC_E_NOSTASH (-5)  /* backward called on a workspace with no matching forward  */

#define AVC_ABI_VERSION 3
int avc_abi_version(void);
/* Compiled-for architecture string, e.g. "sm_100a". */
const char* avc_build_arch(void);

/* ------------------------------------------------------------------------------------------
 * NeuS renderer: SDFNetwork + RenderingNetwork + SingleVarianceNetwork + NeuSRenderer.render
 * (models/fields.py:9-107,111-185,270-276; models/embedder.py:6-51; models/renderer.py:39-69,
 * 133-397).  Supported: mode 'no_view_dir', multires_view 0, squeeze_out, weight_norm, n_outside 0 -- i.e.
 * every conf shipped under confs/ (SURVEY.md fact 1).  The kernels always evaluate the extra colour head
 * (extra_color = True, 179 confs); for the one conf without it (base_models/astrongman.conf, --mode train) the host
 * side fills the head's parameter slot with the constant zero map and blends the background into color_fine itself
 * (models/renderer.py:272-281; avatarclip_b200/renderer.py).
 * ------------------------------------------------------------------------------------------ */
typedef struct avc_neus_cfg {
  /* SDFNetwork ctor (models/fields.py:10-21) */
  int32_t sdf_d_in;        /* 3 */
  int32_t sdf_d_out;       /* d_hidden + 1 in every conf: sdf + feature vector */
  int32_t sdf_d_hidden;
  int32_t sdf_n_layers;    /* number of hidden layers; linears = n_layers + 1 */
  uint32_t sdf_skip_mask;  /* bit l set <=> l in skip_in */
  int32_t sdf_multires;
  float sdf_scale;
  /* RenderingNetwork ctor (models/fields.py:112-122); d_in = 6 (points, normals) */
  int32_t col_d_feature;
  int32_t col_d_hidden;
  int32_t col_n_layers;
  /* NeuSRenderer ctor (models/renderer.py:73-93) */
  int32_t n_samples;
  int32_t n_importance;
  int32_t up_sample_steps;
  /* arithmetic engine for the MLP contractions: 0 = fp32 CUDA-core (FFMA) tiles,
   * 1 = tcgen05 tensor-core tiles with two-term split operands (3 MMAs per product) */
  int32_t engine;
  /* tcgen05 engine only: MMAs per product in the COLOUR net (forward, dgrad, wgrad).  3 (or 0) = the same two-term split
   * as the SDF trunk; 1 = single-pass bf16 on the hi halves (the colour net tolerates it in the
   * rendered RGB; the flat parameter gradient moves from ~4e-5 to 3e-4 .. 1e-3 rel-L2). */
  int32_t color_products;
  /* tcgen05 engine only: MMAs per product in the WEIGHT-GRADIENT tiles (dW += zbar^T in, qt^T ubar, ... : sums over all
   * sample points of a chunk).  3 (or 0) = two-term split operands; 1 = single-pass bf16 on the hi halves. */
  int32_t wgrad_products;
} avc_neus_cfg;

/* Flat parameter vector layout (fp32), identical for the gradient vector:
 *   for each SDF linear l = 0..n_layers:    weight_g[out], weight_v[out*in], bias[out]
 *   for each colour linear l = 0..n_layers: weight_g, weight_v, bias
 *   extra_lin:                              weight_g[3], weight_v[3*d_hidden], bias[3]
 *   variance[1]
 * (the reference's state-dict tensors `linK.weight_g/weight_v/bias`, `extra_lin.*`, `variance`,
 * in named_parameters() order).  avc_neus_param_count returns the total length. */
int avc_neus_param_count(const avc_neus_cfg* cfg, int64_t* n_params);
/* Offset (in floats) of tensor `which` (0 = weight_g, 1 = weight_v, 2 = bias) of linear `layer`
 * of net `net` (0 = sdf, 1 = colour, 2 = extra_lin (layer ignored), 3 = variance). */
int avc_neus_param_offset(const avc_neus_cfg* cfg, int net, int layer, int which, int64_t* offset,
                          int64_t* numel);

/* Bytes of scratch for rendering up to `max_rays_per_chunk` rays per internal chunk. */
int avc_neus_workspace_bytes(const avc_neus_cfg* cfg, int64_t max_rays_per_chunk, size_t* bytes);

typedef struct avc_neus_outputs { /* the dict of models/renderer.py:385-397, all [R, ...] row-major */
  float* color_fine;       /* [R,3]   */
  float* extra_color_fine; /* [R,3]   */
  float* s_val;            /* [R,1]   */
  float* cdf_fine;         /* [R,S]   */
  float* weight_sum;       /* [R,1]   */
  float* weight_max;       /* [R,1]   */
  float* gradients;        /* [R,S,3] */
  float* weights;          /* [R,S]   */
  float* mid_z_vals;       /* [R,S]   */
  float* gradient_error;   /* [1]     */
  float* inside_sphere;    /* [R,S]   */
  float* z_vals;           /* [R,S]  sorted sample depths (saved for the backward) */
} avc_neus_outputs;

/* NeuSRenderer.render forward (models/renderer.py:302-397).
 *   params      flat parameter vector (see above)
 *   rays_o/d    [R,3];  near/far [R]
 *   jitter      [R] values (u-0.5) of renderer.py:317-319, or NULL for perturb = 0
 *   background  NULL, or [3] (bg_kind 1: one colour, main.py:393), or [R] (bg_kind 2: per-ray grey,
 *               main.py:395-405)
 *   z_vals_in   NULL, or [R,S] sorted depths to composite on (skips the placement passes)
 * S = n_samples + n_importance.  The workspace keeps what the backward needs when the call
 * fits one chunk; otherwise the backward recomputes per chunk from out->z_vals. */
int avc_neus_render_fwd(const avc_neus_cfg* cfg, const float* params, const float* rays_o,
                        const float* rays_d, const float* near, const float* far,
                        const float* jitter, const float* background, int bg_kind,
                        const float* z_vals_in, float cos_anneal_ratio, int64_t R,
                        const avc_neus_outputs* out, void* workspace, size_t workspace_bytes,
                        int64_t max_rays_per_chunk, avc_stream_t stream);

typedef struct avc_neus_cotangents { /* d loss / d output; NULL = zero */
  const float* color_fine;       /* [R,3]   */
  const float* extra_color_fine; /* [R,3]   */
  const float* s_val;            /* [R,1]   */
  const float* cdf_fine;         /* [R,S]   */
  const float* weight_sum;       /* [R,1]   */
  const float* weight_max;       /* [R,1]   */
  const float* gradients;        /* [R,S,3] */
  const float* weights;          /* [R,S]   */
  const float* gradient_error;   /* [1]     */
} avc_neus_cotangents;

/* Backward of avc_neus_render_fwd w.r.t. every parameter (the autograd graph the reference builds
 * at models/fields.py:96-107 + main.py:537, second-order terms included).  Overwrites
 * grad_params[n_params].  z_vals / weights etc. are the forward's outputs. */
int avc_neus_render_bwd(const avc_neus_cfg* cfg, const float* params, const float* rays_o,
                        const float* rays_d, const float* background, int bg_kind,
                        float cos_anneal_ratio, int64_t R, const avc_neus_outputs* fwd_out,
                        const avc_neus_cotangents* cot, float* grad_params, void* workspace,
                        size_t workspace_bytes, int64_t max_rays_per_chunk, int32_t flags,
                        avc_stream_t stream);
/* flags for avc_neus_render_bwd: rebuild the forward stash from fwd_out->z_vals even for a
 * single-chunk call (use when the workspace was reused by another call since the forward; the
 * eikonal normaliser is then taken from fwd_out). */
#define AVC_BWD_RECOMPUTE 1

/* SDFNetwork.sdf on arbitrary points (models/fields.py:90-91; used by extract_fields,
 * renderer.py:10-25): sdf_out[P]. */
int avc_neus_sdf_query(const avc_neus_cfg* cfg, const float* params, const float* pts, int64_t P,
                       float* sdf_out, void* workspace, size_t workspace_bytes,
                       avc_stream_t stream);


/* ------------------------------------------------------------------------------------------
 * CLIP ViT-B/32 image tower + cosine loss (openai/CLIP `VisionTransformer`, un-vendored
 * third-party dependency of the reference; call sites main.py:259-261 (load, frozen),
 * :509-526 (resize -> normalise -> encode_image -> cosine against the cached text embedding)).
 * Weights are frozen (main.py:260), so the backward is input-gradient only.
 * GEMM operands are fp16 (clip.load keeps fp16 weights on CUDA), accumulation, the residual
 * stream, LayerNorm and softmax are fp32.
 * ------------------------------------------------------------------------------------------ */
typedef struct avc_clip_cfg {
  int32_t image_size; /* 224 */
  int32_t patch;      /* 32  */
  int32_t width;      /* 768 */
  int32_t layers;     /* 12  */
  int32_t heads;      /* 12  */
  int32_t mlp;        /* 3072 */
  int32_t out_dim;    /* 512 */
} avc_clip_cfg;

typedef struct avc_clip_layer_weights { /* transformer.resblocks.{i}.* ; *_t = transposed copy */
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  const void *w_qkv, *w_qkv_t;   /* fp16 [3W,W], [W,3W]  attn.in_proj_weight  */
  const float* b_qkv;            /* [3W]                 attn.in_proj_bias    */
  const void *w_out, *w_out_t;   /* fp16 [W,W]           attn.out_proj.weight */
  const float* b_out;
  const void *w_fc, *w_fc_t;     /* fp16 [mlp,W], [W,mlp]  mlp.c_fc.weight    */
  const float* b_fc;
  const void *w_proj, *w_proj_t; /* fp16 [W,mlp], [mlp,W]  mlp.c_proj.weight  */
  const float* b_proj;
} avc_clip_layer_weights;

#define AVC_CLIP_MAX_LAYERS 24
typedef struct avc_clip_weights {
  const void *w_patch, *w_patch_t; /* fp16 [W, 3*patch*patch] (conv1.weight flattened), [3*p*p, W] */
  const float *cls, *pos;          /* class_embedding [W], positional_embedding [T,W] */
  const float *ln_pre_g, *ln_pre_b, *ln_post_g, *ln_post_b;
  const float* proj;               /* fp32 [W, out_dim] */
  avc_clip_layer_weights layer[AVC_CLIP_MAX_LAYERS];
} avc_clip_weights;

int avc_clip_workspace_bytes(const avc_clip_cfg* cfg, int32_t B, size_t* bytes);

/* B canvases [B][H][W][3] (fp32, values in [0,1]; the reshape of main.py:510) -> whole-image bilinear
 * resize to image_size^2 (align_corners=False, no antialias) -> Normalize(mean,std) (main.py:261)
 * -> encode_image -> emb_out[B][out_dim]; cos_out[b] = cosine(emb_out[b], text_emb[b]) (main.py:513). */
int avc_clip_loss_fwd(const avc_clip_cfg* cfg, const avc_clip_weights* w, const float* canvases,
                      int32_t H, int32_t W, int32_t B, int32_t input_mode, const float* text_emb,
                      float* emb_out, float* cos_out, void* workspace, size_t workspace_bytes,
                      avc_stream_t stream);
/* input_mode 0: canvases as above.  input_mode 1: `canvases` is an already resized + normalised NCHW
 * image batch [B][3][image_size][image_size] (the argument of perceptor.encode_image, main.py:512);
 * H = W = image_size.
 * Backward: d loss / d input given g_cos[b] = d loss / d cos_out[b] and/or g_emb[b][out_dim] =
 * d loss / d emb_out (either may be NULL); uses what the forward left in the workspace.  Overwrites
 * d_canvases (same shape as the forward input). */
int avc_clip_loss_bwd(const avc_clip_cfg* cfg, const avc_clip_weights* w, int32_t H, int32_t W, int32_t B,
                      int32_t input_mode, const float* text_emb, const float* g_cos, const float* g_emb,
                      float* d_canvases, void* workspace, size_t workspace_bytes, avc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Shading + canvas scatter + non-CLIP losses of Runner.train_clip (main.py:417-497, 528-534) with
 * use_silhouettes = True (every shipped train_clip conf).  The two switches the shipped confs vary
 * (the 18 confs/ablation files ending in _0 / _1 / _2) are the last two fields: zero-initialised = add_no_texture = texture_cast_light =
 * True, the configuration of confs/examples*.
 * Rays are the True pixels of the dilated mask; pix[r] is the flat canvas index (y*W + x) of ray r.
 * ------------------------------------------------------------------------------------------ */
typedef struct avc_loss_inputs {
  /* render outputs (NeuSRenderer.render dict) */
  const float* color_fine;       /* [R,3] */
  const float* extra_color_fine; /* [R,3] */
  const float* gradients;        /* [R,S,3] */
  const float* weights;          /* [R,S] */
  const float* weight_sum;       /* [R] */
  const float* gradient_error;   /* [1] */
  /* per step */
  const int32_t* pix;            /* [R] canvas index of each ray */
  const uint8_t* in_mask;        /* [H*W] 1 where a ray exists (the dilated mask, dataset.py:255-256) */
  const float* true_rgb;         /* [H*W,3] template render resized to the canvas (main.py:376) */
  const float* mask;             /* [H*W] 0/1 (main.py:377-380, 407-410) */
  const float* background;       /* NULL, or [H*W] grey levels for bg_choice 1/2 (main.py:394-405) */
  int32_t bg_choice;             /* 0 white, 1/2 per-pixel grey, 3 black (main.py:387-415) */
  float light_dir[3];            /* sphere_coord(theta+U, phi+U) of main.py:433 (un-normalised) */
  float ambience;                /* main.py:440 */
  const float* view_scalars;     /* NULL, or DEVICE [4] = {light_dir[3], ambience}: overrides the two host fields above
                                    (lets a captured CUDA graph of the step be replayed with new per-view draws) */
  float igr_weight, mask_weight, clip_weight;  /* conf train.* */
  int32_t R, S, H, W;
  /* ABI 3 */
  int32_t plain_texture;         /* 1: train.texture_cast_light = False -- canvas 0 is the extra colour itself
                                    (full_extra_color_fine, main.py:475-477,516), no shading factor, no clamp */
  int32_t no_shading_term;       /* 1: train.add_no_texture = False -- the loss has no CLIP term on canvas 1
                                    (main.py:521,533): the backward ignores d_canvases[1] */
} avc_loss_inputs;

/* scalars written by the forward: [0] color_loss, [1] eikonal, [2] mask_loss (BCE), [3] psnr,
 * [4] base_loss = color + igr*eik + mask_w*bce, [5..7] internal sums (l1, bce, sq), [8] mask_sum */
#define AVC_LOSS_SCALARS 16
/* Forward: fills canvases[2][H][W][3] (0: texture_shading -- or the extra colour when plain_texture --,
 * 1: rand_shading_rgb; main.py:466-477) and scalars[AVC_LOSS_SCALARS].  Canvas 1 is always written (finite values)
 * so that a caller may keep one B = 2 CLIP batch for every configuration. */
int avc_loss_stage_fwd(const avc_loss_inputs* in, float* canvases, float* scalars, avc_stream_t stream);
/* Backward: d_canvases[2][H][W][3] = d loss / d canvases (from the CLIP backward) plus the direct
 * loss terms -> cotangents of the render outputs (struct avc_neus_cotangents, written in full:
 * color_fine, extra_color_fine, gradients, weights, weight_sum, gradient_error must be non-NULL
 * writable buffers; s_val, cdf_fine, weight_max are ignored). */
int avc_loss_stage_bwd(const avc_loss_inputs* in, const float* d_canvases, const float* scalars,
                       const avc_neus_cotangents* cot_out, avc_stream_t stream);

/* SDFNetwork.forward / .sdf_hidden_appearance / .gradient (models/fields.py:72-107) on P arbitrary points (boundary
 * convenience, exact-fp32 tiles regardless of cfg->engine): sdf_feat_out [P][d_out] = (sdf, feature vector) or NULL;
 * grad_out [P][3] = d sdf / d x (raw, un-normalised) or NULL.  Workspace: as avc_neus_sdf_query. */
int avc_neus_sdf_eval(const avc_neus_cfg* cfg, const float* params, const float* pts, int64_t P, float* sdf_feat_out,
                      float* grad_out, void* workspace, size_t workspace_bytes, avc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused Adam over the flat parameter vector (torch.optim.Adam defaults, main.py:145,536-538):
 * p -= lr * mhat / (sqrt(vhat) + eps); `step` is the 1-based step count; grad_scale multiplies g
 * first (1/world_size after the gradient all-reduce).
 * ------------------------------------------------------------------------------------------ */
int avc_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float lr, float beta1, float beta2, float eps, int64_t step, float grad_scale,
                  avc_stream_t stream);
/* Same update with the step counter and the learning rate in DEVICE memory, so that the call can be captured
 * once in a CUDA graph and replayed: state[0] = step count so far (incremented by the call), state[1] = lr
 * (written by the host before each replay), state[2..3] = scratch. */
int avc_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                      float* state, float beta1, float beta2, float eps, float grad_scale, avc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Camera rays: SMPL_Dataset.gen_rays_pose / gen_rays_silhouettes + near_far_from_sphere
 * (models/dataset.py:252-293, 331-342).  pose_c2w is a HOST pointer to the 4x4 row-major camera-to-world matrix
 * (lookat, models/utils.py:9-27).  Pixel grid: linspace(0, full-1, W) x linspace(0, full-1, H).  pix[R] lists the
 * selected canvas pixels (y*W + x; the True entries of the dilated mask) or is NULL for all W*H pixels.
 * ------------------------------------------------------------------------------------------ */
int avc_gen_rays(const float* pose_c2w, float fx, float fy, float cx, float cy, int32_t full_w, int32_t full_h,
                 int32_t W, int32_t H, const int32_t* pix, int32_t R, float* rays_o, float* rays_d, float* near,
                 float* far, avc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SMPL linear-blend skinning: my_lbs(v_shaped, pose, v_template, shapedirs, posedirs, J_regressor, parents,
 * lbs_weights, pose2rot) of models/utils.py:176-224 (batch 1; v_template / shapedirs are unused by the reference
 * function and therefore absent here).  pose: [n_joints*3] axis-angle when pose2rot != 0, else [n_joints][3][3]
 * rotation matrices.  posedirs: [(n_joints-1)*9][V*3].  lbs_weights: [V][n_joints].  n_joints <= 32.
 * Outputs: verts_out [V][3], joints_out [n_joints][3] (J_transformed).
 * ------------------------------------------------------------------------------------------ */
int avc_lbs_workspace_bytes(int32_t n_joints, size_t* bytes);
int avc_lbs_fwd(const float* v_shaped, const float* pose, int32_t pose2rot, const float* J_regressor,
                const int32_t* parents, const float* posedirs, const float* lbs_weights, int32_t V,
                int32_t n_joints, float* verts_out, float* joints_out, void* workspace,
                size_t workspace_bytes, avc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Self-test of the tcgen05 GEMM tiles (engine 1): C[M][N] = A[M][K] . B[N][K]^T, fp32 in/out, operands
 * split into bf16 (hi, lo) pairs; nprod = 1 (hi*hi only) or 3 (hi*hi + hi*lo + lo*hi).
 * workspace >= 4 * (M + N) * round_up(K, 8) + 1024 bytes.
 * ------------------------------------------------------------------------------------------ */
int avc_tc_gemm_nt_test(const float* A, const float* B, int64_t M, int32_t N, int32_t K, int32_t nprod,
                        float* C, void* workspace, size_t workspace_bytes, avc_stream_t stream);
/* Same for the reduction-over-rows tiles (weight gradients): C[N1][N2] += A[P][N1]^T . B[P][N2]; when colsum is not
 * NULL also colsum[N1] += column sums of A (the fused bias gradient).
 * workspace >= 4 * P * (round_up(N1,8) + round_up(N2,8)) + 2048 bytes. */
int avc_tc_gemm_tn_test(const float* A, const float* B, int64_t P, int32_t N1, int32_t N2, int32_t nprod,
                        float* C, float* colsum, void* workspace, size_t workspace_bytes, avc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Per-step view preparation of Runner.train_clip, all on the device.
 *
 * avc_raster_template replaces render_one_batch (models/utils.py:108-125 -> neural_renderer Renderer(camera_mode=
 * 'look'), image_size 256, 2x anti-aliasing, white textures, ambient 0.5 + directional 0.5 along +y, fill_back):
 * verts [V][3] (SMPL frame; the (x, z, -y) re-orientation of utils.py:115-119 happens inside), faces [F][3] int32,
 * eye / at [3] HOST pointers; rgb_out [n][n][3] already flipped like utils.py:124; mask_out [n][n] = (rgb != 0)
 * (main.py:361-364).  avc_dilate_count = ndimage.binary_dilation(mask, 3x3 full structure, iterations) and the
 * pixel count that sizes the canvas (models/dataset.py:255-258).  avc_mask_compact = nearest resize of the dilated
 * mask to W x W (F.interpolate default) + the row-major list of its True pixels (dataset.py:269-273); pix holds at
 * most `cap` entries, count_out the true count.  avc_view_targets = main.py:375-380,407-410 (nearest resize of the
 * template render, mask = channel 0 != 0, or all ones when threshold_mask == 0 i.e. mask_weight == 0).
 * avc_background_field = main.py:392-402: kind 1 clamp(N(0.5, 0.2), 0, 1) per pixel, kind 2 the 0.2 / 0.8
 * chessboard of chess_len-pixel squares blurred by GaussianBlur(kernel (5, 9), sigma); optional gather to the rays
 * (main.py:412-413).  avc_uniform_fill: counter-based U[lo, hi) draws (per-ray jitter, renderer.py:317-319).
 * ------------------------------------------------------------------------------------------ */
int avc_raster_workspace_bytes(int32_t V, int32_t image_size, int32_t supersample, size_t* bytes);
int avc_raster_template(const float* verts, const int32_t* faces, int32_t V, int32_t F, const float* eye,
                        const float* at, int32_t image_size, int32_t supersample, float* rgb_out, uint8_t* mask_out,
                        void* workspace, size_t workspace_bytes, avc_stream_t stream);
int avc_dilate_count(const uint8_t* mask, int32_t n, int32_t iterations, uint8_t* dilated, int32_t* count_out,
                     avc_stream_t stream);
int avc_mask_compact(const uint8_t* dilated, int32_t n, int32_t W, int32_t cap, uint8_t* in_mask, int32_t* pix,
                     int32_t* count_out, avc_stream_t stream);
int avc_view_targets(const float* rgb, int32_t n, int32_t W, int32_t threshold_mask, float* true_rgb, float* mask,
                     avc_stream_t stream);
int avc_background_field(int32_t kind, int32_t H, int32_t W, uint32_t seed, int32_t chess_len, float sigma,
                         float* canvas_bg, const int32_t* pix, int32_t R, float* ray_bg, avc_stream_t stream);
int avc_uniform_fill(uint32_t seed, int32_t n, float lo, float hi, float* out, avc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Iso-surface extraction for NeuSRenderer.extract_geometry (models/renderer.py:27-36,399-404; the reference calls
 * PyMCubes' marching_cubes, third-party, absent).  Marching tetrahedra over the [nx][ny][nz] field u (= -sdf):
 * avc_march_count writes the number of triangles of every grid cube ((nx-1)(ny-1)(nz-1) ints, x-major like the
 * field); the caller turns them into exclusive offsets; avc_march_emit writes 3 vertices per triangle in INDEX
 * coordinates (the caller rescales like renderer.py:33-35) and, per vertex, a 64-bit key of the grid edge it lies
 * on (for welding).  Triangles are oriented with normals towards decreasing u.
 * ------------------------------------------------------------------------------------------ */
/* Profiling aid of the fused value-chain kernel: with AVC_CHAIN_DEBUG=1 in the environment block 0 records cycle
 * counters ([0] MMA warp waiting for its A operand, [1] for weight slabs, [2] MMA warp total, [3] tile-layers,
 * [4] epilogue waiting for the accumulator, [5] epilogue work); this call synchronises the device and copies them. */
int avc_chain_debug_read(long long* out8);
/* Stall probe of the tcgen05 NT tiles: only in a diagnostic build (-DAVC_NT_PROBE=1, tools/nt_probe.py); a regular
 * build returns AVC_E_BADCFG.  host_out[16][8]: per epilogue functor the summed cycles {TMA warp waiting for a free
 * stage, TMA loop, MMA warp waiting for a drained accumulator, MMA warp waiting for operands, MMA loop, one epilogue
 * warp waiting for the accumulator, its loop, CTAs}; reset != 0 clears the counters. */
int avc_nt_probe_read(unsigned long long* host_out, int reset);

int avc_march_count(const float* field, int32_t nx, int32_t ny, int32_t nz, float iso, int32_t* counts,
                    avc_stream_t stream);
int avc_march_emit(const float* field, int32_t nx, int32_t ny, int32_t nz, float iso, const int32_t* offsets,
                   float* verts, int64_t* keys, avc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AVC_B200_H */
