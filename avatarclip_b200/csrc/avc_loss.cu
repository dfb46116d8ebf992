// avc_loss.cu -- shading, canvas scatter and the non-CLIP losses of Runner.train_clip
// (AvatarGen/AppearanceGen/main.py:417-497, 528-534), forward and backward, as two small fused kernels
// each way (the reference runs ~60 eager kernels with boolean-mask scatters here).
//
//   forward   k_canvas_bg   every pixel: background into both canvases; loss terms of pixels without a ray
//             k_shade_fwd   one warp per ray: normal = sum_s w g, Lambert shading, the two shaded colours,
//                           scatter to the canvases, L1 / BCE / PSNR partial sums
//             k_loss_final  scalars
//   backward  k_shade_bwd   one warp per ray: gather d canvas, back through the shading to the cotangents
//                           of color_fine / extra_color_fine / gradients / weights / weight_sum
// The ablation confs (confs/ablation/*_0..2.conf) switch train.texture_cast_light and / or train.add_no_texture off:
// `plain_texture` puts the un-shaded extra colour on canvas 0 (main.py:515-520), `no_shading_term` drops the CLIP term on
// canvas 1 (main.py:521,533).
#include "avc_common.cuh"

using namespace avc;

namespace {

enum { LS_COLOR = 0, LS_EIK = 1, LS_BCE = 2, LS_PSNR = 3, LS_BASE = 4, LS_L1SUM = 5, LS_BCESUM = 6, LS_SQSUM = 7,
       LS_MASKSUM = 8 };

__device__ __forceinline__ float bce_term(float ws, float m) {
  float c = fminf(fmaxf(ws, 1e-3f), 1.0f - 1e-3f);                   // main.py:497 clip
  return -(m * logf(c) + (1.f - m) * logf(1.f - c));                 // F.binary_cross_entropy
}

__global__ void __launch_bounds__(256)
k_canvas_bg(avc_loss_inputs in, float* __restrict__ canvases, float* __restrict__ scalars) {
  const int HW = in.H * in.W;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float l1 = 0.f, bce = 0.f, sq = 0.f, ms = 0.f;
  if (i < HW) {
    float bg = 0.f;
    if (in.bg_choice == 0) bg = 1.f;
    else if ((in.bg_choice == 1 || in.bg_choice == 2) && in.background) bg = in.background[i];
    const bool has_ray = in.in_mask[i] != 0;
    if (!has_ray) {
      float* t = canvases + (size_t)i * 3;
      float* s = canvases + (size_t)HW * 3 + (size_t)i * 3;
      t[0] = t[1] = t[2] = bg;
      s[0] = s[1] = s[2] = bg;
      float m = in.mask[i];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float e = (0.f - in.true_rgb[(size_t)i * 3 + c]);               // full_color_fine is 0 off the mask (:479)
        l1 += fabsf(e * m);
        sq += e * e * m;
      }
      bce = bce_term(0.f, m);                                          // full_weight_sum is 0 off the mask (:483)
    }
    ms = in.mask[i];
  }
  l1 = warp_sum(l1); bce = warp_sum(bce); sq = warp_sum(sq); ms = warp_sum(ms);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(scalars + LS_L1SUM, l1); atomicAdd(scalars + LS_BCESUM, bce);
    atomicAdd(scalars + LS_SQSUM, sq); atomicAdd(scalars + LS_MASKSUM, ms);
  }
}

struct Shade {
  float n[3], r, nh[3], lh[3], dot, diff, shade, shade2, amb;
  bool low, nan_;
};

// per-view draws: device copy (graph replay) when given, else the host fields
__device__ __forceinline__ void view_draws(const avc_loss_inputs& in, float ld[3], float* amb) {
  if (in.view_scalars) {
    ld[0] = in.view_scalars[0]; ld[1] = in.view_scalars[1]; ld[2] = in.view_scalars[2]; *amb = in.view_scalars[3];
  } else {
    ld[0] = in.light_dir[0]; ld[1] = in.light_dir[1]; ld[2] = in.light_dir[2]; *amb = in.ambience;
  }
}

__device__ __forceinline__ Shade shade_terms(const avc_loss_inputs& in, const float n[3], float wsum) {
  Shade s;
  float ldir[3], amb;
  view_draws(in, ldir, &amb);
  s.n[0] = n[0]; s.n[1] = n[1]; s.n[2] = n[2];
  s.r = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  float ln = sqrtf(ldir[0] * ldir[0] + ldir[1] * ldir[1] + ldir[2] * ldir[2]);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    s.nh[a] = n[a] / (s.r + 1e-7f);                                    // main.py:430
    s.lh[a] = ldir[a] / (ln + 1e-7f);                                  // :436
  }
  s.dot = s.nh[0] * s.lh[0] + s.nh[1] * s.lh[1] + s.nh[2] * s.lh[2];
  s.nan_ = isnan(s.dot);
  s.diff = s.nan_ ? 1.0f : fminf(fmaxf(s.dot, 0.f), 1.f);             // :438-439
  s.shade = amb + (1.f - amb) * s.diff;                               // :440-442
  s.amb = amb;
  s.low = wsum < 0.5f;
  s.shade2 = s.low ? 1.0f : s.shade;                                  // :450-452 (l_ratio = 1)
  return s;
}

__global__ void __launch_bounds__(256)
k_shade_fwd(avc_loss_inputs in, float* __restrict__ canvases, float* __restrict__ scalars) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= in.R) return;
  const int lane = threadIdx.x & 31;
  const int HW = in.H * in.W;
  float n[3] = {0.f, 0.f, 0.f};
  for (int j = lane; j < in.S; j += 32) {
    float w = in.weights[(size_t)r * in.S + j];
    const float* g = in.gradients + ((size_t)r * in.S + j) * 3;
    n[0] = fmaf(w, g[0], n[0]); n[1] = fmaf(w, g[1], n[1]); n[2] = fmaf(w, g[2], n[2]);   // main.py:427-429
  }
  n[0] = warp_sum(n[0]); n[1] = warp_sum(n[1]); n[2] = warp_sum(n[2]);
  if (lane != 0) return;
  const float wsum = in.weight_sum[r];
  Shade s = shade_terms(in, n, wsum);
  const int p = in.pix[r];
  float* t = canvases + (size_t)p * 3;
  float* sh = canvases + (size_t)HW * 3 + (size_t)p * 3;
  float l1 = 0.f, sq = 0.f;
  const float m = in.mask[p];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float ex = in.extra_color_fine[(size_t)r * 3 + c];
    t[c] = in.plain_texture ? ex                                       // full_extra_color_fine (:475-477), cast light off
                            : fminf(fmaxf(ex * s.shade2, 0.f), 1.f);   // texture_shading (:453)
    sh[c] = s.low ? ex : s.shade;                                      // rand_shading_rgb (:445-448)
    float e = in.color_fine[(size_t)r * 3 + c] - in.true_rgb[(size_t)p * 3 + c];
    l1 += fabsf(e * m);
    sq += e * e * m;
  }
  atomicAdd(scalars + LS_L1SUM, l1);
  atomicAdd(scalars + LS_SQSUM, sq);
  atomicAdd(scalars + LS_BCESUM, bce_term(wsum, m));
}

__global__ void k_loss_final(avc_loss_inputs in, float* __restrict__ scalars) {
  if (threadIdx.x != 0) return;
  float mask_sum = scalars[LS_MASKSUM] + 1e-5f;                        // main.py:417
  float color = scalars[LS_L1SUM] / mask_sum;                          // :491
  float bce = scalars[LS_BCESUM] / (float)(in.H * in.W);               // :497 (mean over the canvas)
  float eik = in.gradient_error[0];
  scalars[LS_COLOR] = color;
  scalars[LS_EIK] = eik;
  scalars[LS_BCE] = bce;
  scalars[LS_PSNR] = 20.0f * log10f(1.0f / sqrtf(scalars[LS_SQSUM] / (mask_sum * 3.0f)));   // :492
  scalars[LS_BASE] = color + eik * in.igr_weight + bce * in.mask_weight;                       // :528-531
}

__global__ void __launch_bounds__(256)
k_shade_bwd(avc_loss_inputs in, const float* __restrict__ d_canvases, const float* __restrict__ scalars,
            avc_neus_cotangents cot) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= in.R) return;
  const int lane = threadIdx.x & 31;
  const int HW = in.H * in.W;
  float n[3] = {0.f, 0.f, 0.f};
  for (int j = lane; j < in.S; j += 32) {
    float w = in.weights[(size_t)r * in.S + j];
    const float* g = in.gradients + ((size_t)r * in.S + j) * 3;
    n[0] = fmaf(w, g[0], n[0]); n[1] = fmaf(w, g[1], n[1]); n[2] = fmaf(w, g[2], n[2]);
  }
  n[0] = warp_sum(n[0]); n[1] = warp_sum(n[1]); n[2] = warp_sum(n[2]);
  const float wsum = in.weight_sum[r];
  Shade s = shade_terms(in, n, wsum);
  const int p = in.pix[r];
  const float* dt = d_canvases + (size_t)p * 3;
  const float* ds = d_canvases + (size_t)HW * 3 + (size_t)p * 3;
  const float m = in.mask[p];
  const float mask_sum = scalars[LS_MASKSUM] + 1e-5f;
  float d_shade = 0.f;      // adjoint of `shade` (the un-overridden Lambert term)
  float dex[3], dcol[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float ex = in.extra_color_fine[(size_t)r * 3 + c];
    if (in.plain_texture) {
      dex[c] = dt[c];                                                  // canvas 0 is the extra colour itself
    } else {
      float prod = ex * s.shade2;
      float inr = (prod >= 0.f && prod <= 1.f) ? 1.f : 0.f;            // clamp(0,1) backward: inclusive, like torch
      dex[c] = dt[c] * inr * s.shade2;
      if (!s.low) d_shade += dt[c] * inr * ex;                         // shade2 = shade unless low
    }
    if (!in.no_shading_term) {                                         // rand_shading_rgb feeds a CLIP term (:521-526)
      if (s.low) dex[c] += ds[c]; else d_shade += ds[c];
    }
    float e = in.color_fine[(size_t)r * 3 + c] - in.true_rgb[(size_t)p * 3 + c];
    float sg = (e * m > 0.f) ? 1.f : ((e * m < 0.f) ? -1.f : 0.f);     // d|x|/dx, 0 at 0 (l1_loss)
    dcol[c] = sg * m / mask_sum;
  }
  // shade = amb + (1-amb) * clamp(dot, 0, 1) ; dot = nh . lh ; nh = n / (|n| + 1e-7)
  float d_dot = (!s.nan_ && s.dot >= 0.f && s.dot <= 1.f) ? d_shade * (1.f - s.amb) : 0.f;
  float dnh[3] = {d_dot * s.lh[0], d_dot * s.lh[1], d_dot * s.lh[2]};
  float ndn = s.n[0] * dnh[0] + s.n[1] * dnh[1] + s.n[2] * dnh[2];
  float re = s.r + 1e-7f;
  float dn[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) dn[a] = dnh[a] / re - ((s.r > 0.f) ? s.n[a] * ndn / (s.r * re * re) : 0.f);
  for (int j = lane; j < in.S; j += 32) {
    size_t o = (size_t)r * in.S + j;
    float w = in.weights[o];
    const float* g = in.gradients + o * 3;
    float* gg = const_cast<float*>(cot.gradients) + o * 3;
    gg[0] = w * dn[0]; gg[1] = w * dn[1]; gg[2] = w * dn[2];
    const_cast<float*>(cot.weights)[o] = g[0] * dn[0] + g[1] * dn[1] + g[2] * dn[2];
  }
  if (lane == 0) {
    float* gc = const_cast<float*>(cot.color_fine) + (size_t)r * 3;
    float* ge = const_cast<float*>(cot.extra_color_fine) + (size_t)r * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) { gc[c] = dcol[c]; ge[c] = dex[c]; }
    float dws = 0.f;
    if (wsum >= 1e-3f && wsum <= 1.0f - 1e-3f)                          // clip backward (inclusive bounds, like torch)
      dws = (-(m / wsum) + (1.f - m) / (1.f - wsum)) / (float)HW * in.mask_weight;
    const_cast<float*>(cot.weight_sum)[r] = dws;
    if (r == 0) const_cast<float*>(cot.gradient_error)[0] = in.igr_weight;
  }
}

int check_inputs(const avc_loss_inputs* in) {
  if (!in) return AVC_E_NULL;
  if (!in->color_fine || !in->extra_color_fine || !in->gradients || !in->weights || !in->weight_sum ||
      !in->gradient_error || !in->pix || !in->in_mask || !in->true_rgb || !in->mask)
    return AVC_E_NULL;
  if (in->R < 1 || in->S < 1 || in->H < 1 || in->W < 1) return AVC_E_SIZE;
  if (in->bg_choice < 0 || in->bg_choice > 3) return AVC_E_BADCFG;
  if ((in->bg_choice == 1 || in->bg_choice == 2) && !in->background) return AVC_E_NULL;
  if ((in->plain_texture | in->no_shading_term) & ~1) return AVC_E_BADCFG;
  return 0;
}

}  // namespace

extern "C" {

int avc_loss_stage_fwd(const avc_loss_inputs* in, float* canvases, float* scalars, avc_stream_t stream) {
  AVC_TRY(check_inputs(in));
  if (!canvases || !scalars) return AVC_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  AVC_CUDA_TRY(cudaMemsetAsync(scalars, 0, sizeof(float) * AVC_LOSS_SCALARS, st));
  const int HW = in->H * in->W;
  k_canvas_bg<<<(HW + 255) / 256, 256, 0, st>>>(*in, canvases, scalars);
  k_shade_fwd<<<(in->R + 7) / 8, 256, 0, st>>>(*in, canvases, scalars);
  k_loss_final<<<1, 32, 0, st>>>(*in, scalars);
  AVC_LAUNCH_TRY();
  return 0;
}

int avc_loss_stage_bwd(const avc_loss_inputs* in, const float* d_canvases, const float* scalars,
                       const avc_neus_cotangents* cot, avc_stream_t stream) {
  AVC_TRY(check_inputs(in));
  if (!d_canvases || !scalars || !cot) return AVC_E_NULL;
  if (!cot->color_fine || !cot->extra_color_fine || !cot->gradients || !cot->weights || !cot->weight_sum ||
      !cot->gradient_error)
    return AVC_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  k_shade_bwd<<<(in->R + 7) / 8, 256, 0, st>>>(*in, d_canvases, scalars, *cot);
  AVC_LAUNCH_TRY();
  return 0;
}

}  // extern "C"
