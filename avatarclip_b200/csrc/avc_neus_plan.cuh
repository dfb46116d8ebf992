// avc_neus_plan.cuh -- host-side shape bookkeeping for the NeuS path: network dimensions, the flat
// parameter layout (reference state-dict order), the packed-weight layout and the workspace layout.
#pragma once
#include "avc_common.cuh"
#include "avc_neus_kernels.cuh"   // Split16

namespace avc {

constexpr int kMaxLin = 16;     // linears per network
constexpr int kMaxSampleBlocks = 8;  // composite kernels keep ceil(S/32) <= 8 values per lane

struct LinDim {
  int K;     // input width (after the optional skip concat)
  int N;     // output width
  int Kp;    // K rounded up to 4 (leading dimension of the input activation buffer)
  int Np;    // N rounded up to 4 (leading dimension of pre-activation / gradient buffers)
  bool skip; // models/fields.py:79-80: input is cat([x, enc]) / sqrt(2)
  int64_t off_g, off_v, off_b;     // offsets into the flat parameter vector
  // packed (effective, weight-normed) copies, offsets in floats into the pack buffer
  int64_t pk_W;    // [N rows][Kp]   (for the last SDF linear: rows 1.. only, i.e. F rows)
  int64_t pk_WT;   // [K rows][Np]   (for the last SDF linear: [K][Fp])
  int64_t pk_b;    // [Np]
};

struct NeusPlan {
  avc_neus_cfg cfg;
  int E, EP;            // encoded width (39) and its padded leading dimension (40)
  int H, L;             // SDF hidden width; index of the last SDF linear (= n_layers)
  int F, Fp;            // feature width = d_out - 1
  LinDim sdf[kMaxLin];  // sdf[L] describes the full last linear (N = d_out); pk_W/pk_WT hold rows 1..
  int64_t pk_wsdf;      // [Kp_L] row 0 of the last linear;   pk_bsdf: its bias (1 float)
  int64_t pk_bsdf;
  int Hc, Lc;           // colour hidden width; index of the colour head linear (= col_n_layers)
  LinDim col[kMaxLin];  // col[0].K = 6 + F;  packed: pk_W = [Hc][Fp] (feature columns), pk_WT = [F][Hc]
  int64_t pk_c0x;       // [Hc][8]  columns 0..5 of colour lin0 (points, normals), zero padded
  int64_t pk_c0xT;      // [8][Hc]
  LinDim extra;         // extra_lin (3 x Hc)
  int64_t pk_W6;        // [8][Hc]  rows 0-2 = head lin (lin{Lc}), rows 3-5 = extra_lin, rows 6-7 zero
  int64_t pk_b6;        // [8]
  int64_t off_var;      // variance scalar in the flat parameter vector
  int64_t n_params;
  int64_t pack_floats;
  int S, n0, per, steps;
};

static inline int build_plan(const avc_neus_cfg* c, NeusPlan* p) {
  if (!c || !p) return AVC_E_NULL;
  p->cfg = *c;
  if (c->sdf_d_in != 3 || c->sdf_multires < 0 || c->sdf_multires > 10) return AVC_E_BADCFG;
  if (c->sdf_n_layers < 1 || c->sdf_n_layers + 1 > kMaxLin) return AVC_E_BADCFG;
  if (c->col_n_layers < 1 || c->col_n_layers + 1 > kMaxLin) return AVC_E_BADCFG;
  if (c->sdf_d_hidden % 4 || c->col_d_hidden % 4 || c->sdf_d_hidden <= 0 || c->col_d_hidden <= 0) return AVC_E_BADCFG;
  if (c->sdf_d_out < 2 || (c->sdf_d_out - 1) % 4 || c->col_d_feature != c->sdf_d_out - 1) return AVC_E_BADCFG;
  if (!(c->sdf_scale > 0.f)) return AVC_E_BADCFG;
  if (c->n_samples < 2 || c->n_importance < 0 || c->up_sample_steps < 1) return AVC_E_BADCFG;
  if (c->n_importance % c->up_sample_steps) return AVC_E_BADCFG;
  if (c->engine != 0 && c->engine != 1) return AVC_E_BADCFG;
  if (c->color_products != 0 && c->color_products != 1 && c->color_products != 3) return AVC_E_BADCFG;
  if (c->wgrad_products != 0 && c->wgrad_products != 1 && c->wgrad_products != 3) return AVC_E_BADCFG;
  if (c->engine == 1 && (c->sdf_d_hidden % 8 || c->col_d_hidden % 8 || (c->sdf_d_out - 1) % 8)) return AVC_E_BADCFG;
  p->E = 3 * (1 + 2 * c->sdf_multires);
  p->EP = (int)round_up(p->E, 8);   // multiples of 8: bf16 rows stay 16-byte aligned for TMA
  p->H = c->sdf_d_hidden;
  p->L = c->sdf_n_layers;
  p->F = c->sdf_d_out - 1;
  p->Fp = p->F;
  p->Hc = c->col_d_hidden;
  p->Lc = c->col_n_layers;
  p->n0 = c->n_samples;
  p->steps = c->up_sample_steps;
  p->per = c->n_importance / c->up_sample_steps;
  p->S = c->n_samples + c->n_importance;
  if (p->S > 32 * kMaxSampleBlocks) return AVC_E_BADCFG;
  if ((c->sdf_skip_mask & 1u) || (c->sdf_skip_mask >> (p->L + 1))) return AVC_E_BADCFG;

  int64_t po = 0, pk = 0;
  auto take_pk = [&](int64_t n) { int64_t r = pk; pk += round_up(n, 8); return r; };
  // ---- SDF linears (models/fields.py:24-43)
  for (int l = 0; l <= p->L; ++l) {
    LinDim& d = p->sdf[l];
    d.skip = (c->sdf_skip_mask >> l) & 1u;
    d.K = (l == 0) ? p->E : p->H;
    bool next_skip = (l + 1 <= p->L) && ((c->sdf_skip_mask >> (l + 1)) & 1u);
    d.N = (l == p->L) ? c->sdf_d_out : (next_skip ? p->H - p->E : p->H);
    if (d.N <= 0) return AVC_E_BADCFG;
    if (d.skip && l == 0) return AVC_E_BADCFG;
    d.Kp = (int)round_up(d.K, 8);
    d.Np = (int)round_up(d.N, 8);
    d.off_g = po; po += d.N;
    d.off_v = po; po += (int64_t)d.N * d.K;
    d.off_b = po; po += d.N;
    if (l < p->L) {
      d.pk_W = take_pk((int64_t)d.N * d.Kp);
      d.pk_WT = take_pk((int64_t)d.Kp * d.Np);
      d.pk_b = take_pk(d.Np);
    } else {
      d.pk_W = take_pk((int64_t)p->F * d.Kp);
      d.pk_WT = take_pk((int64_t)d.Kp * p->Fp);
      d.pk_b = take_pk(p->Fp);
      p->pk_wsdf = take_pk(d.Kp);
      p->pk_bsdf = take_pk(4);
    }
  }
  // ---- colour linears (models/fields.py:126-149)
  for (int l = 0; l <= p->Lc; ++l) {
    LinDim& d = p->col[l];
    d.skip = false;
    d.K = (l == 0) ? 6 + p->F : p->Hc;
    d.N = (l == p->Lc) ? 3 : p->Hc;
    d.Kp = (int)round_up(d.K, 8);
    d.Np = (int)round_up(d.N, 8);
    d.off_g = po; po += d.N;
    d.off_v = po; po += (int64_t)d.N * d.K;
    d.off_b = po; po += d.N;
    if (l == 0) {
      d.pk_W = take_pk((int64_t)p->Hc * p->Fp);
      d.pk_WT = take_pk((int64_t)p->Fp * p->Hc);
      d.pk_b = take_pk(p->Hc);
      p->pk_c0x = take_pk((int64_t)p->Hc * 8);
      p->pk_c0xT = take_pk((int64_t)8 * p->Hc);
    } else if (l < p->Lc) {
      d.pk_W = take_pk((int64_t)p->Hc * p->Hc);
      d.pk_WT = take_pk((int64_t)p->Hc * p->Hc);
      d.pk_b = take_pk(p->Hc);
    } else {
      d.pk_W = d.pk_WT = d.pk_b = -1;   // heads are packed together into W6/b6
    }
  }
  {
    LinDim& d = p->extra;
    d.skip = false; d.K = p->Hc; d.N = 3; d.Kp = p->Hc; d.Np = 4;
    d.off_g = po; po += 3;
    d.off_v = po; po += (int64_t)3 * p->Hc;
    d.off_b = po; po += 3;
    d.pk_W = d.pk_WT = d.pk_b = -1;
  }
  p->pk_W6 = take_pk((int64_t)8 * p->Hc);
  p->pk_b6 = take_pk(8);
  p->off_var = po; po += 1;
  p->n_params = po;
  p->pack_floats = pk;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Workspace layout for one chunk of Rc rays (P = Rc * S points).
// ---------------------------------------------------------------------------------------------
struct NeusWs {
  // header (persistent across the forward/backward pair)
  float* ctx;        // [64] scalars: see CTX_* below
  float* pack;       // packed effective weights
  float* wbar;       // [n_params] dense dW (in the v slots) / db (in the b slots) accumulators
  float* ray_part;   // [Rmax_total? no: per chunk] [Rc][4] per-ray partial sums (eikonal num/den, inv_s bar)
  // sampling (sample-major [S][Rc])
  float *zA, *zB, *sA, *sB, *wS, *newZ, *newS;
  // per point (ray-major), forward stash
  float* cin;        // [P][8]  x(3) n(3) 0 0
  float* in[kMaxLin];     // in[l]: [P][Kp_l]
  float* z[kMaxLin];      // z[l]:  [P][Np_l], l < L
  float* qt[kMaxLin];     // qt[l]: [P][Np_l], l < L
  float* sdf;        // [P]
  float* feat;       // [P][Fp]
  float* ge;         // [P][EP]
  float* ch[kMaxLin];     // ch[l]: [P][Hc] post-ReLU activation feeding colour linear l (l = 1..Lc)
  float* rgb6;       // [P][8]
  // backward temporaries
  float* y6bar;      // [P][8]
  float* sdfbar;     // [P]
  float* nbar;       // [P][4]
  float* gebar;      // [P][EP]
  float* featbar;    // [P][Fp]
  float* cbar[2];    // [P][Hc]
  float* ubar[2];    // [P][max(Kp)]
  float* zbar[kMaxLin];   // [P][Np_l], l < L
  // tcgen05 engine: two-term bf16 copies of every GEMM operand (same leading dimensions); hi == nullptr otherwise
  __nv_bfloat16 *pk_hi, *pk_lo;            // split of the whole packed-weight buffer (same offsets as `pack`)
  Split16 in16[kMaxLin], qt16[kMaxLin], zbar16[kMaxLin], ch16[kMaxLin];
  Split16 feat16, featbar16, cbar16[2], ubar16[2];
  size_t bytes;
  int64_t Rc, P;
};


static inline void carve_ws(const NeusPlan& pl, int64_t Rc, void* base, NeusWs* w) {
  Carver c(base);
  const int64_t P = Rc * pl.S;
  w->Rc = Rc; w->P = P;
  w->ctx = c.take<float>(CTX_FLOATS);
  w->pack = c.take<float>(pl.pack_floats);
  w->wbar = c.take<float>(pl.n_params);
  w->ray_part = c.take<float>(Rc * 4);
  const int64_t SR = (int64_t)pl.S * Rc;
  w->zA = c.take<float>(SR); w->zB = c.take<float>(SR);
  w->sA = c.take<float>(SR); w->sB = c.take<float>(SR);
  w->wS = c.take<float>(SR);
  w->newZ = c.take<float>((int64_t)pl.per * Rc + 4);
  w->newS = c.take<float>((int64_t)pl.per * Rc + 4);
  w->cin = c.take<float>(P * 8);
  int maxK = pl.EP;
  for (int l = 0; l <= pl.L; ++l) {
    w->in[l] = c.take<float>(P * pl.sdf[l].Kp);
    if (pl.sdf[l].Kp > maxK) maxK = pl.sdf[l].Kp;
  }
  for (int l = 0; l < pl.L; ++l) {
    w->z[l] = c.take<float>(P * pl.sdf[l].Np);
    w->qt[l] = c.take<float>(P * pl.sdf[l].Np);
    w->zbar[l] = c.take<float>(P * pl.sdf[l].Np);
  }
  w->sdf = c.take<float>(P);
  w->feat = c.take<float>(P * pl.Fp);
  w->ge = c.take<float>(P * pl.EP);
  for (int l = 1; l <= pl.Lc; ++l) w->ch[l] = c.take<float>(P * pl.Hc);
  w->ch[0] = nullptr;
  w->rgb6 = c.take<float>(P * 8);
  w->y6bar = c.take<float>(P * 8);
  w->sdfbar = c.take<float>(P);
  w->nbar = c.take<float>(P * 4);
  w->gebar = c.take<float>(P * pl.EP);
  w->featbar = c.take<float>(P * pl.Fp);
  w->cbar[0] = c.take<float>(P * pl.Hc);
  w->cbar[1] = c.take<float>(P * pl.Hc);
  w->ubar[0] = c.take<float>(P * maxK);
  w->ubar[1] = c.take<float>(P * maxK);
  const Split16 none = {nullptr, nullptr, 0};
  w->pk_hi = w->pk_lo = nullptr;
  for (int l = 0; l < kMaxLin; ++l) w->in16[l] = w->qt16[l] = w->zbar16[l] = w->ch16[l] = none;
  w->feat16 = w->featbar16 = w->cbar16[0] = w->cbar16[1] = w->ubar16[0] = w->ubar16[1] = none;
  if (pl.cfg.engine == 1) {
    auto take16 = [&](int64_t rows, int ld) {
      Split16 s;
      s.hi = c.take<__nv_bfloat16>(rows * ld);
      s.lo = c.take<__nv_bfloat16>(rows * ld);
      s.ld = ld;
      return s;
    };
    w->pk_hi = c.take<__nv_bfloat16>(pl.pack_floats);
    w->pk_lo = c.take<__nv_bfloat16>(pl.pack_floats);
    for (int l = 0; l <= pl.L; ++l) w->in16[l] = take16(P, pl.sdf[l].Kp);
    for (int l = 0; l < pl.L; ++l) { w->qt16[l] = take16(P, pl.sdf[l].Np); w->zbar16[l] = take16(P, pl.sdf[l].Np); }
    w->feat16 = take16(P, pl.Fp);
    w->featbar16 = take16(P, pl.Fp);
    for (int l = 1; l <= pl.Lc; ++l) w->ch16[l] = take16(P, pl.Hc);
    w->cbar16[0] = take16(P, pl.Hc); w->cbar16[1] = take16(P, pl.Hc);
    w->ubar16[0] = take16(P, maxK); w->ubar16[1] = take16(P, maxK);
  }
  w->bytes = c.used();
}

}  // namespace avc
