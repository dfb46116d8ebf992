// avc_rays.cu -- camera rays of SMPL_Dataset.gen_rays_pose / gen_rays_silhouettes and near_far_from_sphere
// (AvatarGen/AppearanceGen/models/dataset.py:252-293, 331-342) for a list of canvas pixels.
#include "avc_common.cuh"

using namespace avc;

namespace {

struct RayCam {
  float pose[12];        // rows of the 3x4 camera-to-world matrix (lookat, models/utils.py:9-27)
  float fx, fy, cx, cy;  // intrinsics K (dataset.py:243-247): focal = .5 W / tan(.5 camera_angle_x), cx = .5 W
  float full_w, full_h;  // resolution the intrinsics refer to (256)
  int W, H;              // canvas resolution: pixel grid linspace(0, full-1, W) x linspace(0, full-1, H) (:259-260)
};

__device__ __forceinline__ float linspace_at(float end, int n, int j) {
  if (n == 1) return 0.f;
  float step = end / (float)(n - 1);
  return (j < n / 2) ? step * (float)j : end - step * (float)(n - 1 - j);      // at::linspace (two-sided)
}

__global__ void k_gen_rays(RayCam c, const int* __restrict__ pix, int R, float* __restrict__ rays_o,
                           float* __restrict__ rays_d, float* __restrict__ near, float* __restrict__ far) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  int p = pix ? pix[r] : r;
  int y = p / c.W, x = p - y * c.W;
  float px = linspace_at(c.full_w - 1.f, c.W, x), py = linspace_at(c.full_h - 1.f, c.H, y);
  float v0 = (px - c.cx) / c.fx, v1 = -(py - c.cy) / c.fy, v2 = -1.f;           // dataset.py:264-266
  float inv = 1.0f / sqrtf(v0 * v0 + v1 * v1 + v2 * v2);
  v0 *= inv; v1 *= inv; v2 *= inv;
  float d[3], o[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    d[i] = v0 * c.pose[i * 4] + v1 * c.pose[i * 4 + 1] + v2 * c.pose[i * 4 + 2];   // rays_v @ R^T (:268)
    o[i] = c.pose[i * 4 + 3];
    rays_d[r * 3 + i] = d[i];
    rays_o[r * 3 + i] = o[i];
  }
  // near_far_from_sphere (:331-342): mid = -(o.d)/(d.d); near = max(mid - 1, 0); far = mid + 1
  float a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  float b = 2.f * (o[0] * d[0] + o[1] * d[1] + o[2] * d[2]);
  float mid = 0.5f * (-b) / a;
  near[r] = fmaxf(mid - 1.f, 0.f);
  far[r] = mid + 1.f;
}

}  // namespace

extern "C" int avc_gen_rays(const float* pose_c2w /*host, 16 floats row-major*/, float fx, float fy, float cx, float cy,
                            int32_t full_w, int32_t full_h, int32_t W, int32_t H, const int32_t* pix, int32_t R,
                            float* rays_o, float* rays_d, float* near, float* far, avc_stream_t stream) {
  if (!pose_c2w || !rays_o || !rays_d || !near || !far) return AVC_E_NULL;
  if (R < 1 || W < 1 || H < 1 || full_w < 1 || full_h < 1) return AVC_E_SIZE;
  if (!pix && R != W * H) return AVC_E_SIZE;
  RayCam c;
  for (int i = 0; i < 12; ++i) c.pose[i] = pose_c2w[i];
  c.fx = fx; c.fy = fy; c.cx = cx; c.cy = cy; c.full_w = (float)full_w; c.full_h = (float)full_h; c.W = W; c.H = H;
  k_gen_rays<<<(R + 255) / 256, 256, 0, (cudaStream_t)stream>>>(c, pix, R, rays_o, rays_d, near, far);
  AVC_LAUNCH_TRY();
  return 0;
}
