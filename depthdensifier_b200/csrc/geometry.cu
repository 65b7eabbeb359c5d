// Stand-alone projection helpers of the reference script, kept as C-ABI entry points so the drop-in
// `project_points` / `unproject_points` (scripts/test.py:58-76, :79-90) also run on the device, in float64
// like the reference (pycolmap matrices and numpy promote everything to float64 there).
#include "common.cuh"

namespace ddn {

// scripts/test.py:58-76: Xc = [R|t] [X;1]; depths = Xc.z; uv = (K (Xc / (z + 1e-8)))[:2].  No validity
// handling - the caller gates on depth > 0.
__global__ void project_points_kernel(int64_t n, const double* __restrict__ xyz, const double* __restrict__ pose,
                                      const double* __restrict__ kmat, double* __restrict__ uv, double* __restrict__ depth) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = xyz[i * 3 + 0], y = xyz[i * 3 + 1], z = xyz[i * 3 + 2];
  double c[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) c[r] = fma(pose[r * 4 + 2], z, fma(pose[r * 4 + 1], y, pose[r * 4 + 0] * x)) + pose[r * 4 + 3];
  const double den = c[2] + 1e-8;
  const double xn = c[0] / den, yn = c[1] / den, zn = c[2] / den;
  uv[i * 2 + 0] = fma(kmat[2], zn, fma(kmat[1], yn, kmat[0] * xn));
  uv[i * 2 + 1] = fma(kmat[5], zn, fma(kmat[4], yn, kmat[3] * xn));
  depth[i] = c[2];
}

// scripts/test.py:79-90: ((u - cx) / fx * d, (v - cy) / fy * d, d) in float64; d is the float32 depth.
__global__ void unproject_points_kernel(int64_t n, const double* __restrict__ uv, const float* __restrict__ depth, double fx,
                                        double fy, double cx, double cy, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double d = (double)depth[i];
  out[i * 3 + 0] = __dmul_rn(__ddiv_rn(uv[i * 2 + 0] - cx, fx), d);
  out[i * 3 + 1] = __dmul_rn(__ddiv_rn(uv[i * 2 + 1] - cy, fy), d);
  out[i * 3 + 2] = d;
}

}  // namespace ddn

extern "C" {

int ddn_project_points(int64_t n_points, const double* points3d, const double* cam_from_world, const double* kmat,
                       double* points2d, double* depths, void* stream) {
  using namespace ddn;
  DDN_REQUIRE(n_points >= 0, "n_points");
  if (n_points == 0) return DDN_OK;
  DDN_REQUIRE(points3d && cam_from_world && kmat && points2d && depths, "null pointer");
  project_points_kernel<<<(unsigned)((n_points + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_points, points3d, cam_from_world,
                                                                                            kmat, points2d, depths);
  return after_launch("project_points_kernel");
}

int ddn_unproject_points(int64_t n_points, const double* points2d, const float* depth, const double* params4, double* points3d_cam,
                         void* stream) {
  using namespace ddn;
  DDN_REQUIRE(n_points >= 0, "n_points");
  if (n_points == 0) return DDN_OK;
  DDN_REQUIRE(points2d && depth && params4 && points3d_cam, "null pointer");
  DDN_REQUIRE(params4[0] != 0.0 && params4[1] != 0.0, "focal length is zero");
  unproject_points_kernel<<<(unsigned)((n_points + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      n_points, points2d, depth, params4[0], params4[1], params4[2], params4[3], points3d_cam);
  return after_launch("unproject_points_kernel");
}

}  // extern "C"
