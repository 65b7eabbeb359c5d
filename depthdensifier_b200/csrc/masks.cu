// Gradient-mask and world-normal helpers next to the hot path:
//   ddn_gradient_mask      compute_depth_normal_gradient_mask, src/depthdensifier/initilizer.py:236-328:
//                          zero-padded separable Gaussian (torch conv2d), 3x3 Sobel magnitude relative to the
//                          smoothed depth > threshold, OR magnitude of torch.gradient over the normal
//                          channels > threshold
//   ddn_transform_normals  COLMAPVisualizer._transform_normals, src/depthdensifier/visualizer.py:346-376:
//                          n_world = R^T n_cam / (|R^T n_cam| + 1e-8) in float64
#include "common.cuh"

namespace ddn {

constexpr int kMaxTaps = 129;
struct Taps {
  float w[kMaxTaps];
  int n;
};

// one axis of the zero-padded correlation F.conv2d(x, k, padding=n/2), float32
__global__ void conv_axis_zero_kernel(int H, int W, int axis, Taps t, const float* __restrict__ in, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int r = t.n / 2;
  float acc = 0.f;
  for (int k = 0; k < t.n; ++k) {
    const int xx = axis == 1 ? x + k - r : x, yy = axis == 0 ? y + k - r : y;
    if (xx >= 0 && xx < W && yy >= 0 && yy < H) acc = __fadd_rn(acc, __fmul_rn(in[(size_t)yy * W + xx], t.w[k]));
  }
  out[(size_t)y * W + x] = acc;
}

__device__ __forceinline__ float at0(const float* __restrict__ f, int H, int W, int y, int x) {
  return (x >= 0 && x < W && y >= 0 && y < H) ? f[(size_t)y * W + x] : 0.f;
}

__device__ __forceinline__ float tgrad(const float* __restrict__ f, int i, int n, size_t stride) {  // torch.gradient, edge_order 1
  if (n == 1) return 0.f;
  if (i == 0) return f[stride] - f[0];
  if (i == n - 1) return f[(size_t)i * stride] - f[(size_t)(i - 1) * stride];
  return (f[(size_t)(i + 1) * stride] - f[(size_t)(i - 1) * stride]) * 0.5f;
}

__global__ void gradient_mask_kernel(int H, int W, const float* __restrict__ smooth, const float* __restrict__ normal,
                                     float depth_threshold, float normal_threshold, uint8_t* __restrict__ mask) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  bool e = false;
  if (smooth != nullptr) {
    // Sobel as cross-correlation with zero padding (initilizer.py:283-290)
    const float a = at0(smooth, H, W, y - 1, x - 1), b = at0(smooth, H, W, y - 1, x), c = at0(smooth, H, W, y - 1, x + 1);
    const float d = at0(smooth, H, W, y, x - 1), f = at0(smooth, H, W, y, x + 1);
    const float g = at0(smooth, H, W, y + 1, x - 1), h = at0(smooth, H, W, y + 1, x), i = at0(smooth, H, W, y + 1, x + 1);
    const float dx = (c - a) + 2.f * (f - d) + (i - g);
    const float dy = (g - a) + 2.f * (h - b) + (i - c);
    const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    const float rel = __fdiv_rn(mag, __fadd_rn(smooth[(size_t)y * W + x], 1e-6f));
    e = rel > depth_threshold;
  }
  if (normal != nullptr) {
    float acc = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float gx = tgrad(normal + (size_t)y * W * 3 + ch, x, W, 3);
      const float gy = tgrad(normal + (size_t)x * 3 + ch, y, H, (size_t)W * 3);
      acc = __fadd_rn(acc, __fmul_rn(gx, gx));
      acc = __fadd_rn(acc, __fmul_rn(gy, gy));
    }
    e |= __fsqrt_rn(acc) > normal_threshold;
  }
  mask[(size_t)y * W + x] = e ? 1 : 0;
}

struct Rot {
  double r[9];  // R_cam_from_world, row-major
};

__global__ void transform_normals_kernel(int64_t n, Rot R, const float* __restrict__ normal, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = normal[i * 3 + 0], b = normal[i * 3 + 1], c = normal[i * 3 + 2];
  // R_world_from_cam = R^T: row j of R^T = column j of R
  double w[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) w[j] = R.r[0 * 3 + j] * a + R.r[1 * 3 + j] * b + R.r[2 * 3 + j] * c;
  const double nrm = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]) + 1e-8;
  out[i * 3 + 0] = w[0] / nrm;
  out[i * 3 + 1] = w[1] / nrm;
  out[i * 3 + 2] = w[2] / nrm;
}

}  // namespace ddn

extern "C" {

int ddn_gradient_mask(int64_t height, int64_t width, const float* depth, const float* normal, const float* taps_host,
                      int32_t n_taps, float depth_threshold, float normal_threshold, uint8_t* mask_out, void* workspace,
                      int64_t workspace_bytes, void* stream) {
  using namespace ddn;
  DDN_REQUIRE(height > 0 && width > 0 && height * width < (1ll << 31), "shape");
  DDN_REQUIRE(mask_out != nullptr && (depth != nullptr || normal != nullptr), "null pointer");
  DDN_REQUIRE(n_taps >= 0 && n_taps <= kMaxTaps && (n_taps == 0 || (n_taps % 2 == 1 && taps_host != nullptr)), "taps");
  const int H = (int)height, W = (int)width;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((W + 127) / 128, H), block(128);
  const float* smooth = depth;
  if (depth != nullptr && n_taps > 0) {
    const int64_t plane = align_up(height * width * 4, 256);
    DDN_REQUIRE(workspace != nullptr && workspace_bytes >= 2 * plane + 256, "workspace too small (2 float planes)");
    char* base = reinterpret_cast<char*>(align_up((int64_t)(uintptr_t)workspace, 256));
    float* a = reinterpret_cast<float*>(base);
    float* b = reinterpret_cast<float*>(base + plane);
    Taps t;
    t.n = n_taps;
    for (int i = 0; i < n_taps; ++i) t.w[i] = taps_host[i];
    conv_axis_zero_kernel<<<grid, block, 0, st>>>(H, W, 1, t, depth, a);  // kernel_x first (initilizer.py:277)
    DDN_TRY(after_launch("conv_axis_zero_kernel"));
    conv_axis_zero_kernel<<<grid, block, 0, st>>>(H, W, 0, t, a, b);
    DDN_TRY(after_launch("conv_axis_zero_kernel"));
    smooth = b;
  }
  gradient_mask_kernel<<<grid, block, 0, st>>>(H, W, smooth, normal, depth_threshold, normal_threshold, mask_out);
  return after_launch("gradient_mask_kernel");
}

int ddn_transform_normals(int64_t n_points, const float* normal_cam, const double* cam_from_world_host, int32_t row_stride,
                          double* normal_world, void* stream) {
  using namespace ddn;
  DDN_REQUIRE(n_points >= 0, "n_points");
  if (n_points == 0) return DDN_OK;
  DDN_REQUIRE(normal_cam && cam_from_world_host && normal_world && row_stride >= 3, "null pointer / row stride");
  Rot R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R.r[i * 3 + j] = cam_from_world_host[i * row_stride + j];
  transform_normals_kernel<<<(unsigned)((n_points + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_points, R, normal_cam,
                                                                                                 normal_world);
  return after_launch("transform_normals_kernel");
}

}  // extern "C"
