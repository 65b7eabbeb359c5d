// Per-pixel kernels of the reference's alternative refiner, FastPCHIPRefiner
// (src/depthdensifier/fast_pchip_refiner.py):
//
//   edge mask   :187-273  Gaussian smoothing (scipy.ndimage.gaussian_filter: separable, 'reflect' boundary,
//                         double accumulation with the symmetric taps added first, result rounded to the
//                         input type after EACH axis), np.gradient magnitudes, thresholds, and two
//                         dilations with the cross structuring element (= one diamond of radius 2)
//   apply       :300-385, :550-579  cubic-Hermite remap with 0.3-scaled secant tangents in float32 with the
//                         reference's exact operation order, 70/30 blend on edge pixels, 3x3 median,
//                         zero outside the mask
//
// The O(C) correspondence logic of that class (project the sparse points, MAD outlier rejection, np.unique)
// stays on the host in depthdensifier_b200/fast_pchip_refiner.py, as it does in the reference.
// All float arithmetic uses explicit round-to-nearest intrinsics: no FMA contraction, so the results are
// bit-identical to numpy / torch-CPU (tests/test_gpu_pchip.py).
#include "common.cuh"

namespace ddn {

constexpr int kMaxRadius = 64;

struct GaussWeights {
  double w[kMaxRadius + 1];  // w[0] = centre, w[k] = tap at distance k
  int radius;
};

__device__ __forceinline__ int reflect_index(int i, int n) {  // scipy 'reflect': d c b a | a b c d | d c b a
  while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
  return i;
}

// one axis of scipy.ndimage.correlate1d with a symmetric kernel: tmp = x[c] w[c]; tmp += (x[c-k] + x[c+k]) w[k]
// for k = radius .. 1 (the order scipy's loop runs), in double; the result is rounded to T.
template <typename T>
__global__ void gauss_axis_kernel(int H, int W, int axis, GaussWeights gw, const T* __restrict__ in, T* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int n = axis == 0 ? H : W, c = axis == 0 ? y : x;
  auto at = [&](int i) -> double {
    const int r = reflect_index(i, n);
    return (double)(axis == 0 ? in[(size_t)r * W + x] : in[(size_t)y * W + r]);
  };
  double tmp = __dmul_rn(at(c), gw.w[0]);
  for (int k = gw.radius; k >= 1; --k) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(at(c - k), at(c + k)), gw.w[k]));
  out[(size_t)y * W + x] = (T)tmp;
}

// np.gradient along one axis, edge_order 1: central differences / 2 inside, one-sided at the borders
template <typename T>
__device__ __forceinline__ T grad1(const T* __restrict__ f, int i, int n, size_t stride) {
  if (n == 1) return (T)0;
  if (i == 0) return f[stride] - f[0];
  if (i == n - 1) return f[(size_t)i * stride] - f[(size_t)(i - 1) * stride];
  return (f[(size_t)(i + 1) * stride] - f[(size_t)(i - 1) * stride]) / (T)2;
}

// depth mode (:226-268): |grad n| > 0.3 over the three normal channels, OR relative gradient of the smoothed
// depth > threshold (zero outside the mask)
__global__ void depth_edge_kernel(int H, int W, const float* __restrict__ smooth, const uint8_t* __restrict__ mask,
                                  const float* __restrict__ normal, float edge_threshold, uint8_t* __restrict__ edge) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const size_t p = (size_t)y * W + x;
  bool e = false;
  if (normal != nullptr) {
    float acc = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float g0 = grad1(normal + (size_t)x * 3 + ch, y, H, (size_t)W * 3);  // axis 0
      const float g1 = grad1(normal + (size_t)y * W * 3 + ch, x, W, 3);          // axis 1
      acc = ch == 0 ? __fmul_rn(g0, g0) : __fadd_rn(acc, __fmul_rn(g0, g0));
      acc = __fadd_rn(acc, __fmul_rn(g1, g1));
    }
    e = __fsqrt_rn(acc) > 0.3f;
  }
  const float dy = grad1(smooth + x, y, H, (size_t)W), dx = grad1(smooth + (size_t)y * W, x, W, 1);
  const float gmag = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
  float rel = __fdiv_rn(gmag, __fadd_rn(smooth[p], 1e-6f));
  if (mask != nullptr && mask[p] == 0) rel = 0.f;
  e |= rel > edge_threshold;  // NaN compares false, as in numpy
  edge[p] = e ? 1 : 0;
}

// image mode (:187-219): float64 gradient magnitude of the smoothed grey image > threshold, AND mask
__global__ void image_edge_kernel(int H, int W, const double* __restrict__ smooth, const uint8_t* __restrict__ mask,
                                  double threshold, uint8_t* __restrict__ edge) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const size_t p = (size_t)y * W + x;
  const double dy = grad1(smooth + x, y, H, (size_t)W), dx = grad1(smooth + (size_t)y * W, x, W, 1);
  bool e = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) > threshold;
  if (mask != nullptr) e = e && mask[p] != 0;
  edge[p] = e ? 1 : 0;
}

// two dilations with the 4-connected cross = one dilation with the diamond |dx| + |dy| <= 2, border value 0
__global__ void dilate2_kernel(int H, int W, const uint8_t* __restrict__ in, uint8_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  bool e = false;
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy) {
    const int yy = y + dy, span = 2 - (dy < 0 ? -dy : dy);
    if (yy < 0 || yy >= H) continue;
    for (int dx = -span; dx <= span; ++dx) {
      const int xx = x + dx;
      if (xx >= 0 && xx < W) e |= in[(size_t)yy * W + xx] != 0;
    }
  }
  out[(size_t)y * W + x] = e ? 1 : 0;
}

// ---- apply ---------------------------------------------------------------------------------------------
// first / last pixel (row-major) of the two query sets of :556-571: non-edge and edge pixels inside the mask
__global__ void pchip_endpoints_kernel(int64_t n, const uint8_t* __restrict__ mask, const uint8_t* __restrict__ edge,
                                       int* __restrict__ ends /* [4]: first/last non-edge, first/last edge */) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int lo[2] = {0x7fffffff, 0x7fffffff}, hi[2] = {-1, -1};
  if (i < n && mask[i]) {
    const int s = edge[i] ? 1 : 0;
    lo[s] = hi[s] = (int)i;
  }
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int l = __reduce_min_sync(0xffffffffu, lo[s]), h = __reduce_max_sync(0xffffffffu, hi[s]);
    if ((threadIdx.x & 31) == 0) {
      if (l != 0x7fffffff) atomicMin(ends + 2 * s, l);
      if (h >= 0) atomicMax(ends + 2 * s + 1, h);
    }
  }
}

__global__ void pchip_init_ends_kernel(int* ends) {
  if (threadIdx.x < 4) ends[threadIdx.x] = (threadIdx.x & 1) ? -1 : 0x7fffffff;
}

struct Knots {
  const float* x;
  const float* y;
  int n;
};

// slopes_padded[i] of :316-321: secant slope of interval i-1 for 1 <= i <= n-1, the first / last slope at 0 / n
__device__ __forceinline__ float padded_slope(const Knots& k, int i) {
  const int j = min(max(i - 1, 0), k.n - 2);
  return __fdiv_rn(__fadd_rn(k.y[j + 1], -k.y[j]), __fadd_rn(__fadd_rn(k.x[j + 1], -k.x[j]), 1e-8f));
}

__device__ __forceinline__ int interval_of(const Knots& k, float d) {  // clamp(searchsorted_left(x, d), 1, n-1)
  int lo = 0, len = k.n;
  while (len > 0) {
    const int half = len >> 1;
    const bool less = k.x[lo + half] < d;
    lo = less ? lo + half + 1 : lo;
    len = less ? len - half - 1 : half;
  }
  return min(max(lo, 1), k.n - 1);
}

// :323-363 in the reference's operation order (every product and sum rounded to float32)
__device__ __forceinline__ float hermite_eval(const Knots& k, float d, float sl_first, float sr_last) {
  const int i = interval_of(k, d);
  const float xl = k.x[i - 1], xr = k.x[i], yl = k.y[i - 1], yr = k.y[i];
  const float sl = padded_slope(k, i - 1), sr = padded_slope(k, i);
  const float h = __fadd_rn(__fadd_rn(xr, -xl), 1e-8f);
  const float t = __fdiv_rn(__fadd_rn(d, -xl), h);
  const float t2 = __fmul_rn(t, t), t3 = __fmul_rn(t2, t);
  const float h00 = __fadd_rn(__fadd_rn(__fmul_rn(2.f, t3), -__fmul_rn(3.f, t2)), 1.f);
  const float h10 = __fadd_rn(__fadd_rn(t3, -__fmul_rn(2.f, t2)), t);
  const float h01 = __fadd_rn(__fmul_rn(-2.f, t3), __fmul_rn(3.f, t2));
  const float h11 = __fadd_rn(t3, -t2);
  float r = __fmul_rn(h00, yl);
  r = __fadd_rn(r, __fmul_rn(__fmul_rn(__fmul_rn(h10, h), sl), 0.3f));
  r = __fadd_rn(r, __fmul_rn(h01, yr));
  r = __fadd_rn(r, __fmul_rn(__fmul_rn(__fmul_rn(h11, h), sr), 0.3f));
  const float x0 = k.x[0], xn = k.x[k.n - 1];
  if (d <= x0) r = __fadd_rn(k.y[0], __fmul_rn(__fmul_rn(sl_first, __fadd_rn(d, -x0)), 0.3f));
  if (d >= xn) r = __fadd_rn(k.y[k.n - 1], __fmul_rn(__fmul_rn(sr_last, __fadd_rn(d, -xn)), 0.3f));
  return fmaxf(r, 1e-3f);
}

constexpr int kPTileW = 64, kPTileH = 32, kPThreads = 256;
constexpr int kPHaloW = kPTileW + 2, kPHaloH = kPTileH + 2;

__device__ __forceinline__ void cswap(float& a, float& b) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  a = lo;
  b = hi;
}
__device__ __forceinline__ float median9(float p0, float p1, float p2, float p3, float p4, float p5, float p6, float p7, float p8) {
  cswap(p1, p2); cswap(p4, p5); cswap(p7, p8);
  cswap(p0, p1); cswap(p3, p4); cswap(p6, p7);
  cswap(p1, p2); cswap(p4, p5); cswap(p7, p8);
  cswap(p0, p3); cswap(p5, p8); cswap(p4, p7);
  cswap(p3, p6); cswap(p1, p4); cswap(p2, p5);
  cswap(p4, p7); cswap(p4, p2); cswap(p6, p4);
  cswap(p4, p2);
  return p4;
}

// One CTA = a 64 x 32 tile + halo: transformed values into shared memory, then the 3x3 median (scipy's default
// 'reflect' boundary duplicates the border pixel for a 3-wide window, i.e. clamped indices).
__global__ void __launch_bounds__(kPThreads)
pchip_apply_kernel(int H, int W, int tiles_x, const float* __restrict__ depth, const uint8_t* __restrict__ mask,
                   const uint8_t* __restrict__ edge, Knots gk, const int* __restrict__ ends, float* __restrict__ refined) {
  extern __shared__ __align__(16) float s_knots[];  // x | y (when they fit)
  __shared__ float s_val[kPHaloH][kPHaloW];
  __shared__ float s_tan[4];  // sl_first / sr_last of the non-edge and of the edge query set
  const int tid = threadIdx.x;
  const int ty0 = (blockIdx.x / tiles_x) * kPTileH, tx0 = (blockIdx.x % tiles_x) * kPTileW;
  Knots k = gk;
  if (gk.n * 2 * sizeof(float) <= 64 * 1024) {
    for (int i = tid; i < gk.n; i += kPThreads) {
      s_knots[i] = gk.x[i];
      s_knots[gk.n + i] = gk.y[i];
    }
    k.x = s_knots;
    k.y = s_knots + gk.n;
  }
  __syncthreads();
  if (tid < 4) {
    // tangents the reference takes from the first / last QUERY of each call (:352-358)
    const int pix = ends[tid];
    float s = 0.f;
    if (pix >= 0 && pix != 0x7fffffff) {
      const int i = interval_of(k, depth[pix]);
      s = (tid & 1) ? padded_slope(k, i) : padded_slope(k, i - 1);
    }
    s_tan[tid] = s;
  }
  __syncthreads();
  const bool any_non_edge = ends[1] >= 0;  // :556-559: without non-edge pixels the canvas starts as zeros
  for (int i = tid; i < kPHaloW * kPHaloH; i += kPThreads) {
    const int hy = i / kPHaloW, hx = i - hy * kPHaloW;
    const int y = min(max(ty0 + hy - 1, 0), H - 1), x = min(max(tx0 + hx - 1, 0), W - 1);
    const size_t g = (size_t)y * W + x;
    const float d = depth[g];
    float val = any_non_edge ? d : 0.f;  // pixels outside the mask keep the original depth until the final masking
    if (mask[g]) {
      if (!edge[g]) {
        val = hermite_eval(k, d, s_tan[0], s_tan[1]);
      } else {
        const float tr = hermite_eval(k, d, s_tan[2], s_tan[3]);
        val = __fadd_rn(__fmul_rn(0.7f, d), __fmul_rn(0.3f, tr));
      }
    }
    s_val[hy][hx] = val;
  }
  __syncthreads();
  for (int i = tid; i < kPTileW * kPTileH; i += kPThreads) {
    const int ly = i / kPTileW, lx = i - ly * kPTileW;
    const int y = ty0 + ly, x = tx0 + lx;
    if (y >= H || x >= W) continue;
    const float o = median9(s_val[ly][lx], s_val[ly][lx + 1], s_val[ly][lx + 2], s_val[ly + 1][lx], s_val[ly + 1][lx + 1],
                            s_val[ly + 1][lx + 2], s_val[ly + 2][lx], s_val[ly + 2][lx + 1], s_val[ly + 2][lx + 2]);
    const size_t g = (size_t)y * W + x;
    refined[g] = mask[g] ? o : 0.f;
  }
}

static int gauss_weights_from_host(const double* w, int radius, GaussWeights* g) {
  DDN_REQUIRE(w != nullptr && radius >= 0 && radius <= kMaxRadius, "gaussian radius must be in [0, 64]");
  g->radius = radius;
  for (int i = 0; i <= radius; ++i) g->w[i] = w[i];
  return DDN_OK;
}

}  // namespace ddn

extern "C" {

int ddn_pchip_workspace_bytes(int64_t height, int64_t width, int64_t* bytes_out) {
  using namespace ddn;
  DDN_REQUIRE(bytes_out != nullptr && height > 0 && width > 0, "shape");
  *bytes_out = align_up(height * width * 8, 256) * 2 + align_up(height * width, 256) + 512;
  return DDN_OK;
}

int ddn_pchip_edge_mask(int64_t height, int64_t width, const float* depth, const uint8_t* mask, const float* normal,
                        const double* gray, const double* gauss_weights_host, int32_t radius, float edge_threshold,
                        double image_edge_threshold, uint8_t* edge_out, void* workspace, int64_t workspace_bytes,
                        void* stream) {
  using namespace ddn;
  DDN_REQUIRE(height > 0 && width > 0 && height * width < (1ll << 31), "shape");
  DDN_REQUIRE((depth != nullptr || gray != nullptr) && edge_out != nullptr && workspace != nullptr, "null pointer");
  int64_t need = 0;
  DDN_TRY(ddn_pchip_workspace_bytes(height, width, &need));
  if (workspace_bytes < need) {
    set_error("pchip workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)need);
    return DDN_ERR_WORKSPACE_TOO_SMALL;
  }
  GaussWeights gw;
  DDN_TRY(gauss_weights_from_host(gauss_weights_host, radius, &gw));
  cudaStream_t st = (cudaStream_t)stream;
  const int H = (int)height, W = (int)width;
  char* base = reinterpret_cast<char*>(align_up((int64_t)(uintptr_t)workspace, 256));
  const int64_t plane = align_up(height * width * 8, 256);
  uint8_t* raw = reinterpret_cast<uint8_t*>(base + 2 * plane);
  dim3 grid((W + 127) / 128, H), block(128);
  if (gray != nullptr) {  // image mode: everything in float64
    double* a = reinterpret_cast<double*>(base);
    double* b = reinterpret_cast<double*>(base + plane);
    gauss_axis_kernel<double><<<grid, block, 0, st>>>(H, W, 0, gw, gray, a);
    DDN_TRY(after_launch("gauss_axis_kernel"));
    gauss_axis_kernel<double><<<grid, block, 0, st>>>(H, W, 1, gw, a, b);
    DDN_TRY(after_launch("gauss_axis_kernel"));
    image_edge_kernel<<<grid, block, 0, st>>>(H, W, b, mask, image_edge_threshold, raw);
    DDN_TRY(after_launch("image_edge_kernel"));
  } else {
    float* a = reinterpret_cast<float*>(base);
    float* b = reinterpret_cast<float*>(base + plane);
    gauss_axis_kernel<float><<<grid, block, 0, st>>>(H, W, 0, gw, depth, a);
    DDN_TRY(after_launch("gauss_axis_kernel"));
    gauss_axis_kernel<float><<<grid, block, 0, st>>>(H, W, 1, gw, a, b);
    DDN_TRY(after_launch("gauss_axis_kernel"));
    depth_edge_kernel<<<grid, block, 0, st>>>(H, W, b, mask, normal, edge_threshold, raw);
    DDN_TRY(after_launch("depth_edge_kernel"));
  }
  dilate2_kernel<<<grid, block, 0, st>>>(H, W, raw, edge_out);
  return after_launch("dilate2_kernel");
}

int ddn_pchip_apply(int64_t height, int64_t width, const float* depth, const uint8_t* mask, const uint8_t* edge,
                    const float* knots_x, const float* knots_y, int32_t n_knots, float* refined, void* workspace,
                    int64_t workspace_bytes, void* stream) {
  using namespace ddn;
  DDN_REQUIRE(height > 0 && width > 0 && height * width < (1ll << 31), "shape");
  DDN_REQUIRE(depth && mask && edge && knots_x && knots_y && refined && workspace, "null pointer");
  DDN_REQUIRE(n_knots >= 2, "at least two knots");
  DDN_REQUIRE(workspace_bytes >= 256, "workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const int H = (int)height, W = (int)width;
  int* ends = reinterpret_cast<int*>(align_up((int64_t)(uintptr_t)workspace, 256));
  pchip_init_ends_kernel<<<1, 32, 0, st>>>(ends);
  DDN_TRY(after_launch("pchip_init_ends_kernel"));
  const int64_t n = height * width;
  pchip_endpoints_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, mask, edge, ends);
  DDN_TRY(after_launch("pchip_endpoints_kernel"));
  const int tiles_x = (W + kPTileW - 1) / kPTileW, tiles_y = (H + kPTileH - 1) / kPTileH;
  const size_t smem = (size_t)n_knots * 8 <= 64 * 1024 ? (size_t)n_knots * 8 : 0;
  DDN_TRY(check_cuda(cudaFuncSetAttribute(pchip_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                     "cudaFuncSetAttribute(pchip_apply)"));
  Knots k{knots_x, knots_y, n_knots};
  pchip_apply_kernel<<<tiles_x * tiles_y, kPThreads, smem, st>>>(H, W, tiles_x, depth, mask, edge, k, ends, refined);
  return after_launch("pchip_apply_kernel");
}

}  // extern "C"
