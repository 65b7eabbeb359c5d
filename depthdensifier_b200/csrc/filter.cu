// Stages 2+3: fused pixel back-projection and multi-view consistency vote (kernel K4), plus the
// float64 set-up kernel that turns poses into per-(source view, neighbour) float32 tables.
//
// Reference semantics (see include/ddn_b200.h): scripts/test.py:79-90, 205-233 (back-projection),
// scripts/test.py:58-76, 273-330 (reproject, grazing gate, truncating lookup, one-sided floater vote).
//
// HBM-bound design (no dense contraction anywhere on this path, so no tensor cores):
//   * one CTA owns 256*PX consecutive pixels of ONE source view, so the K neighbour tables are
//     CTA-uniform and live in shared memory;
//   * normals in / xyz out are AoS float3: they move through a shared staging buffer as float4
//     vectors (fully coalesced 16 B per lane) and are read/written per pixel with stride-3 LDS/STS
//     (conflict free);
//   * neighbour depth gathers go through the read-only path; consecutive lanes hit consecutive
//     pixels of the neighbour map, and CTAs are scheduled view by view so the K neighbour maps of
//     the views in flight stay L2 resident;
//   * float->int truncation uses FADD.RZ with 2^23 instead of F2I (keeps the XU pipe free for the
//     one MUFU.RCP and one MUFU.SQRT per pair).
#include "common.cuh"

namespace ddn {

constexpr int kFilterThreads = 256;
constexpr int kFilterPX = 4;  // pixels per thread
constexpr int kFilterChunk = kFilterThreads * kFilterPX;

// ------------------------------------------------------------------------------------------------
// Table set-up (float64, one thread per (source view, neighbour))
// pair_table[s][k][24]: rows 0..2 of M = R_t R_s^T Kinv_s with t_ts appended (12 floats),
//   camera centre of t (3), target view index (int bits), fx_t fy_t cx_t cy_t, own-view flag, pad.
// src_table[s][16]: rows of R_s^T Kinv_s with c_s appended (12 floats), cx_s, cy_s, pad.
// ------------------------------------------------------------------------------------------------
__global__ void build_pair_tables_kernel(int n_total, int src_begin, int n_src, int k_nbr,
                                         const double* __restrict__ poses,
                                         const double* __restrict__ intr,
                                         const int32_t* __restrict__ nbr, float* __restrict__ pair_table,
                                         float* __restrict__ src_table) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int n_pairs = n_src * k_nbr;
  if (idx >= n_pairs + n_src) return;
  const bool is_src_row = idx >= n_pairs;
  const int sl = is_src_row ? idx - n_pairs : idx / k_nbr;
  const int s = src_begin + sl;
  const double* Ps = poses + (size_t)s * 12;
  const double fx = intr[s * 4 + 0], fy = intr[s * 4 + 1], cx = intr[s * 4 + 2], cy = intr[s * 4 + 3];
  // Kinv_s columns: (1/fx, 0, 0), (0, 1/fy, 0), (-cx/fx, -cy/fy, 1)
  double Rs[3][3], ts[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Rs[i][j] = Ps[i * 4 + j];
    ts[i] = Ps[i * 4 + 3];
  }
  if (is_src_row) {
    float* o = src_table + (size_t)sl * 16;
    for (int i = 0; i < 3; ++i) {
      // row i of R_s^T: (Rs[0][i], Rs[1][i], Rs[2][i]);  c_s = -R_s^T t_s
      double r0 = Rs[0][i], r1 = Rs[1][i], r2 = Rs[2][i];
      o[i * 4 + 0] = (float)(r0 / fx);
      o[i * 4 + 1] = (float)(r1 / fy);
      o[i * 4 + 2] = (float)(-r0 * cx / fx - r1 * cy / fy + r2);
      o[i * 4 + 3] = (float)(-(r0 * ts[0] + r1 * ts[1] + r2 * ts[2]));
    }
    o[12] = (float)cx;
    o[13] = (float)cy;
    o[14] = (float)fx;
    o[15] = (float)fy;
    return;
  }
  const int k = idx - sl * k_nbr;
  const int t = nbr[(size_t)s * k_nbr + k];
  float* o = pair_table + (size_t)idx * DDN_PAIR_TABLE_FLOATS;
  if (t < 0 || t >= n_total) {
    for (int i = 0; i < DDN_PAIR_TABLE_FLOATS; ++i) o[i] = 0.f;
    o[15] = __int_as_float(-1);
    return;
  }
  const double* Pt = poses + (size_t)t * 12;
  double Rt[3][3], tt[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Rt[i][j] = Pt[i * 4 + j];
    tt[i] = Pt[i * 4 + 3];
  }
  for (int i = 0; i < 3; ++i) {
    double r[3];  // row i of R_t R_s^T
    for (int j = 0; j < 3; ++j) r[j] = Rt[i][0] * Rs[j][0] + Rt[i][1] * Rs[j][1] + Rt[i][2] * Rs[j][2];
    o[i * 4 + 0] = (float)(r[0] / fx);
    o[i * 4 + 1] = (float)(r[1] / fy);
    o[i * 4 + 2] = (float)(-r[0] * cx / fx - r[1] * cy / fy + r[2]);
    o[i * 4 + 3] = (float)(tt[i] - (r[0] * ts[0] + r[1] * ts[1] + r[2] * ts[2]));
  }
  for (int i = 0; i < 3; ++i)  // c_t = -R_t^T t_t
    o[12 + i] = (float)(-(Rt[0][i] * tt[0] + Rt[1][i] * tt[1] + Rt[2][i] * tt[2]));
  o[15] = __int_as_float(t);
  o[16] = (float)intr[t * 4 + 0];
  o[17] = (float)intr[t * 4 + 1];
  o[18] = (float)intr[t * 4 + 2];
  o[19] = (float)intr[t * 4 + 3];
  o[20] = (t == s) ? 1.f : 0.f;
  o[21] = o[22] = o[23] = 0.f;
}

// ------------------------------------------------------------------------------------------------
// K4
// ------------------------------------------------------------------------------------------------
struct FilterParams {
  const float* refined_all;  // [V,H,W]
  const float* normal;       // [n_src,H,W,3]
  const float* pair_table;   // [n_src,K,24]
  const float* src_table;    // [n_src,16]
  float* xyz;                // [n_src,Hs,Ws,3]
  uint8_t* votes;            // [n_src,Hs,Ws]
  int* bbox;                 // [6] ordered-int encoded, or nullptr
  int src_begin, n_src, H, W, Hs, Ws, K, stride;
  int vote_threshold;
  float depth_threshold, grazing_cos, two_sided_tau;
  int normals_in_world;
};

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// trunc(u) for 0 <= u < 2^22, biased by 0x4B000000, on the FMA pipe (no F2I).
__device__ __forceinline__ int trunc_biased(float u) { return __float_as_int(__fadd_rz(u, 8388608.0f)); }
constexpr int kTruncBias = 0x4B000000;

// Cooperative global<->shared copy of n floats, float4 when the global address is 16 B aligned.
template <bool kLoad>
__device__ __forceinline__ void stage_floats(float* smem, float* gptr, int n) {
  const int tid = threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(gptr) & 15) == 0) {
    const int n4 = n >> 2;
    float4* g4 = reinterpret_cast<float4*>(gptr);
    float4* s4 = reinterpret_cast<float4*>(smem);
    for (int i = tid; i < n4; i += kFilterThreads) {
      if (kLoad)
        s4[i] = __ldcs(g4 + i);
      else
        __stcs(g4 + i, s4[i]);
    }
    for (int i = (n4 << 2) + tid; i < n; i += kFilterThreads) {
      if (kLoad)
        smem[i] = __ldcs(gptr + i);
      else
        __stcs(gptr + i, smem[i]);
    }
  } else {
    for (int i = tid; i < n; i += kFilterThreads) {
      if (kLoad)
        smem[i] = __ldcs(gptr + i);
      else
        __stcs(gptr + i, smem[i]);
    }
  }
}

template <bool kBilinear, bool kStride1>
__global__ void __launch_bounds__(kFilterThreads, 3) backproject_filter_kernel(const FilterParams p) {
  extern __shared__ __align__(16) float smem[];
  float* s_stage = smem;                          // kFilterChunk*3 floats
  float* s_src = smem + kFilterChunk * 3;         // 16 floats
  float* s_pair = s_src + 16;                     // K*24 floats
  __shared__ int s_bbox[6];

  const int tid = threadIdx.x;
  const int sl = blockIdx.y;
  const int s = p.src_begin + sl;
  const int Ps = p.Hs * p.Ws;  // source-grid pixels per view
  const int chunk0 = blockIdx.x * kFilterChunk;
  const int n_here = min(kFilterChunk, Ps - chunk0);
  const size_t HW = (size_t)p.H * p.W;
  const float* __restrict__ depth_s = p.refined_all + (size_t)s * HW;

  for (int i = tid; i < p.K * DDN_PAIR_TABLE_FLOATS; i += kFilterThreads)
    s_pair[i] = __ldg(p.pair_table + (size_t)sl * p.K * DDN_PAIR_TABLE_FLOATS + i);
  if (tid < 16) s_src[tid] = __ldg(p.src_table + (size_t)sl * 16 + tid);
  if (tid < 6) s_bbox[tid] = tid < 3 ? 0x7fffffff : (int)0x80000000;
  if (kStride1) {
    stage_floats<true>(s_stage, const_cast<float*>(p.normal) + ((size_t)sl * HW + chunk0) * 3, n_here * 3);
  }
  __syncthreads();

  float d[kFilterPX], P[kFilterPX], Q[kFilterPX], nx[kFilterPX], ny[kFilterPX], nz[kFilterPX], nXw[kFilterPX];
  int px[kFilterPX], py[kFilterPX];
  int nvotes[kFilterPX];
  float bmin[3] = {INFINITY, INFINITY, INFINITY}, bmax[3] = {-INFINITY, -INFINITY, -INFINITY};
  float kx[kFilterPX], ky[kFilterPX], kz[kFilterPX];

#pragma unroll
  for (int j = 0; j < kFilterPX; ++j) {
    const int l = j * kFilterThreads + tid;
    const int pix = chunk0 + l;
    const bool in = l < n_here;
    int x = 0, y = 0;
    float dd = 0.f;
    float n0 = 0.f, n1 = 0.f, n2 = 0.f;
    if (in) {
      const int ys = pix / p.Ws;
      const int xs = pix - ys * p.Ws;
      if (kStride1) {
        x = xs;
        y = ys;
        dd = __ldg(depth_s + pix);
        n0 = s_stage[l * 3 + 0];
        n1 = s_stage[l * 3 + 1];
        n2 = s_stage[l * 3 + 2];
      } else {
        x = xs * p.stride;
        y = ys * p.stride;
        const size_t g = (size_t)y * p.W + x;
        dd = __ldg(depth_s + g);
        const float* np_ = p.normal + ((size_t)sl * HW + g) * 3;
        n0 = __ldg(np_ + 0);
        n1 = __ldg(np_ + 1);
        n2 = __ldg(np_ + 2);
      }
    }
    const bool valid = in && dd > 0.f;
    d[j] = valid ? dd : 0.f;
    px[j] = x;
    py[j] = y;
    P[j] = d[j] * (float)x;
    Q[j] = d[j] * (float)y;
    // world position (scripts/test.py:79-90 then :233), fp32 with float64-precomputed rows
    const float X = fmaf(s_src[0], P[j], fmaf(s_src[1], Q[j], fmaf(s_src[2], d[j], s_src[3])));
    const float Y = fmaf(s_src[4], P[j], fmaf(s_src[5], Q[j], fmaf(s_src[6], d[j], s_src[7])));
    const float Z = fmaf(s_src[8], P[j], fmaf(s_src[9], Q[j], fmaf(s_src[10], d[j], s_src[11])));
    kx[j] = X;
    ky[j] = Y;
    kz[j] = Z;
    if (p.normals_in_world) {
      // n_w = R_s^T n_c; rows of R_s^T are (src[0]*fx, src[1]*fy, ...) - recover from the table
      const float fx = s_src[14], fy = s_src[15], cx = s_src[12], cy = s_src[13];
      float r[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        r[i][0] = s_src[i * 4 + 0] * fx;
        r[i][1] = s_src[i * 4 + 1] * fy;
        r[i][2] = s_src[i * 4 + 2] + r[i][0] * cx / fx + r[i][1] * cy / fy;
      }
      const float w0 = r[0][0] * n0 + r[0][1] * n1 + r[0][2] * n2;
      const float w1 = r[1][0] * n0 + r[1][1] * n1 + r[1][2] * n2;
      const float w2 = r[2][0] * n0 + r[2][1] * n1 + r[2][2] * n2;
      n0 = w0;
      n1 = w1;
      n2 = w2;
    }
    nx[j] = n0;
    ny[j] = n1;
    nz[j] = n2;
    nXw[j] = fmaf(n0, X, fmaf(n1, Y, n2 * Z));
    nvotes[j] = valid ? 0 : 255;
    if (in) {
      if (kStride1) {
        s_stage[l * 3 + 0] = valid ? X : 0.f;
        s_stage[l * 3 + 1] = valid ? Y : 0.f;
        s_stage[l * 3 + 2] = valid ? Z : 0.f;
      } else {
        float* o = p.xyz + ((size_t)sl * Ps + pix) * 3;
        o[0] = valid ? X : 0.f;
        o[1] = valid ? Y : 0.f;
        o[2] = valid ? Z : 0.f;
      }
    }
  }

  const float Wf = (float)p.W, Hf = (float)p.H;
  const float thr = p.depth_threshold;
  const float gcos = p.grazing_cos;
  const float tau = p.two_sided_tau;

  for (int k = 0; k < p.K; ++k) {
    const float4* t4 = reinterpret_cast<const float4*>(s_pair + k * DDN_PAIR_TABLE_FLOATS);
    const float4 r0 = t4[0], r1 = t4[1], r2 = t4[2], cc = t4[3], kk = t4[4];
    const int t = __float_as_int(cc.w);
    if (t < 0) continue;
    const bool own = s_pair[k * DDN_PAIR_TABLE_FLOATS + 20] != 0.f;
    const float* __restrict__ depth_t = p.refined_all + (size_t)t * HW;
#pragma unroll
    for (int j = 0; j < kFilterPX; ++j) {
      const float X = fmaf(r0.x, P[j], fmaf(r0.y, Q[j], fmaf(r0.z, d[j], r0.w)));
      const float Y = fmaf(r1.x, P[j], fmaf(r1.y, Q[j], fmaf(r1.z, d[j], r1.w)));
      const float Z = fmaf(r2.x, P[j], fmaf(r2.y, Q[j], fmaf(r2.z, d[j], r2.w)));
      // grazing gate (scripts/test.py:284-295): dot(n, -(Xw - c_t)/|Xw - c_t|) > cos
      const float nc = fmaf(nx[j], cc.x, fmaf(ny[j], cc.y, nz[j] * cc.z));
      const float dn = nc - nXw[j];
      const float len = sqrt_approx(fmaf(X, X, fmaf(Y, Y, Z * Z)));
      bool ok = (d[j] > 0.f) && (dn > gcos * len);
      float zq = Z;
      float D = 0.f;
      if (own) {
        // Own view: the reference normalises by (z + 1e-8) before applying K (scripts/test.py:71-75), so
        // u = x * z/(z+1e-8) lands ~x*1e-8/z BELOW the integer x (far above float64 round-off) and the
        // truncation at :308-309 looks up pixel (x-1, y-1) for x, y >= 1.  Reproduced in integer
        // arithmetic; z is the pixel's own depth.  (x == 0 or y == 0 is a round-off tie in the reference.)
        const int ux = max(px[j] - 1, 0);
        const int vy = max(py[j] - 1, 0);
        zq = d[j];
        if (ok) D = __ldg(depth_t + (size_t)vy * p.W + ux);
      } else {
        const float inv = rcp_approx(Z);
        const float u = fmaf(kk.x, X * inv, kk.z);
        const float v = fmaf(kk.y, Y * inv, kk.w);
        ok = ok && (Z > 0.f) && (u >= 0.f) && (u < Wf) && (v >= 0.f) && (v < Hf);
        if (!kBilinear) {
          if (ok) {
            const int ui = trunc_biased(u) - kTruncBias;
            const int vi = trunc_biased(v) - kTruncBias;
            D = __ldg(depth_t + vi * p.W + ui);
          }
        } else {
          if (ok) {
            // N3: 4 taps at floor(u), floor(v), +1 clamped; all taps must be > 0
            const int x0 = trunc_biased(u) - kTruncBias;
            const int y0 = trunc_biased(v) - kTruncBias;
            const int x1 = min(x0 + 1, p.W - 1);
            const int y1 = min(y0 + 1, p.H - 1);
            const float fxw = u - (float)x0, fyw = v - (float)y0;
            const float ta = __ldg(depth_t + y0 * p.W + x0);
            const float tb = __ldg(depth_t + y0 * p.W + x1);
            const float tc = __ldg(depth_t + y1 * p.W + x0);
            const float td = __ldg(depth_t + y1 * p.W + x1);
            const bool all = (ta > 0.f) && (tb > 0.f) && (tc > 0.f) && (td > 0.f);
            const float gx = 1.f - fxw, gy = 1.f - fyw;
            const float acc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(gx, gy), ta), __fmul_rn(__fmul_rn(fxw, gy), tb)),
                                                  __fmul_rn(__fmul_rn(gx, fyw), tc)),
                                        __fmul_rn(__fmul_rn(fxw, fyw), td));
            D = all ? acc : 0.f;
          }
        }
      }
      bool bad;
      if (tau > 0.f)
        bad = fabsf(zq - D) > tau * D;
      else
        bad = zq < __fmul_rn(thr, D);  // float32 product, NEP-50 (scripts/test.py:320)
      nvotes[j] += (ok && D > 0.f && bad) ? 1 : 0;
    }
  }

#pragma unroll
  for (int j = 0; j < kFilterPX; ++j) {
    const int l = j * kFilterThreads + tid;
    if (l < n_here) {
      p.votes[(size_t)sl * Ps + chunk0 + l] = (uint8_t)min(nvotes[j], 255);
      if (nvotes[j] < p.vote_threshold) {
        bmin[0] = fminf(bmin[0], kx[j]);
        bmin[1] = fminf(bmin[1], ky[j]);
        bmin[2] = fminf(bmin[2], kz[j]);
        bmax[0] = fmaxf(bmax[0], kx[j]);
        bmax[1] = fmaxf(bmax[1], ky[j]);
        bmax[2] = fmaxf(bmax[2], kz[j]);
      }
    }
  }
  if (p.bbox != nullptr) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        bmin[i] = fminf(bmin[i], __shfl_xor_sync(0xffffffffu, bmin[i], o));
        bmax[i] = fmaxf(bmax[i], __shfl_xor_sync(0xffffffffu, bmax[i], o));
      }
    }
    if ((tid & 31) == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (bmin[i] <= bmax[i]) {
          atomicMin(&s_bbox[i], float_to_ordered(bmin[i]));
          atomicMax(&s_bbox[3 + i], float_to_ordered(bmax[i]));
        }
      }
    }
  }
  __syncthreads();
  if (kStride1) stage_floats<false>(s_stage, p.xyz + ((size_t)sl * Ps + chunk0) * 3, n_here * 3);
  if (p.bbox != nullptr && tid < 6) {
    if (tid < 3) {
      if (s_bbox[tid] != 0x7fffffff) atomicMin(p.bbox + tid, s_bbox[tid]);
    } else {
      if (s_bbox[tid] != (int)0x80000000) atomicMax(p.bbox + tid, s_bbox[tid]);
    }
  }
}

__global__ void bbox_init_kernel(int* bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = float_to_ordered(INFINITY);
  else if (threadIdx.x < 6) bbox[threadIdx.x] = float_to_ordered(-INFINITY);
}

}  // namespace ddn

extern "C" {

int ddn_build_pair_tables(int64_t n_views_total, int64_t src_begin, int64_t n_src, int64_t k_nbr,
                          const double* cam_from_world, const double* intr, const int32_t* nbr,
                          float* pair_table, float* src_table, void* stream) {
  using namespace ddn;
  DDN_REQUIRE(n_views_total > 0 && n_src >= 0 && k_nbr > 0, "view counts");
  DDN_REQUIRE(src_begin >= 0 && src_begin + n_src <= n_views_total, "source range");
  DDN_REQUIRE(cam_from_world && intr && nbr && pair_table && src_table, "null pointer");
  if (n_src == 0) return DDN_OK;
  const int total = (int)(n_src * k_nbr + n_src);
  build_pair_tables_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      (int)n_views_total, (int)src_begin, (int)n_src, (int)k_nbr, cam_from_world, intr, nbr, pair_table, src_table);
  return after_launch("build_pair_tables_kernel");
}

int ddn_bbox_init(float* bbox, void* stream) {
  using namespace ddn;
  DDN_REQUIRE(bbox != nullptr, "null bbox");
  bbox_init_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<int*>(bbox));
  return after_launch("bbox_init_kernel");
}

int ddn_backproject_filter(const ddn_filter_config* cfg, int64_t n_views_total, int64_t src_begin,
                           int64_t n_src, int64_t height, int64_t width, int64_t k_nbr,
                           const float* refined_all, const float* normal, const int32_t* nbr,
                           const float* pair_table, const float* src_table, int32_t vote_threshold,
                           float* xyz, uint8_t* votes, float* bbox, void* stream) {
  using namespace ddn;
  (void)nbr;
  DDN_REQUIRE(cfg != nullptr, "null config");
  DDN_REQUIRE(n_views_total > 0 && n_src >= 0 && k_nbr > 0 && k_nbr <= 1024, "view counts");
  DDN_REQUIRE(src_begin >= 0 && src_begin + n_src <= n_views_total, "source range");
  DDN_REQUIRE(height > 0 && width > 0 && height * width < (1ll << 31), "image size");
  DDN_REQUIRE(width < (1 << 22) && height < (1 << 22), "image side too large for the truncation trick");
  DDN_REQUIRE(cfg->stride >= 1, "stride");
  DDN_REQUIRE(refined_all && normal && pair_table && src_table && xyz && votes, "null pointer");
  if (n_src == 0) return DDN_OK;
  FilterParams p;
  p.refined_all = refined_all;
  p.normal = normal;
  p.pair_table = pair_table;
  p.src_table = src_table;
  p.xyz = xyz;
  p.votes = votes;
  p.bbox = reinterpret_cast<int*>(bbox);
  p.src_begin = (int)src_begin;
  p.n_src = (int)n_src;
  p.H = (int)height;
  p.W = (int)width;
  p.stride = cfg->stride;
  p.Hs = (p.H + p.stride - 1) / p.stride;
  p.Ws = (p.W + p.stride - 1) / p.stride;
  p.K = (int)k_nbr;
  p.vote_threshold = vote_threshold;
  p.depth_threshold = cfg->depth_threshold;
  p.grazing_cos = cfg->grazing_cos;
  p.two_sided_tau = cfg->two_sided_tau;
  p.normals_in_world = cfg->normals_in_world;
  const int Ps = p.Hs * p.Ws;
  dim3 grid((Ps + kFilterChunk - 1) / kFilterChunk, (unsigned)n_src);
  DDN_REQUIRE(n_src <= 65535, "too many source views per call");
  const size_t smem = (size_t)(kFilterChunk * 3 + 16 + p.K * DDN_PAIR_TABLE_FLOATS) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  const bool bil = cfg->sample_mode == 1;
  const bool s1 = cfg->stride == 1;
#define DDN_LAUNCH_FILTER(B, S)                                                                            \
  do {                                                                                                     \
    DDN_TRY(check_cuda(cudaFuncSetAttribute(backproject_filter_kernel<B, S>,                               \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),       \
                       "cudaFuncSetAttribute"));                                                           \
    backproject_filter_kernel<B, S><<<grid, kFilterThreads, smem, st>>>(p);                                \
  } while (0)
  if (bil && s1) DDN_LAUNCH_FILTER(true, true);
  else if (bil) DDN_LAUNCH_FILTER(true, false);
  else if (s1) DDN_LAUNCH_FILTER(false, true);
  else DDN_LAUNCH_FILTER(false, false);
#undef DDN_LAUNCH_FILTER
  return after_launch("backproject_filter_kernel");
}

}  // extern "C"
