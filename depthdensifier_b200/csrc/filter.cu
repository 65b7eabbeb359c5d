// Stages 2+3: fused pixel back-projection and multi-view consistency vote (kernel K4), plus the
// float64 set-up kernel that turns poses into per-(source view, neighbour) float32 tables.
//
// Reference semantics (see include/ddn_b200.h): scripts/test.py:79-90, 205-233 (back-projection),
// scripts/test.py:58-76, 273-330 (reproject, grazing gate, truncating lookup, one-sided floater vote).
//
// HBM-bound design (no dense contraction anywhere on this path, so no tensor cores):
//   * one CTA owns 256*PX consecutive pixels of ONE source view, so the K neighbour tables are
//     CTA-uniform and live in shared memory;
//   * xyz out is AoS float3: it moves through a shared staging buffer as float4 vectors (fully
//     coalesced 16 B per lane) and is written per pixel with stride-3 STS (conflict free); the
//     normal map is only read for vote candidates (rare), not streamed;
//   * pixel groups without any point (sky, masked regions - contiguous in practice) are skipped per
//     warp;
//   * neighbour depth gathers go through the read-only path; consecutive lanes hit consecutive
//     pixels of the neighbour map, and CTAs are scheduled view by view so the K neighbour maps of
//     the views in flight stay L2 resident;
//   * float->int truncation uses FADD.RZ with 2^23 instead of F2I (keeps the XU pipe free for the
//     one MUFU.RCP and one MUFU.SQRT per pair).
#include <string.h>

#include "fuse_common.cuh"

namespace ddn {

#ifndef DDN_K4_MINBLOCKS
#define DDN_K4_MINBLOCKS 5
#endif
constexpr int kFilterThreads = 256;
#ifndef DDN_K4_PX
#define DDN_K4_PX 4
#endif
constexpr int kFilterPX = DDN_K4_PX;  // pixels per thread
constexpr int kFilterChunk = kFilterThreads * kFilterPX;

// ------------------------------------------------------------------------------------------------
// Table set-up (float64, one thread per (source view, neighbour))
// pair_table[s][k][24]: rows 0..2 of K_t [M | t_ts], M = R_t R_s^T Kinv_s (12 floats) - i.e. the pixel
//   (x, y) at depth d maps to (U, V, Z) = rows * (d x, d y, d, 1) and lands at u = U/Z, v = V/Z;
//   camera centre of t (3), target view index (int bits), fx_t fy_t cx_t cy_t, own-view flag, and (float 21,
//   uint bits) K4's gather offset of the entry: t*H*W - 0x4B000000*(W+1) mod 2^32 (see pair_gather).
//   The entries of a source view are COMPACTED: valid neighbours other than the view itself first
//   (n_hot of them, in table order), then the n_own entries of the view itself (only the
//   reference-parity table K = V lists it); entry 0 carries n_hot and n_own in floats 22, 23 (int bits).  K4's hot loop therefore runs over n_hot entries without any validity test.
// src_table[s][16]: rows of R_s^T Kinv_s with c_s appended (12 floats), cx_s, cy_s, fx_s, fy_s.
// ------------------------------------------------------------------------------------------------
__global__ void build_pair_tables_kernel(int n_total, int src_begin, int n_src, int k_nbr, unsigned hw, unsigned width,
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
  const int32_t* row = nbr + (size_t)s * k_nbr;
  const int t = row[k];
  int n_hot = 0, n_own = 0, hot_before = 0, own_before = 0;
  for (int q = 0; q < k_nbr; ++q) {
    const int tq = row[q];
    const bool own = tq == s, hot = tq >= 0 && tq < n_total && !own;
    n_hot += hot ? 1 : 0;
    n_own += own ? 1 : 0;
    hot_before += (hot && q < k) ? 1 : 0;
    own_before += (own && q < k) ? 1 : 0;
  }
  if (k == 0) {
    float* e0 = pair_table + (size_t)sl * k_nbr * DDN_PAIR_TABLE_FLOATS;
    e0[22] = __int_as_float(n_hot);
    e0[23] = __int_as_float(n_own);
  }
  if (t < 0 || t >= n_total) return;
  const int pos = (t == s) ? n_hot + own_before : hot_before;
  float* o = pair_table + ((size_t)sl * k_nbr + pos) * DDN_PAIR_TABLE_FLOATS;
  const double* Pt = poses + (size_t)t * 12;
  double Rt[3][3], tt[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Rt[i][j] = Pt[i * 4 + j];
    tt[i] = Pt[i * 4 + 3];
  }
  double rows[3][4];
  for (int i = 0; i < 3; ++i) {
    double r[3];  // row i of R_t R_s^T
    for (int j = 0; j < 3; ++j) r[j] = Rt[i][0] * Rs[j][0] + Rt[i][1] * Rs[j][1] + Rt[i][2] * Rs[j][2];
    rows[i][0] = r[0] / fx;
    rows[i][1] = r[1] / fy;
    rows[i][2] = -r[0] * cx / fx - r[1] * cy / fy + r[2];
    rows[i][3] = tt[i] - (r[0] * ts[0] + r[1] * ts[1] + r[2] * ts[2]);
  }
  const double fxt = intr[t * 4 + 0], fyt = intr[t * 4 + 1], cxt = intr[t * 4 + 2], cyt = intr[t * 4 + 3];
  for (int j = 0; j < 4; ++j) {
    o[0 * 4 + j] = (float)(fxt * rows[0][j] + cxt * rows[2][j]);
    o[1 * 4 + j] = (float)(fyt * rows[1][j] + cyt * rows[2][j]);
    o[2 * 4 + j] = (float)rows[2][j];
  }
  for (int i = 0; i < 3; ++i)  // c_t = -R_t^T t_t
    o[12 + i] = (float)(-(Rt[0][i] * tt[0] + Rt[1][i] * tt[1] + Rt[2][i] * tt[2]));
  o[15] = __int_as_float(t);
  o[16] = (float)intr[t * 4 + 0];
  o[17] = (float)intr[t * 4 + 1];
  o[18] = (float)intr[t * 4 + 2];
  o[19] = (float)intr[t * 4 + 3];
  o[20] = (t == s) ? 1.f : 0.f;
  o[21] = __uint_as_float((unsigned)t * hw - 0x4B000000u * (width + 1u));
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
  FuseDev mark;              // occupancy marking of kept points (mark.units == nullptr: off)
  int src_begin, n_src, H, W, Hs, Ws, K, stride;
  int vote_threshold;
  float depth_threshold, grazing_cos, two_sided_tau;
  int normals_in_world;
  unsigned wbits, hbits;  // bit patterns of (float)W, (float)H for the bounds test on raw bits
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

// One (pixel, neighbour) evaluation, in two steps so that the gathers of a thread's four pixels are all in
// flight before the first one is consumed.  pair_gather: project, bounds test, issue the depth lookup.
// pair_test: the candidate flag ("would vote if the grazing gate passes").
template <bool kBilinear, bool kFoldMask>
__device__ __forceinline__ void pair_gather(const float4& r0, const float4& r1, const float4& r2, float P, float Q, float d,
                                            const float* __restrict__ depth_all, unsigned off_k, unsigned W, int H,
                                            unsigned wbits, unsigned hbits, float& Z, float& D, bool& inb) {
  const float U = fmaf(r0.x, P, fmaf(r0.y, Q, fmaf(r0.z, d, r0.w)));
  const float V = fmaf(r1.x, P, fmaf(r1.y, Q, fmaf(r1.z, d, r1.w)));
  Z = fmaf(r2.x, P, fmaf(r2.y, Q, fmaf(r2.z, d, r2.w)));
  const float inv = rcp_approx(Z);
  const float u = U * inv;
  const float v = V * inv;
  // 0 <= u < W and 0 <= v < H on the raw bits (negative floats and NaN compare as huge unsigned);
  // invalid pixels carry NaN depth, so Z > 0 rejects them too.
  inb = (__float_as_uint(u) < wbits) & (__float_as_uint(v) < hbits) & (Z > 0.f);
  const unsigned ub = (unsigned)trunc_biased(u), vb = (unsigned)trunc_biased(v);
  if (kFoldMask && !inb) Z = INFINITY;  // folds the bounds mask into the one-sided test: inf < thr*D is false
  if (!kBilinear) {
    // 32-bit element offset from the start of refined_all; off_k = t*H*W - bias*(W+1) removes the 2^23
    // biases by modular arithmetic.  Out-of-bounds lanes read element 0 and are masked by `inb`.
    const unsigned off = inb ? vb * W + ub + off_k : 0u;
    D = __ldg(depth_all + off);
  } else {
    const float* __restrict__ depth_t = depth_all + (off_k + (unsigned)kTruncBias * (W + 1u));
    // N3: 4 taps at floor(u), floor(v), +1 clamped; all taps must be > 0
    const int x0 = inb ? ub - kTruncBias : 0, y0 = inb ? vb - kTruncBias : 0;
    const int x1 = min(x0 + 1, (int)W - 1), y1 = min(y0 + 1, H - 1);
    const float ta = __ldg(depth_t + y0 * W + x0), tb = __ldg(depth_t + y0 * W + x1);
    const float tc = __ldg(depth_t + y1 * W + x0), td = __ldg(depth_t + y1 * W + x1);
    const float fxw = u - (float)x0, fyw = v - (float)y0;
    const float gx = 1.f - fxw, gy = 1.f - fyw;
    const float acc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(gx, gy), ta), __fmul_rn(__fmul_rn(fxw, gy), tb)),
                                          __fmul_rn(__fmul_rn(gx, fyw), tc)),
                                __fmul_rn(__fmul_rn(fxw, fyw), td));
    D = ((ta > 0.f) & (tb > 0.f) & (tc > 0.f) & (td > 0.f)) ? acc : 0.f;
  }
}

// cm |= bit when a < b, as one compare and one predicated OR (the compiler would otherwise build the bit with
// two selects)
__device__ __forceinline__ void or_bit_if_less(unsigned& cm, float a, float b, unsigned bit) {
  asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(cm) : "f"(a), "f"(b), "r"(bit));
}

template <bool kTwoSided>
__device__ __forceinline__ bool pair_test(float Z, float D, bool inb, float thr, float tau) {
  if (kTwoSided) return inb & (D > 0.f) & (fabsf(Z - D) > tau * D);
  // one-sided floater test z < float32(thr * D) (scripts/test.py:319-321, NEP-50 product in float32).
  // The lookup-valid gate D > 0 (:315) is implied: inb has Z > 0 and thr > 0, so Z < thr*D needs D > 0.
  return inb & (Z < __fmul_rn(thr, D));
}

// Normal of one source pixel for the grazing gate (scripts/test.py:291), read on demand: only vote
// candidates need it, and they are rare (floaters), so the 12 B/pixel normal map is not streamed.
// normals_in_world rotates it by R_s^T (rows recovered from the source table).
__device__ __forceinline__ void load_normal(const float* __restrict__ normal_view, size_t pixel, bool in_world,
                                            const float* __restrict__ s_src, float& n0, float& n1, float& n2) {
  const float* np_ = normal_view + pixel * 3;
  n0 = __ldg(np_ + 0);
  n1 = __ldg(np_ + 1);
  n2 = __ldg(np_ + 2);
  if (in_world) {
    const float fx = s_src[14], fy = s_src[15], cx = s_src[12], cy = s_src[13];
    float w[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float ra = s_src[i * 4 + 0] * fx, rb = s_src[i * 4 + 1] * fy;
      const float rc = s_src[i * 4 + 2] + s_src[i * 4 + 0] * cx + s_src[i * 4 + 1] * cy;
      w[i] = ra * n0 + rb * n1 + rc * n2;
    }
    n0 = w[0], n1 = w[1], n2 = w[2];
  }
}

// Bulk asynchronous copy shared -> global (the TMA engine's 1-D form, SASS UBLKCP): with -DDDN_K4_BULK=1 one elected
// lane hands a warp's 1536-byte xyz slice to the copy engine instead of 32 lanes moving it with three LDS.128 + three
// streaming STG.128 each.  Measured at cfg 2 (profiles/README.md, round 2): 2.29 ms against 2.23 ms for the per-lane
// copy (3.03 vs 2.97 ms with the occupancy mark fused in) - the proxy fence and the wait for the engine's read before
// the CTA may retire cost more than the twelve instructions per thread they replace - so the per-lane copy stays.
#ifndef DDN_K4_BULK
#define DDN_K4_BULK 0
#endif
__device__ __forceinline__ void fence_proxy_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_shared_to_global(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\tcp.async.bulk.commit_group;"
               :
               : "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Pixel ownership.  A CTA owns kFilterChunk consecutive source-grid pixels of one view, a warp 128 of them.
//   layout 0 ("strided"):  thread pixel j = warp base + j*32 + lane.  Every load, gather and byte store of a
//                          warp instruction covers 32 consecutive pixels; xyz leaves through a per-warp
//                          shared staging buffer as float4 (no block barrier).
//   layout 1 ("adjacent"): thread pixel j = warp base + lane*4 + j.  One LDG.128 for the depths, one STG.32
//                          for the votes, three STG.128 for xyz straight from registers, no shared staging;
//                          the gathers of one instruction are 4 pixels apart.
// -DDDN_K4_LAYOUT selects the default; both produce identical results.
#ifndef DDN_K4_LAYOUT
#define DDN_K4_LAYOUT 0
#endif
constexpr int kWarpPix = 32 * kFilterPX;
static_assert(kFilterPX == 4, "the layouts below are written for 4 pixels per thread");

template <bool kBilinear, bool kStride1, bool kTwoSided, int kLayout>
__global__ void __launch_bounds__(kFilterThreads, kBilinear ? 3 : DDN_K4_MINBLOCKS) backproject_filter_kernel(const FilterParams p) {
  extern __shared__ __align__(16) float smem[];
  float* s_src = smem;                     // 16 floats
  float* s_pair = s_src + 16;              // K*24 floats
  float* s_stage = s_pair + p.K * DDN_PAIR_TABLE_FLOATS;  // layout 0: 8 warps x 128 px x 3 floats
  __shared__ int s_bbox[6];
  __shared__ int s_marked;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sl = blockIdx.y;
  const int s = p.src_begin + sl;
  const int Ps = p.Hs * p.Ws;  // source-grid pixels per view
  const int chunk0 = blockIdx.x * kFilterChunk;
  const size_t HW = (size_t)p.H * p.W;
  const float* __restrict__ depth_s = p.refined_all + (size_t)s * HW;
  const float* __restrict__ normal_s = p.normal + (size_t)sl * HW * 3;

  // tables -> shared memory, as float4 (entries are 96 bytes; the gather offset t*H*W - bias*(W+1) of every
  // entry was put into float 21 by build_pair_tables_kernel)
  {
    const float4* g4 = reinterpret_cast<const float4*>(p.pair_table + (size_t)sl * p.K * DDN_PAIR_TABLE_FLOATS);
    float4* s4 = reinterpret_cast<float4*>(s_pair);
    for (int i = tid; i < p.K * (DDN_PAIR_TABLE_FLOATS / 4); i += kFilterThreads) s4[i] = __ldg(g4 + i);
    if (tid < 4) reinterpret_cast<float4*>(s_src)[tid] = __ldg(reinterpret_cast<const float4*>(p.src_table + (size_t)sl * 16) + tid);
    if (tid < 6) s_bbox[tid] = tid < 3 ? 0x7fffffff : (int)0x80000000;
    if (tid == 6) s_marked = 0;
  }

  // first pixel of the thread on the source grid: q0 = y0 * Ws + x0 (one integer division per thread)
  const int wbase = chunk0 + warp * kWarpPix;
  const int q0 = wbase + (kLayout == 1 ? lane * kFilterPX : lane);
  constexpr int kStep = kLayout == 1 ? 1 : 32;
  int ys_run = q0 / p.Ws;
  int xs_run = q0 - ys_run * p.Ws;

  // Per-pixel state kept in registers: depth d (NaN = no point), P = d*x, Q = d*y, vote count.
  float d[kFilterPX], P[kFilterPX], Q[kFilterPX];
  int nvotes[kFilterPX];
  unsigned src_pix[kFilterPX];  // element index of the pixel in its own full-resolution map
  const float qnan = __int_as_float(0x7fc00000);
  {
    float dd[kFilterPX];
    int px[kFilterPX], py[kFilterPX];
#pragma unroll
    for (int j = 0; j < kFilterPX; ++j) {
      px[j] = kStride1 ? xs_run : xs_run * p.stride;
      py[j] = kStride1 ? ys_run : ys_run * p.stride;
      src_pix[j] = kStride1 ? (unsigned)(q0 + j * kStep) : (unsigned)py[j] * (unsigned)p.W + (unsigned)px[j];
      xs_run += kStep;
      while (xs_run >= p.Ws) {
        xs_run -= p.Ws;
        ++ys_run;
      }
    }
    if (kLayout == 1 && kStride1 && q0 + kFilterPX <= Ps && ((reinterpret_cast<uintptr_t>(depth_s + q0) & 15) == 0)) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(depth_s + q0));
      dd[0] = v.x, dd[1] = v.y, dd[2] = v.z, dd[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < kFilterPX; ++j) dd[j] = (q0 + j * kStep) < Ps ? __ldg(depth_s + src_pix[j]) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kFilterPX; ++j) {
      const bool valid = dd[j] > 0.f;
      d[j] = valid ? dd[j] : qnan;
      asm volatile("" : "+f"(d[j]));  // keep the NaN-tagged depth in a register (no per-neighbour recompute)
      P[j] = d[j] * (float)px[j];
      Q[j] = d[j] * (float)py[j];
      nvotes[j] = 0;
    }
  }
  __syncthreads();  // tables visible
  // warp-uniform flags: pixel group j of this warp has at least one point (sky / masked regions are
  // contiguous, so whole groups drop out of the pair loop)
  bool live[kFilterPX];
  bool any_live = false;
#pragma unroll
  for (int j = 0; j < kFilterPX; ++j) {
    live[j] = __any_sync(0xffffffffu, d[j] > 0.f);
    any_live |= live[j];
  }

  const unsigned wbits = p.wbits, hbits = p.hbits;
  const float thr = p.depth_threshold, gcos = p.grazing_cos, tau = p.two_sided_tau;
  const bool in_world = p.normals_in_world != 0;
  const int n_hot = __float_as_int(s_pair[22]), n_own = __float_as_int(s_pair[23]);

  // world position of pixel j (scripts/test.py:79-90 then :233), fp32 with float64-precomputed rows;
  // the same expression ddn_align_views' bounding-box epilogue evaluates (backproject_px)
  auto world_of = [&](int j, float& X, float& Y, float& Z) { backproject_pqd(s_src, P[j], Q[j], d[j], X, Y, Z); };

  // Hot loop: branch-free candidate test for every (pixel, neighbour); the four gathers of a thread
  // issue back to back.  Candidates ("would vote if the grazing gate passes") are only recorded as a
  // bit per neighbour; the gate itself (scripts/test.py:284-295) needs the pixel's normal and is
  // resolved after each block of 32 neighbours, once per pixel that has a candidate at all.
  if (any_live) {
    bool all_live = true;
#pragma unroll
    for (int j = 0; j < kFilterPX; ++j) all_live &= live[j];
    for (int k0 = 0; k0 < n_hot; k0 += 32) {
      const int k1 = min(k0 + 32, n_hot);
      unsigned cm[kFilterPX];
#pragma unroll
      for (int j = 0; j < kFilterPX; ++j) cm[j] = 0u;
      unsigned bit = 1u;
      for (int k = k0; k < k1; ++k, bit <<= 1) {
        const float4* t4 = reinterpret_cast<const float4*>(s_pair + k * DDN_PAIR_TABLE_FLOATS);
        const float4 r0 = t4[0], r1 = t4[1], r2 = t4[2];
        const unsigned off_k = __float_as_uint(s_pair[k * DDN_PAIR_TABLE_FLOATS + 21]);
        float Z[kFilterPX], D[kFilterPX];
        bool inb[kFilterPX];
        if (all_live) {
#pragma unroll
          for (int j = 0; j < kFilterPX; ++j)
            pair_gather<kBilinear, !kBilinear && !kTwoSided>(r0, r1, r2, P[j], Q[j], d[j], p.refined_all, off_k, (unsigned)p.W, p.H, wbits, hbits, Z[j], D[j],
                                   inb[j]);
#pragma unroll
          for (int j = 0; j < kFilterPX; ++j) {
            if (!kBilinear && !kTwoSided) {
              or_bit_if_less(cm[j], Z[j], __fmul_rn(thr, D[j]), bit);  // Z is +inf for out-of-bounds lanes
              continue;
            }
            if (pair_test<kTwoSided>(Z[j], D[j], inb[j], thr, tau)) cm[j] |= bit;
          }
        } else {
#pragma unroll
          for (int j = 0; j < kFilterPX; ++j) {
            if (live[j]) {
              pair_gather<kBilinear, !kBilinear && !kTwoSided>(r0, r1, r2, P[j], Q[j], d[j], p.refined_all, off_k, (unsigned)p.W, p.H, wbits, hbits, Z[j],
                                     D[j], inb[j]);
              if (pair_test<kTwoSided>(Z[j], D[j], inb[j], thr, tau)) cm[j] |= bit;
            }
          }
        }
      }
      // dot(n, -(Xw - c_t)/|Xw - c_t|) > cos  <=>  dot(n, c_t - Xw) > cos * |c_t - Xw|
#pragma unroll
      for (int j = 0; j < kFilterPX; ++j) {
        if (cm[j]) {
          float n0, n1, n2, wx, wy, wz;
          load_normal(normal_s, src_pix[j], in_world, s_src, n0, n1, n2);
          world_of(j, wx, wy, wz);
          unsigned m = cm[j];
          while (m) {
            const int k = k0 + __ffs(m) - 1;
            m &= m - 1;
            const float4 cc = reinterpret_cast<const float4*>(s_pair + k * DDN_PAIR_TABLE_FLOATS)[3];
            const float ex = cc.x - wx, ey = cc.y - wy, ez = cc.z - wz;
            const float dn = fmaf(n0, ex, fmaf(n1, ey, n2 * ez));
            const float len = sqrt_approx(fmaf(ex, ex, fmaf(ey, ey, ez * ez)));
            nvotes[j] += (dn > gcos * len) ? 1 : 0;
          }
        }
      }
    }
  }

  if (n_own > 0) {
    // Own view (only present in the reference-parity table K = V).  The reference normalises by (z + 1e-8)
    // before applying K (scripts/test.py:71-75), so u = x * z/(z+1e-8) lands ~x*1e-8/z BELOW the integer x
    // (far above float64 round-off) and the truncation at :308-309 looks up pixel (x-1, y-1) for
    // x, y >= 1.  Reproduced in integer arithmetic; z is the pixel's own depth.  (x == 0 or y == 0 is a
    // round-off tie in the reference.)
    for (int k = n_hot; k < n_hot + n_own; ++k) {
      const float4* t4 = reinterpret_cast<const float4*>(s_pair + k * DDN_PAIR_TABLE_FLOATS);
      const float4 cc = t4[3];
#pragma unroll
      for (int j = 0; j < kFilterPX; ++j) {
        if (d[j] > 0.f) {
          const int py = (int)(src_pix[j] / (unsigned)p.W), px = (int)(src_pix[j] - (unsigned)py * (unsigned)p.W);
          const int ux = max(px - 1, 0), vy = max(py - 1, 0);
          const float D = __ldg(depth_s + (size_t)vy * p.W + ux);
          const bool bad = kTwoSided ? (fabsf(d[j] - D) > tau * D) : (d[j] < __fmul_rn(thr, D));
          if (D > 0.f && bad) {
            float n0, n1, n2, wx, wy, wz;
            load_normal(normal_s, src_pix[j], in_world, s_src, n0, n1, n2);
            world_of(j, wx, wy, wz);
            const float ex = cc.x - wx, ey = cc.y - wy, ez = cc.z - wz;
            const float dn = fmaf(n0, ex, fmaf(n1, ey, n2 * ez));
            const float len = sqrt_approx(fmaf(ex, ex, fmaf(ey, ey, ez * ez)));
            nvotes[j] += (dn > gcos * len) ? 1 : 0;
          }
        }
      }
    }
  }

  // ---- epilogue: world positions (recomputed from the registers), votes, bounding box, occupancy ----
  bool bulk_pending = false;  // lane 0: a bulk copy out of this warp's staging slice is in flight
  float X[kFilterPX], Y[kFilterPX], Zw[kFilterPX];
  bool keep[kFilterPX];
  unsigned vote_bytes = 0;
#pragma unroll
  for (int j = 0; j < kFilterPX; ++j) {
    const bool valid = d[j] > 0.f;
    world_of(j, X[j], Y[j], Zw[j]);
    if (!valid) X[j] = 0.f, Y[j] = 0.f, Zw[j] = 0.f;
    keep[j] = valid && nvotes[j] < p.vote_threshold;
    vote_bytes |= (valid ? (unsigned)min(nvotes[j], 254) : 255u) << (8 * j);
  }
  const size_t out0 = (size_t)sl * Ps;  // first pixel of the view in the outputs
  if (kLayout == 1) {
    uint8_t* vo = p.votes + out0 + q0;
    if (q0 + kFilterPX <= Ps && ((reinterpret_cast<uintptr_t>(vo) & 3) == 0)) {
      *reinterpret_cast<unsigned*>(vo) = vote_bytes;
    } else {
#pragma unroll
      for (int j = 0; j < kFilterPX; ++j)
        if (q0 + j < Ps) vo[j] = (uint8_t)(vote_bytes >> (8 * j));
    }
    float* xo = p.xyz + (out0 + q0) * 3;
    if (q0 + kFilterPX <= Ps && ((reinterpret_cast<uintptr_t>(xo) & 15) == 0)) {
      float4* x4 = reinterpret_cast<float4*>(xo);
      __stcs(x4 + 0, make_float4(X[0], Y[0], Zw[0], X[1]));
      __stcs(x4 + 1, make_float4(Y[1], Zw[1], X[2], Y[2]));
      __stcs(x4 + 2, make_float4(Zw[2], X[3], Y[3], Zw[3]));
    } else {
#pragma unroll
      for (int j = 0; j < kFilterPX; ++j)
        if (q0 + j < Ps) {
          __stcs(xo + j * 3 + 0, X[j]);
          __stcs(xo + j * 3 + 1, Y[j]);
          __stcs(xo + j * 3 + 2, Zw[j]);
        }
    }
  } else {
#pragma unroll
    for (int j = 0; j < kFilterPX; ++j)
      if (q0 + j * 32 < Ps) p.votes[out0 + q0 + j * 32] = (uint8_t)(vote_bytes >> (8 * j));
    // xyz through the warp's staging slice: stride-3 STS (conflict free), float4 out (16 B per lane, coalesced)
    float* sw = s_stage + warp * (kWarpPix * 3);
#pragma unroll
    for (int j = 0; j < kFilterPX; ++j) {
      const int l = j * 32 + lane;
      sw[l * 3 + 0] = X[j];
      sw[l * 3 + 1] = Y[j];
      sw[l * 3 + 2] = Zw[j];
    }
#if DDN_K4_BULK
    fence_proxy_async_shared();  // the slice was written through the generic proxy: make it visible to the copy engine
#endif
    __syncwarp();
    const int n_w = min(kWarpPix, Ps - wbase);  // pixels of this warp that exist (<= 0: none)
    float* xo = p.xyz + (out0 + wbase) * 3;
    if (n_w == kWarpPix && ((reinterpret_cast<uintptr_t>(xo) & 15) == 0)) {
#if DDN_K4_BULK
      if (lane == 0) {
        bulk_store_shared_to_global(xo, sw, kWarpPix * 3 * sizeof(float));
        bulk_pending = true;
      }
#else
      const float4* s4 = reinterpret_cast<const float4*>(sw);
      float4* g4 = reinterpret_cast<float4*>(xo);
#pragma unroll
      for (int q = 0; q < 3; ++q) __stcs(g4 + q * 32 + lane, s4[q * 32 + lane]);
#endif
    } else {
      for (int i = lane; i < n_w * 3; i += 32) __stcs(xo + i, sw[i]);
    }
  }

  if (p.bbox != nullptr && any_live) {
    float bmin[3] = {INFINITY, INFINITY, INFINITY}, bmax[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < kFilterPX; ++j) {
      if (keep[j]) {
        bmin[0] = fminf(bmin[0], X[j]), bmax[0] = fmaxf(bmax[0], X[j]);
        bmin[1] = fminf(bmin[1], Y[j]), bmax[1] = fmaxf(bmax[1], Y[j]);
        bmin[2] = fminf(bmin[2], Zw[j]), bmax[2] = fmaxf(bmax[2], Zw[j]);
      }
    }
    // warp reduction with REDUX on the order-preserving int encoding, then one shared atomic per warp
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int lo = __reduce_min_sync(0xffffffffu, float_to_ordered(bmin[i]));
      const int hi = __reduce_max_sync(0xffffffffu, float_to_ordered(bmax[i]));
      if (lane == 0 && lo <= hi) {
        atomicMin(&s_bbox[i], lo);
        atomicMax(&s_bbox[3 + i], hi);
      }
    }
  }

  // Stage 4's mark pass, fused: the occupancy bit of every kept point.  A cell equal to that of the pixel
  // to the left (same thread, or the neighbouring lane) is already set by that pixel.
  if (p.mark.units != nullptr && any_live) {
    const GridDev g = *p.mark.grid;
    if (g.n_units != 0) {
      const float rv = 1.0f / g.voxel;
      uint64_t cell[kFilterPX];
      int mine = 0;
#pragma unroll
      for (int j = 0; j < kFilterPX; ++j) {
        cell[j] = kNoCell;
        if (keep[j]) {
          uint32_t kx, ky, kz;
          cell[j] = cell_of_point(g, rv, X[j], Y[j], Zw[j], kx, ky, kz);
        }
        mine += cell[j] != kNoCell ? 1 : 0;
      }
#pragma unroll
      for (int j = 0; j < kFilterPX; ++j) {
        // left neighbour: layout 1 - previous pixel of the thread, or the last pixel of the previous lane;
        // layout 0 - the same group in the previous lane
        uint64_t left = __shfl_up_sync(0xffffffffu, kLayout == 1 ? cell[kFilterPX - 1] : cell[j], 1);
        if (lane == 0) left = kNoCell;
        if (kLayout == 1 && j > 0) left = cell[j - 1];
        if (cell[j] != kNoCell && cell[j] != left) set_cell_bit(p.mark.units, p.mark.dirty, cell[j], left);
      }
      mine = __reduce_add_sync(0xffffffffu, mine);
      if (lane == 0 && mine) atomicAdd(&s_marked, mine);
    }
  }
  __syncthreads();
  if (p.bbox != nullptr && tid < 6) {
    if (tid < 3) {
      if (s_bbox[tid] != 0x7fffffff) atomicMin(p.bbox + tid, s_bbox[tid]);
    } else {
      if (s_bbox[tid] != (int)0x80000000) atomicMax(p.bbox + tid, s_bbox[tid]);
    }
  }
  if (tid == 6 && s_marked) atomicAdd(p.mark.counts, (unsigned long long)s_marked);
  if (bulk_pending) bulk_store_wait_read();  // the staging slice must stay intact until the copy engine has read it
}

__global__ void bbox_init_kernel(int* bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = float_to_ordered(INFINITY);
  else if (threadIdx.x < 6) bbox[threadIdx.x] = float_to_ordered(-INFINITY);
}

}  // namespace ddn

extern "C" {

int ddn_build_pair_tables(int64_t n_views_total, int64_t src_begin, int64_t n_src, int64_t k_nbr, int64_t height,
                          int64_t width, const double* cam_from_world, const double* intr, const int32_t* nbr,
                          float* pair_table, float* src_table, void* stream) {
  using namespace ddn;
  DDN_REQUIRE(n_views_total > 0 && n_src >= 0 && k_nbr > 0, "view counts");
  DDN_REQUIRE(height > 0 && width > 0 && height * width < (1ll << 31), "image size");
  DDN_REQUIRE(src_begin >= 0 && src_begin + n_src <= n_views_total, "source range");
  DDN_REQUIRE(cam_from_world && intr && nbr && pair_table && src_table, "null pointer");
  if (n_src == 0) return DDN_OK;
  const int total = (int)(n_src * k_nbr + n_src);
  build_pair_tables_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      (int)n_views_total, (int)src_begin, (int)n_src, (int)k_nbr, (unsigned)(height * width), (unsigned)width, cam_from_world, intr,
      nbr, pair_table, src_table);
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
                           float* xyz, uint8_t* votes, float* bbox, const ddn_fuse_session* mark, void* stream) {
  using namespace ddn;
  (void)nbr;
  DDN_REQUIRE(cfg != nullptr, "null config");
  DDN_REQUIRE(n_views_total > 0 && n_src >= 0 && k_nbr > 0 && k_nbr <= 1024, "view counts (at most 1024 neighbours per view)");
  DDN_REQUIRE(src_begin >= 0 && src_begin + n_src <= n_views_total, "source range");
  DDN_REQUIRE(height > 0 && width > 0 && height * width < (1ll << 31), "image size");
  DDN_REQUIRE(n_views_total * height * width < (1ll << 32), "refined_all must hold fewer than 2^32 pixels (32-bit gather offsets)");
  DDN_REQUIRE(width < (1 << 22) && height < (1 << 22), "image side too large for the truncation trick");
  DDN_REQUIRE(cfg->stride >= 1, "stride");
  DDN_REQUIRE(cfg->pixel_layout == 0 || cfg->pixel_layout == 1, "pixel_layout");
  DDN_REQUIRE(refined_all && normal && pair_table && src_table && xyz && votes, "null pointer");
  DDN_REQUIRE((reinterpret_cast<uintptr_t>(pair_table) & 15) == 0 && (reinterpret_cast<uintptr_t>(src_table) & 15) == 0,
              "pair_table / src_table must be 16-byte aligned");
  if (n_src == 0) return DDN_OK;
  FilterParams p;
  p.refined_all = refined_all;
  p.normal = normal;
  p.pair_table = pair_table;
  p.src_table = src_table;
  p.xyz = xyz;
  p.votes = votes;
  p.bbox = reinterpret_cast<int*>(bbox);
  p.mark.grid = nullptr, p.mark.units = nullptr, p.mark.dirty = nullptr, p.mark.counts = nullptr;
  if (mark != nullptr) {
    DDN_REQUIRE(mark->grid && mark->units && mark->counts, "mark session: null buffer");
    p.mark.grid = reinterpret_cast<const GridDev*>(mark->grid);
    p.mark.units = reinterpret_cast<uint32_t*>(mark->units);
    p.mark.dirty = mark->dirty;
    p.mark.counts = reinterpret_cast<unsigned long long*>(mark->counts);
  }
  p.src_begin = (int)src_begin;
  p.n_src = (int)n_src;
  p.H = (int)height;
  p.W = (int)width;
  p.stride = cfg->stride;
  p.Hs = (p.H + p.stride - 1) / p.stride;
  p.Ws = (p.W + p.stride - 1) / p.stride;
  p.K = (int)k_nbr;
  p.vote_threshold = vote_threshold < 255 ? vote_threshold : 255;
  p.depth_threshold = cfg->depth_threshold;
  p.grazing_cos = cfg->grazing_cos;
  p.two_sided_tau = cfg->two_sided_tau;
  p.normals_in_world = cfg->normals_in_world;
  {
    const float wf = (float)p.W, hf = (float)p.H;
    memcpy(&p.wbits, &wf, 4);
    memcpy(&p.hbits, &hf, 4);
  }
  const int Ps = p.Hs * p.Ws;
  dim3 grid((Ps + kFilterChunk - 1) / kFilterChunk, (unsigned)n_src);
  DDN_REQUIRE(n_src <= 65535, "too many source views per call");
  const int layout = cfg->pixel_layout;
  const size_t smem = (size_t)(16 + p.K * DDN_PAIR_TABLE_FLOATS + (layout == 0 ? kFilterChunk * 3 : 0)) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  const bool bil = cfg->sample_mode == 1;
  const bool s1 = cfg->stride == 1;
  const bool two = cfg->two_sided_tau > 0.f;
  DDN_REQUIRE(two || cfg->depth_threshold > 0.f, "depth_threshold must be positive");
#define DDN_LAUNCH_FILTER(B, S, T, L)                                                                      \
  do {                                                                                                     \
    DDN_TRY(check_cuda(cudaFuncSetAttribute(backproject_filter_kernel<B, S, T, L>,                         \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),       \
                       "cudaFuncSetAttribute"));                                                           \
    backproject_filter_kernel<B, S, T, L><<<grid, kFilterThreads, smem, st>>>(p);                          \
  } while (0)
#define DDN_LAUNCH_FILTER_L(B, S, T)       \
  do {                                     \
    if (layout == 1) DDN_LAUNCH_FILTER(B, S, T, 1); \
    else DDN_LAUNCH_FILTER(B, S, T, 0);    \
  } while (0)
  if (two) {
    if (bil && s1) DDN_LAUNCH_FILTER_L(true, true, true);
    else if (bil) DDN_LAUNCH_FILTER_L(true, false, true);
    else if (s1) DDN_LAUNCH_FILTER_L(false, true, true);
    else DDN_LAUNCH_FILTER_L(false, false, true);
  } else {
    if (bil && s1) DDN_LAUNCH_FILTER_L(true, true, false);
    else if (bil) DDN_LAUNCH_FILTER_L(true, false, false);
    else if (s1) DDN_LAUNCH_FILTER_L(false, true, false);
    else DDN_LAUNCH_FILTER_L(false, false, false);
  }
#undef DDN_LAUNCH_FILTER_L
#undef DDN_LAUNCH_FILTER
  return after_launch("backproject_filter_kernel");
}

}  // extern "C"
