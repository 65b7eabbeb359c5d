// Stage 4: voxel-grid fusion (new capability; the reference only concatenates
// points, scripts/test.py:353-359).  DENSE-RANK PATH - no sort.
//
// The voxel grid of a scene is small enough for one occupancy BIT per cell (cfg 2: 2400 x 2400 x 600
// cells = 0.43 GB of bits in 180 GB of HBM).  Bit order == key order (x fastest, then y, then z), so
// the rank of a set bit among all set bits IS the voxel's position in the key-sorted output.  That
// replaces the radix sort of 168 M (key, index) pairs by streaming passes:
//
//   mark        every participating point sets the bit of its cell (RED.OR; lane-neighbour dedupe)
//   rank        popcount scan over the bitmap: one exclusive prefix per 256-bit group (= one 32 B
//               sector); the pass also emits the sorted keys and zeroes the accumulators it hands out
//   accumulate  every point looks up its slot (sector + prefix, both L2-resident because consecutive
//               pixels fall into neighbouring cells) and adds integer fixed-point sums with 64-bit
//               RED.ADD.  A thread owns 8 consecutive points and merges runs of equal cells in
//               registers first, which removes about half of the atomics.
//   finalize    one thread per voxel: mean = centre + sum/count, colour = round-half-up
//
// Integer sums make the result independent of the order of points and of how they are split over
// ranks (ddn_voxel_partials / ddn_voxel_merge use the same passes with a different output / input).
// Grids with more than 2^35 cells take the sort path in fuse_sort.cu.
#include "fuse_common.cuh"

namespace ddn {

constexpr int kGroupBits = 256;              // rank granularity: 8 words = one 32 B sector
constexpr int kTileGroups = 2048;            // groups per CTA in the rank passes (64 KB of bitmap)
constexpr int kScanThreads = 256;
constexpr int kAccWords = 5;                 // sx, sy, sz, r:g, b:count (u64 each)
constexpr uint64_t kDenseMaxCells = 1ull << 35;  // 4 GiB of bits
constexpr uint64_t kNoCell = ~0ull;

__device__ __forceinline__ uint64_t cell_of_point(const GridDev& g, float rv, float x, float y, float z, uint32_t& kx,
                                                  uint32_t& ky, uint32_t& kz) {
  const float fx = voxel_coord(x, g.ox, g.voxel, rv);
  const float fy = voxel_coord(y, g.oy, g.voxel, rv);
  const float fz = voxel_coord(z, g.oz, g.voxel, rv);
  const bool inside = fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)g.nx && fy < (float)g.ny && fz < (float)g.nz;
  if (!inside) return kNoCell;
  kx = (uint32_t)fx;
  ky = (uint32_t)fy;
  kz = (uint32_t)fz;
  return (uint64_t)kx + (uint64_t)g.nx * ((uint64_t)ky + (uint64_t)g.ny * (uint64_t)kz);
}

__device__ __forceinline__ uint64_t cell_of_key(const GridDev& g, uint64_t key) {
  const uint64_t kx = key & 0x1fffff, ky = (key >> 21) & 0x1fffff, kz = (key >> 42) & 0x1fffff;
  if ((key >> 63) || kx >= (uint64_t)g.nx || ky >= (uint64_t)g.ny || kz >= (uint64_t)g.nz) return kNoCell;
  return kx + (uint64_t)g.nx * (ky + (uint64_t)g.ny * kz);
}

// Set the bit of `cell`; lanes whose left neighbour holds the same cell skip the atomic.  Returns the
// number of participating lanes of the warp (valid in every lane).
__device__ __forceinline__ int mark_cell(uint32_t* __restrict__ bitmap, uint64_t cell) {
  const uint64_t prev = __shfl_up_sync(0xffffffffu, cell, 1);
  const bool valid = cell != kNoCell;
  const bool first = (threadIdx.x & 31) == 0 || prev != cell;
  if (valid && first) atomicOr(bitmap + (cell >> 5), 1u << (cell & 31));
  return __popc(__ballot_sync(0xffffffffu, valid));
}

__global__ void __launch_bounds__(256)
mark_points_kernel(GridDev g, float rv, int64_t n, const float* __restrict__ xyz, const uint8_t* __restrict__ votes, int thr,
                   uint32_t* __restrict__ bitmap, unsigned long long* __restrict__ n_in) {
  __shared__ int s_count;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  uint64_t cell = kNoCell;
  if (i < n && (votes == nullptr || (int)__ldg(votes + i) < thr)) {
    uint32_t kx, ky, kz;
    cell = cell_of_point(g, rv, __ldg(xyz + i * 3 + 0), __ldg(xyz + i * 3 + 1), __ldg(xyz + i * 3 + 2), kx, ky, kz);
  }
  const int c = mark_cell(bitmap, cell);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_count, c);
  __syncthreads();
  if (threadIdx.x == 0 && s_count) atomicAdd(n_in, (unsigned long long)s_count);
}

__global__ void __launch_bounds__(256)
mark_records_kernel(GridDev g, int64_t n, const uint64_t* __restrict__ keys, uint32_t* __restrict__ bitmap,
                    unsigned long long* __restrict__ n_in) {
  __shared__ int s_count;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const uint64_t cell = i < n ? cell_of_key(g, __ldg(keys + i)) : kNoCell;
  const int c = mark_cell(bitmap, cell);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_count, c);
  __syncthreads();
  if (threadIdx.x == 0 && s_count) atomicAdd(n_in, (unsigned long long)s_count);
}

// ---- rank: popcount scan over the bitmap ---------------------------------------------------------
__device__ __forceinline__ int popc8(const uint4& a, const uint4& b) {
  return __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
}

__global__ void __launch_bounds__(kScanThreads)
tile_count_kernel(const uint4* __restrict__ bitmap4, uint32_t groups, uint32_t* __restrict__ tile_sums) {
  __shared__ int s_warp[kScanThreads / 32];
  const uint32_t base = blockIdx.x * kTileGroups;
  int sum = 0;
#pragma unroll
  for (int j = 0; j < kTileGroups / kScanThreads; ++j) {
    const uint32_t gi = base + j * kScanThreads + threadIdx.x;
    if (gi < groups) sum += popc8(__ldg(bitmap4 + 2 * (size_t)gi), __ldg(bitmap4 + 2 * (size_t)gi + 1));
  }
  sum = __reduce_add_sync(0xffffffffu, sum);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) t += s_warp[w];
    tile_sums[blockIdx.x] = (uint32_t)t;
  }
}

// exclusive scan of the tile sums in place (one CTA), total -> counts_out[1]
__global__ void __launch_bounds__(1024) tile_scan_kernel(uint32_t* __restrict__ tile_sums, int tiles, int64_t* __restrict__ counts_out) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < tiles; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < tiles ? tile_sums[i] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += t;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    const uint32_t excl = carry + (warp ? s_warp[warp - 1] : 0u) + inc - v;
    if (i < tiles) tile_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) counts_out[1] = (int64_t)s_carry;
}

// Per group: exclusive rank prefix.  Per set bit: canonical key of the cell -> keys[slot], and the
// slot's accumulators are zeroed (so no separate memset sized by a device-side count is needed).
__global__ void __launch_bounds__(kScanThreads)
group_prefix_kernel(GridDev g, const uint4* __restrict__ bitmap4, uint32_t groups, const uint32_t* __restrict__ tile_excl,
                    uint32_t* __restrict__ group_prefix, uint64_t* __restrict__ keys, unsigned long long* __restrict__ accum) {
  __shared__ uint32_t s_warp[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = blockIdx.x * kTileGroups;
  uint32_t carry = tile_excl[blockIdx.x];
  const uint64_t nxy = (uint64_t)g.nx * (uint64_t)g.ny;
#pragma unroll 1
  for (int j = 0; j < kTileGroups / kScanThreads; ++j) {
    const uint32_t gi = base + j * kScanThreads + threadIdx.x;
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (gi < groups) {
      const uint4 a = __ldg(bitmap4 + 2 * (size_t)gi), b = __ldg(bitmap4 + 2 * (size_t)gi + 1);
      w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w, w[4] = b.x, w[5] = b.y, w[6] = b.z, w[7] = b.w;
    }
    uint32_t cnt = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) cnt += __popc(w[i]);
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int q = 0; q < kScanThreads / 32; ++q) {
      const uint32_t s = s_warp[q];
      before += q < warp ? s : 0u;
      total += s;
    }
    __syncthreads();
    uint32_t slot = carry + before + inc - cnt;
    carry += total;
    if (gi < groups) group_prefix[gi] = slot;
    if (cnt) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t bits = w[i];
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          const uint64_t cell = (uint64_t)gi * kGroupBits + i * 32 + b;
          const uint64_t kz = cell / nxy, rem = cell - kz * nxy;
          const uint64_t ky = rem / (uint64_t)g.nx, kx = rem - ky * (uint64_t)g.nx;
          keys[slot] = kx | (ky << 21) | (kz << 42);
          if (accum != nullptr) {
            unsigned long long* a = accum + (size_t)slot * kAccWords;
#pragma unroll
            for (int q = 0; q < kAccWords; ++q) a[q] = 0ull;
          }
          ++slot;
        }
      }
    }
  }
}

__device__ __forceinline__ uint32_t slot_of_cell(uint64_t cell, const uint4* __restrict__ bitmap4,
                                                 const uint32_t* __restrict__ group_prefix) {
  const size_t gi = (size_t)(cell >> 8);
  const int wi = (int)(cell >> 5) & 7;
  const uint32_t below = (1u << (cell & 31)) - 1u;
  const uint4 a = __ldg(bitmap4 + 2 * gi), b = __ldg(bitmap4 + 2 * gi + 1);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t s = __ldg(group_prefix + gi);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t m = i < wi ? 0xffffffffu : (i == wi ? below : 0u);
    s += __popc(w[i] & m);
  }
  return s;
}

struct RunAcc {
  long long sx, sy, sz;
  uint32_t r, g, b, n;
};

__device__ __forceinline__ void flush_run(uint64_t cell, const RunAcc& acc, const uint4* __restrict__ bitmap4,
                                          const uint32_t* __restrict__ group_prefix, unsigned long long* __restrict__ accum) {
  if (cell == kNoCell || acc.n == 0) return;
  unsigned long long* a = accum + (size_t)slot_of_cell(cell, bitmap4, group_prefix) * kAccWords;
  atomicAdd(a + 0, (unsigned long long)acc.sx);
  atomicAdd(a + 1, (unsigned long long)acc.sy);
  atomicAdd(a + 2, (unsigned long long)acc.sz);
  atomicAdd(a + 3, ((unsigned long long)acc.r << 32) | acc.g);
  atomicAdd(a + 4, ((unsigned long long)acc.b << 32) | acc.n);
}

constexpr int kAccPX = 8;  // consecutive points per thread

template <bool kVec>
__global__ void __launch_bounds__(256)
accumulate_points_kernel(GridDev g, float rv, int64_t n, const float* __restrict__ xyz, const uint8_t* __restrict__ rgb,
                         const uint8_t* __restrict__ votes, int thr, const uint4* __restrict__ bitmap4,
                         const uint32_t* __restrict__ group_prefix, unsigned long long* __restrict__ accum) {
  const int64_t base = ((int64_t)blockIdx.x * 256 + threadIdx.x) * kAccPX;
  if (base >= n) return;
  float p[kAccPX * 3];
  uint8_t c[kAccPX * 3];
  uint8_t v[kAccPX];
  const int m = (int)min((int64_t)kAccPX, n - base);
  if (kVec && m == kAccPX) {
    const float4* x4 = reinterpret_cast<const float4*>(xyz + base * 3);
#pragma unroll
    for (int q = 0; q < kAccPX * 3 / 4; ++q) {
      const float4 t = __ldcs(x4 + q);
      p[q * 4 + 0] = t.x, p[q * 4 + 1] = t.y, p[q * 4 + 2] = t.z, p[q * 4 + 3] = t.w;
    }
    const uint2* c2 = reinterpret_cast<const uint2*>(rgb + base * 3);
#pragma unroll
    for (int q = 0; q < kAccPX * 3 / 8; ++q) {
      const uint2 t = __ldcs(c2 + q);
#pragma unroll
      for (int e = 0; e < 4; ++e) c[q * 8 + e] = (uint8_t)(t.x >> (8 * e)), c[q * 8 + 4 + e] = (uint8_t)(t.y >> (8 * e));
    }
    if (votes != nullptr) {
      const uint2 t = __ldcs(reinterpret_cast<const uint2*>(votes + base));
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = (uint8_t)(t.x >> (8 * e)), v[4 + e] = (uint8_t)(t.y >> (8 * e));
    }
  } else {
#pragma unroll
    for (int j = 0; j < kAccPX; ++j) {
      const bool in = j < m;
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        p[j * 3 + e] = in ? __ldg(xyz + (base + j) * 3 + e) : 0.f;
        c[j * 3 + e] = in ? __ldg(rgb + (base + j) * 3 + e) : (uint8_t)0;
      }
      v[j] = (in && votes != nullptr) ? __ldg(votes + base + j) : (uint8_t)0;
    }
  }
  uint64_t cur = kNoCell;
  RunAcc acc = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < kAccPX; ++j) {
    const bool take = j < m && (votes == nullptr || (int)v[j] < thr);
    uint32_t kx = 0, ky = 0, kz = 0;
    const uint64_t cell = take ? cell_of_point(g, rv, p[j * 3 + 0], p[j * 3 + 1], p[j * 3 + 2], kx, ky, kz) : kNoCell;
    if (cell != cur) {
      flush_run(cur, acc, bitmap4, group_prefix, accum);
      acc = {0, 0, 0, 0, 0, 0, 0};
      cur = cell;
    }
    if (cell != kNoCell) {
      // p - centre is exact in float32 for points inside the voxel
      acc.sx += voxel_offset_fix(p[j * 3 + 0], voxel_centre(g.ox, kx, g.voxel), g.voxel);
      acc.sy += voxel_offset_fix(p[j * 3 + 1], voxel_centre(g.oy, ky, g.voxel), g.voxel);
      acc.sz += voxel_offset_fix(p[j * 3 + 2], voxel_centre(g.oz, kz, g.voxel), g.voxel);
      acc.r += c[j * 3 + 0];
      acc.g += c[j * 3 + 1];
      acc.b += c[j * 3 + 2];
      acc.n += 1;
    }
  }
  flush_run(cur, acc, bitmap4, group_prefix, accum);
}

__global__ void __launch_bounds__(256)
accumulate_records_kernel(GridDev g, int64_t n, const uint64_t* __restrict__ keys, const long long* __restrict__ in_sums,
                          const uint32_t* __restrict__ in_rgb, const int32_t* __restrict__ in_count,
                          const uint4* __restrict__ bitmap4, const uint32_t* __restrict__ group_prefix,
                          unsigned long long* __restrict__ accum) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint64_t cell = cell_of_key(g, __ldg(keys + i));
  RunAcc acc;
  acc.sx = in_sums[i * 3 + 0], acc.sy = in_sums[i * 3 + 1], acc.sz = in_sums[i * 3 + 2];
  acc.r = in_rgb[i * 3 + 0], acc.g = in_rgb[i * 3 + 1], acc.b = in_rgb[i * 3 + 2];
  acc.n = (uint32_t)in_count[i];
  flush_run(cell, acc, bitmap4, group_prefix, accum);
}

// One thread per voxel.  The colour fields are 32 bits wide: a voxel with 2^24 or more points could
// have overflowed them, which is reported (counts_out[0] = -1) instead of returned as a wrong colour.
template <bool kPartialOut>
__global__ void __launch_bounds__(256)
finalize_kernel(GridDev g, const unsigned long long* __restrict__ accum, const uint64_t* __restrict__ keys,
                int64_t* __restrict__ counts, float* __restrict__ out_xyz, uint8_t* __restrict__ out_rgb,
                int32_t* __restrict__ out_count, long long* __restrict__ part_sums, uint32_t* __restrict__ part_rgb) {
  const int64_t mv = counts[1];
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < mv; r += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long* a = accum + (size_t)r * kAccWords;
    const long long sx = (long long)a[0], sy = (long long)a[1], sz = (long long)a[2];
    const unsigned long long rg = a[3], bn = a[4];
    const uint32_t cnt = (uint32_t)bn;
    if (cnt >= (1u << 24)) counts[0] = -1;
    out_count[r] = (int32_t)cnt;
    if (kPartialOut) {
      part_sums[r * 3 + 0] = sx;
      part_sums[r * 3 + 1] = sy;
      part_sums[r * 3 + 2] = sz;
      part_rgb[r * 3 + 0] = (uint32_t)(rg >> 32);
      part_rgb[r * 3 + 1] = (uint32_t)rg;
      part_rgb[r * 3 + 2] = (uint32_t)(bn >> 32);
    } else {
      const uint64_t key = keys[r];
      const uint32_t kx = (uint32_t)(key & 0x1fffff), ky = (uint32_t)((key >> 21) & 0x1fffff), kz = (uint32_t)((key >> 42) & 0x1fffff);
      finalize_voxel(g, voxel_centre(g.ox, kx, g.voxel), voxel_centre(g.oy, ky, g.voxel), voxel_centre(g.oz, kz, g.voxel), sx, sy,
                     sz, rg >> 32, rg & 0xffffffffull, bn >> 32, (long long)cnt, out_xyz + r * 3, out_rgb + r * 3);
    }
  }
}

__global__ void canonical_key_kernel(GridDev g, int64_t n, const float* __restrict__ xyz, uint64_t* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = xyz[i * 3 + 0], y = xyz[i * 3 + 1], z = xyz[i * 3 + 2];
  const int64_t kx = (int64_t)floorf(__fdiv_rn(__fsub_rn(x, g.ox), g.voxel));
  const int64_t ky = (int64_t)floorf(__fdiv_rn(__fsub_rn(y, g.oy), g.voxel));
  const int64_t kz = (int64_t)floorf(__fdiv_rn(__fsub_rn(z, g.oz), g.voxel));
  const bool ok = kx >= 0 && ky >= 0 && kz >= 0 && kx < (1 << 21) && ky < (1 << 21) && kz < (1 << 21);
  keys[i] = ok ? ((uint64_t)kx | ((uint64_t)ky << 21) | ((uint64_t)kz << 42)) : ~0ull;
}

// ---- host side ---------------------------------------------------------------------------------
struct DenseLayout {
  uint64_t cells, groups, tiles;
  size_t bitmap, bitmap_bytes, prefix, tile_sums, accum, total;
};

static uint64_t grid_cells(const GridDev& g) { return (uint64_t)g.nx * (uint64_t)g.ny * (uint64_t)g.nz; }
static bool use_dense(const GridDev& g) { return grid_cells(g) <= kDenseMaxCells; }

static void dense_layout(const GridDev& g, int64_t n, DenseLayout* L) {
  L->cells = grid_cells(g);
  L->groups = (L->cells + kGroupBits - 1) / kGroupBits;
  L->tiles = (L->groups + kTileGroups - 1) / kTileGroups;
  const uint64_t max_vox = (uint64_t)n < L->cells ? (uint64_t)n : L->cells;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (size_t)align_up((int64_t)bytes, 256);
    return o;
  };
  L->bitmap_bytes = (size_t)L->groups * (kGroupBits / 8);
  L->bitmap = take(L->bitmap_bytes);
  L->prefix = take((size_t)L->groups * 4);
  L->tile_sums = take((size_t)(L->tiles + 1) * 4);
  L->accum = take((size_t)max_vox * kAccWords * 8);
  L->total = off + 256;
}

struct DenseSource {
  // points
  const float* xyz = nullptr;
  const uint8_t* rgb = nullptr;
  const uint8_t* votes = nullptr;
  int thr = 0;
  // records
  const uint64_t* rec_keys = nullptr;
  const long long* rec_sums = nullptr;
  const uint32_t* rec_rgb = nullptr;
  const int32_t* rec_count = nullptr;
};

static int dense_fuse(const GridDev& g, int64_t n, const DenseSource& src, uint64_t* out_keys, float* out_xyz, uint8_t* out_rgb,
                      int32_t* out_count, long long* part_sums, uint32_t* part_rgb, int64_t* counts_out, void* workspace,
                      int64_t workspace_bytes, cudaStream_t st) {
  DenseLayout L;
  dense_layout(g, n, &L);
  if ((int64_t)L.total > workspace_bytes) {
    set_error("fuse workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)L.total);
    return DDN_ERR_WORKSPACE_TOO_SMALL;
  }
  char* base = reinterpret_cast<char*>(align_up((int64_t)(uintptr_t)workspace, 256));
  uint32_t* bitmap = (uint32_t*)(base + L.bitmap);
  const uint4* bitmap4 = (const uint4*)bitmap;
  uint32_t* prefix = (uint32_t*)(base + L.prefix);
  uint32_t* tile_sums = (uint32_t*)(base + L.tile_sums);
  unsigned long long* accum = (unsigned long long*)(base + L.accum);
  const float rv = 1.0f / g.voxel;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  const bool points = src.rec_keys == nullptr;

  DDN_TRY(check_cuda(cudaMemsetAsync(bitmap, 0, L.bitmap_bytes, st), "memset bitmap"));
  DDN_TRY(check_cuda(cudaMemsetAsync(counts_out, 0, 16, st), "memset counts"));
  if (points)
    mark_points_kernel<<<blocks, 256, 0, st>>>(g, rv, n, src.xyz, src.votes, src.thr, bitmap, (unsigned long long*)counts_out);
  else
    mark_records_kernel<<<blocks, 256, 0, st>>>(g, n, src.rec_keys, bitmap, (unsigned long long*)counts_out);
  DDN_TRY(after_launch("mark_kernel"));
  tile_count_kernel<<<(unsigned)L.tiles, kScanThreads, 0, st>>>(bitmap4, (uint32_t)L.groups, tile_sums);
  DDN_TRY(after_launch("tile_count_kernel"));
  tile_scan_kernel<<<1, 1024, 0, st>>>(tile_sums, (int)L.tiles, counts_out);
  DDN_TRY(after_launch("tile_scan_kernel"));
  group_prefix_kernel<<<(unsigned)L.tiles, kScanThreads, 0, st>>>(g, bitmap4, (uint32_t)L.groups, tile_sums, prefix, out_keys, accum);
  DDN_TRY(after_launch("group_prefix_kernel"));
  if (points) {
    const unsigned ablocks = (unsigned)((n + 256 * kAccPX - 1) / (256 * kAccPX));
    const bool vec = ((uintptr_t)src.xyz % 16 == 0) && ((uintptr_t)src.rgb % 8 == 0) && (src.votes == nullptr || (uintptr_t)src.votes % 8 == 0);
    if (vec)
      accumulate_points_kernel<true><<<ablocks, 256, 0, st>>>(g, rv, n, src.xyz, src.rgb, src.votes, src.thr, bitmap4, prefix, accum);
    else
      accumulate_points_kernel<false><<<ablocks, 256, 0, st>>>(g, rv, n, src.xyz, src.rgb, src.votes, src.thr, bitmap4, prefix, accum);
  } else {
    accumulate_records_kernel<<<blocks, 256, 0, st>>>(g, n, src.rec_keys, src.rec_sums, src.rec_rgb, src.rec_count, bitmap4, prefix,
                                                      accum);
  }
  DDN_TRY(after_launch("accumulate_kernel"));
  if (part_sums != nullptr)
    finalize_kernel<true><<<kNumSMs * 8, 256, 0, st>>>(g, accum, out_keys, counts_out, out_xyz, out_rgb, out_count, part_sums, part_rgb);
  else
    finalize_kernel<false><<<kNumSMs * 8, 256, 0, st>>>(g, accum, out_keys, counts_out, out_xyz, out_rgb, out_count, nullptr, nullptr);
  return after_launch("finalize_kernel");
}

}  // namespace ddn

extern "C" {

int ddn_fuse_workspace_bytes(const ddn_voxel_grid* grid_host, int64_t n_points, int64_t* bytes_out) {
  using namespace ddn;
  DDN_REQUIRE(bytes_out != nullptr, "null bytes_out");
  DDN_REQUIRE(n_points >= 0 && n_points < (1ll << 31) - 1024, "n_points");
  const int64_t n = n_points > 0 ? n_points : 1;
  if (grid_host != nullptr) {
    GridDev g;
    DDN_TRY(grid_from_host(grid_host, &g));
    if (use_dense(g)) {
      DenseLayout L;
      dense_layout(g, n, &L);
      *bytes_out = (int64_t)L.total;
      return DDN_OK;
    }
  }
  return sort_fuse_workspace_bytes(n, bytes_out);
}

int ddn_voxel_fuse(const ddn_voxel_grid* grid_host, int64_t n_points, const float* xyz, const uint8_t* rgb,
                   const uint8_t* votes, int32_t vote_threshold, uint64_t* out_keys, float* out_xyz,
                   uint8_t* out_rgb, int32_t* out_count, int64_t* counts_out, void* workspace,
                   int64_t workspace_bytes, void* stream) {
  using namespace ddn;
  GridDev g;
  DDN_TRY(grid_from_host(grid_host, &g));
  DDN_REQUIRE(n_points >= 0 && n_points < (1ll << 31) - 1024, "n_points");
  DDN_REQUIRE(counts_out != nullptr, "null counts_out");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_points == 0) return check_cuda(cudaMemsetAsync(counts_out, 0, 16, st), "memset counts");
  DDN_REQUIRE(xyz && rgb && out_keys && out_xyz && out_rgb && out_count && workspace, "null pointer");
  if (!use_dense(g))
    return sort_fuse_points(g, n_points, xyz, rgb, votes, vote_threshold, out_keys, out_xyz, out_rgb, out_count, counts_out,
                            workspace, workspace_bytes, st, nullptr, nullptr);
  DenseSource src;
  src.xyz = xyz, src.rgb = rgb, src.votes = votes, src.thr = vote_threshold;
  return dense_fuse(g, n_points, src, out_keys, out_xyz, out_rgb, out_count, nullptr, nullptr, counts_out, workspace,
                    workspace_bytes, st);
}

int ddn_voxel_partials(const ddn_voxel_grid* grid_host, int64_t n_points, const float* xyz, const uint8_t* rgb,
                       const uint8_t* votes, int32_t vote_threshold, uint64_t* part_keys, int64_t* part_sums,
                       uint32_t* part_rgb, int32_t* part_count, int64_t* counts_out, void* workspace,
                       int64_t workspace_bytes, void* stream) {
  using namespace ddn;
  GridDev g;
  DDN_TRY(grid_from_host(grid_host, &g));
  DDN_REQUIRE(n_points >= 0 && n_points < (1ll << 31) - 1024, "n_points");
  DDN_REQUIRE(counts_out != nullptr, "null counts_out");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_points == 0) return check_cuda(cudaMemsetAsync(counts_out, 0, 16, st), "memset counts");
  DDN_REQUIRE(xyz && rgb && part_keys && part_sums && part_rgb && part_count && workspace, "null pointer");
  if (!use_dense(g))
    return sort_fuse_points(g, n_points, xyz, rgb, votes, vote_threshold, part_keys, nullptr, nullptr, part_count, counts_out,
                            workspace, workspace_bytes, st, (long long*)part_sums, part_rgb);
  DenseSource src;
  src.xyz = xyz, src.rgb = rgb, src.votes = votes, src.thr = vote_threshold;
  return dense_fuse(g, n_points, src, part_keys, nullptr, nullptr, part_count, (long long*)part_sums, part_rgb, counts_out,
                    workspace, workspace_bytes, st);
}

int ddn_voxel_merge(const ddn_voxel_grid* grid_host, int64_t n_records, const uint64_t* part_keys,
                    const int64_t* part_sums, const uint32_t* part_rgb, const int32_t* part_count,
                    uint64_t* out_keys, float* out_xyz, uint8_t* out_rgb, int32_t* out_count, int64_t* counts_out,
                    void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace ddn;
  GridDev g;
  DDN_TRY(grid_from_host(grid_host, &g));
  DDN_REQUIRE(n_records >= 0 && n_records < (1ll << 31) - 1024, "n_records");
  DDN_REQUIRE(counts_out != nullptr, "null counts_out");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_records == 0) return check_cuda(cudaMemsetAsync(counts_out, 0, 16, st), "memset counts");
  DDN_REQUIRE(part_keys && part_sums && part_rgb && part_count && out_keys && out_xyz && out_rgb && out_count && workspace,
              "null pointer");
  if (!use_dense(g))
    return sort_merge_records(g, n_records, part_keys, (const long long*)part_sums, part_rgb, part_count, out_keys, out_xyz,
                              out_rgb, out_count, counts_out, workspace, workspace_bytes, st);
  DenseSource src;
  src.rec_keys = part_keys, src.rec_sums = (const long long*)part_sums, src.rec_rgb = part_rgb, src.rec_count = part_count;
  return dense_fuse(g, n_records, src, out_keys, out_xyz, out_rgb, out_count, nullptr, nullptr, counts_out, workspace,
                    workspace_bytes, st);
}

int ddn_voxel_keys(const ddn_voxel_grid* grid_host, int64_t n_points, const float* xyz, uint64_t* keys, void* stream) {
  using namespace ddn;
  DDN_REQUIRE(grid_host != nullptr && grid_host->voxel > 0.f, "grid");
  DDN_REQUIRE(n_points >= 0, "n_points");
  if (n_points == 0) return DDN_OK;
  DDN_REQUIRE(xyz && keys, "null pointer");
  GridDev g;
  g.voxel = grid_host->voxel;
  g.ox = grid_host->origin[0];
  g.oy = grid_host->origin[1];
  g.oz = grid_host->origin[2];
  g.bx = g.by = g.bz = 21;
  g.nx = g.ny = g.nz = 1 << 21;
  canonical_key_kernel<<<(unsigned)((n_points + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, n_points, xyz, keys);
  return after_launch("canonical_key_kernel");
}

}  // extern "C"
