// Stage 4: voxel-grid fusion (new capability; the reference only concatenates
// points, scripts/test.py:353-359).  DENSE-RANK PATH - no sort.
//
// The voxel grid of a scene is small enough for one occupancy BIT per cell (cfg 2: about 2400 x 2400 x 600
// cells = 0.6 GB of occupancy units in 180 GB of HBM).  Bit order == key order (x fastest, then y, then z), so
// the rank of a set bit among all set bits IS the voxel's position in the key-sorted output.  That
// replaces the radix sort of 168 M (key, index) pairs by streaming passes:
//
//   mark        every participating point sets the bit of its cell (RED.OR; a cell equal to the
//               previous point's is skipped)
//   rank        popcount scan over the occupancy: one exclusive prefix per 96-bit unit, stored in the
//               unit's fourth word; the pass also emits the sorted keys
//   accumulate  integer fixed-point sums per voxel with 64-bit RED.ADD.  The L2 atomic units bound this
//               pass, so a warp first merges the points of its 16 x 2 pixel tile that share a cell
//               (MATCH.ANY + shuffles); the group's first lane looks the slot up (ONE 16 B load:
//               96 bits + prefix) and issues the atomics.
//   finalize    one thread per voxel: mean = centre + sum/count, colour = round-half-up
//
// Integer sums make the result independent of the order of points and of how they are split over
// ranks (ddn_voxel_partials / ddn_voxel_merge use the same passes with a different output / input).
// Grids with more than 2^35 cells take the sort path in fuse_sort.cu.
#include <algorithm>

#include "fuse_common.cuh"

namespace ddn {

// Occupancy + rank live in ONE array of 16-byte units: words x, y, z = 96 occupancy bits, word w = the
// exclusive rank prefix of the unit (written by the rank pass).  A slot lookup is a single LDG.128.
constexpr int kUnitBits = 96;
constexpr int kUnitsPerThread = 8;
constexpr int kScanThreads = 256;
constexpr int kTileUnits = kScanThreads * kUnitsPerThread;  // units per CTA in the rank passes (32 KB)
// Ownership granularity of the multi-GPU form: a "tile" of the C ABI is kOwnUnits consecutive units
// (24,576 cells in key order), fine enough to cut a dense z-layer of the grid into balanced shares.
constexpr int kOwnUnits = kScanThreads;
constexpr int kOwnPerScanTile = kTileUnits / kOwnUnits;
constexpr int kAccWords = 5;                                // sx, sy, sz, r:g, b:count (u64 each)
// Partial-sum RECORD exchanged between ranks: {key, sx, sy, sz, r:g, b:count} = DDN_RECORD_WORDS u64.  In
// partial mode the accumulators ARE words 1..5 of the output records (stride 6), so there is no
// finalisation pass and a destination's share of the sorted records is one contiguous slice.
constexpr int kRecWords = DDN_RECORD_WORDS;
static_assert(kRecWords == kAccWords + 1, "record = key + accumulators");
constexpr uint64_t kDenseMaxCells = 1ull << 35;             // 5.7 GB of units
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

__device__ __forceinline__ void set_cell_bit(uint32_t* __restrict__ units, uint64_t cell) {
  const uint32_t w32 = (uint32_t)(cell >> 5);  // word index in a plain bitmap
  const uint32_t unit = w32 / 3u;
  atomicOr(units + (size_t)unit * 4 + (w32 - unit * 3u), 1u << (cell & 31));
}

constexpr int kMarkPX = 4;  // consecutive points per thread

// mark: a thread owns 4 consecutive points; a cell equal to its predecessor (in the thread, or the last
// cell of the previous lane) is not marked again.  Consecutive pixels of a depth map fall into the same
// or neighbouring voxels, so this removes most of the atomics.
template <bool kVec>
__global__ void __launch_bounds__(256)
mark_points_kernel(GridDev g, float rv, int64_t n, const float* __restrict__ xyz, const uint8_t* __restrict__ votes, int thr,
                   uint32_t* __restrict__ units, unsigned long long* __restrict__ n_in) {
  __shared__ int s_count;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  const int64_t base = ((int64_t)blockIdx.x * 256 + threadIdx.x) * kMarkPX;
  uint64_t cell[kMarkPX];
  int mine = 0;
  bool take[kMarkPX];
  if (kVec && base + kMarkPX <= n) {
    uint32_t vv = 0;
    if (votes != nullptr) vv = __ldcs(reinterpret_cast<const uint32_t*>(votes + base));
#pragma unroll
    for (int j = 0; j < kMarkPX; ++j) take[j] = votes == nullptr || (int)((vv >> (8 * j)) & 0xff) < thr;
    if (take[0] | take[1] | take[2] | take[3]) {
      const float4* x4 = reinterpret_cast<const float4*>(xyz + base * 3);
      const float4 a = __ldcs(x4), b = __ldcs(x4 + 1), c = __ldcs(x4 + 2);
      const float p[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
      uint32_t kx, ky, kz;
#pragma unroll
      for (int j = 0; j < kMarkPX; ++j)
        cell[j] = take[j] ? cell_of_point(g, rv, p[j * 3], p[j * 3 + 1], p[j * 3 + 2], kx, ky, kz) : kNoCell;
    } else {
#pragma unroll
      for (int j = 0; j < kMarkPX; ++j) cell[j] = kNoCell;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kMarkPX; ++j) {
      const int64_t i = base + j;
      cell[j] = kNoCell;
      if (i < n && (votes == nullptr || (int)__ldg(votes + i) < thr)) {
        uint32_t kx, ky, kz;
        cell[j] = cell_of_point(g, rv, __ldg(xyz + i * 3 + 0), __ldg(xyz + i * 3 + 1), __ldg(xyz + i * 3 + 2), kx, ky, kz);
      }
    }
  }
  uint64_t prev = __shfl_up_sync(0xffffffffu, cell[kMarkPX - 1], 1);
  if ((threadIdx.x & 31) == 0) prev = kNoCell;
#pragma unroll
  for (int j = 0; j < kMarkPX; ++j) {
    if (cell[j] != kNoCell) {
      ++mine;
      if (cell[j] != prev) set_cell_bit(units, cell[j]);
    }
    prev = cell[j];
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_count, mine);
  __syncthreads();
  if (threadIdx.x == 0 && s_count) atomicAdd(n_in, (unsigned long long)s_count);
}

__global__ void __launch_bounds__(256)
mark_records_kernel(GridDev g, int64_t n, const uint64_t* __restrict__ records, uint32_t* __restrict__ units,
                    uint64_t cell_begin, uint64_t cell_end, unsigned long long* __restrict__ n_in) {
  __shared__ int s_count;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  uint64_t cell = i < n ? cell_of_key(g, __ldg(records + i * kRecWords)) : kNoCell;
  if (cell < cell_begin || cell >= cell_end) cell = kNoCell;  // not owned by this call's tile range
  if (cell != kNoCell) set_cell_bit(units, cell);
  const int c = __popc(__ballot_sync(0xffffffffu, cell != kNoCell));
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_count, c);
  __syncthreads();
  if (threadIdx.x == 0 && s_count) atomicAdd(n_in, (unsigned long long)s_count);
}

// ---- rank: popcount scan over the units --------------------------------------------------------
__device__ __forceinline__ int popc3(const uint4& u) { return __popc(u.x) + __popc(u.y) + __popc(u.z); }

__global__ void __launch_bounds__(kScanThreads)
tile_count_kernel(const uint4* __restrict__ units, uint32_t n_units, uint32_t tile_begin, uint32_t* __restrict__ tile_sums) {
  __shared__ int s_warp[kScanThreads / 32];
  const uint32_t base = (tile_begin + blockIdx.x) * kTileUnits;
  int sum = 0;
#pragma unroll
  for (int j = 0; j < kUnitsPerThread; ++j) {
    const uint32_t ui = base + j * kScanThreads + threadIdx.x;
    if (ui < n_units) sum += popc3(__ldg(units + ui));
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
  if (threadIdx.x == 0) {
    counts_out[1] = (int64_t)s_carry;
    tile_sums[tiles] = s_carry;  // exclusive prefix with the total appended: [tiles + 1] entries
  }
}

// Per unit: exclusive rank prefix -> word w of the unit.  Per set bit: canonical key of the cell ->
// keys[slot].  The cell coordinates are decoded once per non-empty unit and then stepped along x.
__global__ void __launch_bounds__(kScanThreads)
unit_prefix_kernel(GridDev g, uint4* __restrict__ units, uint32_t n_units, uint32_t tile_begin,
                   const uint32_t* __restrict__ tile_excl, uint64_t* __restrict__ keys, int key_stride,
                   uint32_t* __restrict__ own_prefix, uint32_t own_tiles) {
  __shared__ uint32_t s_warp[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = (tile_begin + blockIdx.x) * kTileUnits;
  uint32_t carry = tile_excl[blockIdx.x];
  const uint32_t own0 = (tile_begin + blockIdx.x) * kOwnPerScanTile;
  // empty tile (most of the grid is): its units keep the zero prefix word of the memset and are never
  // looked up, nothing to emit
  if (tile_excl[blockIdx.x + 1] == carry) {
    if (own_prefix != nullptr && threadIdx.x <= kOwnPerScanTile && own0 + threadIdx.x <= own_tiles &&
        (threadIdx.x < kOwnPerScanTile || blockIdx.x == gridDim.x - 1))
      own_prefix[own0 + threadIdx.x] = carry;
    return;
  }
  const uint64_t nxy = (uint64_t)g.nx * (uint64_t)g.ny;
#pragma unroll 1
  for (int j = 0; j < kUnitsPerThread; ++j) {
    const uint32_t ui = base + j * kScanThreads + threadIdx.x;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (ui < n_units) u = __ldg(units + ui);
    const uint32_t cnt = (uint32_t)popc3(u);
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
    if (own_prefix != nullptr && threadIdx.x == 0 && own0 + j <= own_tiles) own_prefix[own0 + j] = carry;
    carry += total;
    if (ui < n_units) units[ui].w = slot;
    if (cnt) {
      const uint64_t cell0 = (uint64_t)ui * kUnitBits;
      uint32_t kz = (uint32_t)(cell0 / nxy);
      const uint64_t rem = cell0 - (uint64_t)kz * nxy;
      uint32_t ky = (uint32_t)(rem / (uint64_t)g.nx);
      const uint32_t kx0 = (uint32_t)(rem - (uint64_t)ky * (uint64_t)g.nx);
      const uint32_t w[3] = {u.x, u.y, u.z};
      uint32_t row_off = 0;  // bits of this unit that belong to earlier rows
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        uint32_t bits = w[i];
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          uint32_t kx = kx0 + (uint32_t)(i * 32 + b) - row_off;
          while (kx >= (uint32_t)g.nx) {  // the unit straddles a row end
            kx -= (uint32_t)g.nx;
            row_off += (uint32_t)g.nx;
            if (++ky >= (uint32_t)g.ny) ky = 0, ++kz;
          }
          keys[(size_t)(slot++) * key_stride] = (uint64_t)kx | ((uint64_t)ky << 21) | ((uint64_t)kz << 42);
        }
      }
    }
  }
  if (own_prefix != nullptr && threadIdx.x == 0 && blockIdx.x == gridDim.x - 1 && own0 + kOwnPerScanTile <= own_tiles)
    own_prefix[own0 + kOwnPerScanTile] = carry;  // total, when the ownership tiles end exactly at this scan tile
}

// accumulators of the counts[1] voxels -> 0 (device-side count, no host round trip)
__global__ void __launch_bounds__(256)
zero_accum_kernel(ulonglong2* __restrict__ accum2, const int64_t* __restrict__ counts, int words_per_voxel) {
  const int64_t n2 = (counts[1] * words_per_voxel + 1) / 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x)
    accum2[i] = make_ulonglong2(0ull, 0ull);
}

__device__ __forceinline__ uint32_t slot_of_cell(uint64_t cell, const uint4* __restrict__ units) {
  const uint32_t w32 = (uint32_t)(cell >> 5);
  const uint32_t unit = w32 / 3u;
  const int wi = (int)(w32 - unit * 3u);
  const uint32_t below = (1u << (cell & 31)) - 1u;
  const uint4 u = __ldg(units + unit);
  uint32_t s = u.w;
  s += __popc(u.x & (wi > 0 ? 0xffffffffu : below));
  s += wi > 0 ? __popc(u.y & (wi > 1 ? 0xffffffffu : below)) : 0;
  s += wi > 1 ? __popc(u.z & below) : 0;
  return s;
}

// accumulate: one point per lane, aggregated ACROSS THE WARP before touching memory.  With a row length
// (points are pixels of [rows, row_len] images) a warp covers a 16 x 2 pixel tile, otherwise 32
// consecutive points (the tile shape is kAccTileW).  Lanes whose points fall into the same cell are found with MATCH.ANY; every lane
// walks its peer mask with shuffles (32-bit partial sums: at most 32 points of |offset| <= 2^19), and
// the lowest lane of each group looks the slot up and issues the five 64-bit REDs.  The L2 atomic units
// are the bound of this pass, so points per RED group is what matters: ~2 at cfg 2.
#ifndef DDN_ACC_TPW
#define DDN_ACC_TPW 8
#endif
constexpr int kAccTilesPerWarp = DDN_ACC_TPW;
constexpr int kAccTileW = 16;  // pixel tile of a warp: 16 x 2 (measured at cfg 2: 4 x 8 4.46, 8 x 4 4.35, 16 x 2 4.34, 32 x 1 4.54 ms)

template <int kTW>  // tile width in pixels (tile = kTW x 32/kTW); 0 = no row structure, 32 consecutive points
__global__ void __launch_bounds__(256)
accumulate_points_kernel(GridDev g, float rv, int64_t n, int row_len, const float* __restrict__ xyz,
                         const uint8_t* __restrict__ rgb, const uint8_t* __restrict__ votes, int thr,
                         const uint4* __restrict__ units, unsigned long long* __restrict__ accum, int stride) {
  const int lane = threadIdx.x & 31;
  const float fix_scale = voxel_fix_scale(g.voxel);
  // n < 2^31, so tile indices fit 32 bits
  const uint32_t warp = blockIdx.x * 8u + (threadIdx.x >> 5);
  constexpr bool kTiled = kTW > 0;
  constexpr int kTWs = kTiled ? kTW : 32, kTH = 32 / kTWs;
  uint32_t tiles_x = 0, n_tiles;
  if (kTiled) {
    tiles_x = ((uint32_t)row_len + kTWs - 1u) / kTWs;
    const uint32_t n_rows = (uint32_t)((n + row_len - 1) / row_len);
    n_tiles = tiles_x * ((n_rows + kTH - 1u) / kTH);
  } else {
    n_tiles = (uint32_t)((n + 31) / 32);
  }
  const int dx = lane % kTWs, dy = lane / kTWs;
  const uint32_t t_begin = warp * kAccTilesPerWarp;
  if (t_begin >= n_tiles) return;
  const uint32_t t_end = min(t_begin + (uint32_t)kAccTilesPerWarp, n_tiles);
  uint32_t ty = 0, tx = 0;
  if (kTiled) {
    ty = t_begin / tiles_x;
    tx = t_begin - ty * tiles_x;
  }
  // The point of the NEXT tile is loaded (vote, position and colour at once - no dependent round trips)
  // before the current tile is processed, so its latency hides behind the aggregation and the atomics.
  struct In {
    float x, y, z;
    uint32_t rg, bb;
    bool take;
  };
  auto load_tile = [&](uint32_t t) -> In {
    int64_t i;
    bool in;
    if (kTiled) {
      const int x = (int)tx * kTWs + dx;
      i = (int64_t)(ty * kTH + dy) * row_len + x;
      in = x < row_len && i < n;
      if (++tx == tiles_x) tx = 0, ++ty;
    } else {
      i = (int64_t)t * 32 + lane;
      in = i < n;
    }
    In r = {0.f, 0.f, 0.f, 0u, 0u, false};
    if (in) {
      const int v = votes != nullptr ? (int)__ldg(votes + i) : 0;
      r.x = __ldg(xyz + i * 3 + 0), r.y = __ldg(xyz + i * 3 + 1), r.z = __ldg(xyz + i * 3 + 2);
      r.rg = ((uint32_t)__ldg(rgb + i * 3 + 0) << 16) | (uint32_t)__ldg(rgb + i * 3 + 1);
      r.bb = (uint32_t)__ldg(rgb + i * 3 + 2);
      r.take = v < thr || votes == nullptr;
    }
    return r;
  };
  In nxt = load_tile(t_begin);
#pragma unroll 1
  for (uint32_t t = t_begin; t < t_end; ++t) {
    const In cur = nxt;
    if (t + 1 < t_end) nxt = load_tile(t + 1);
    uint64_t cell = kNoCell;
    int ox = 0, oy = 0, oz = 0;
    const uint32_t rg = cur.rg, bb = cur.bb;
    if (cur.take) {
      uint32_t kx, ky, kz;
      cell = cell_of_point(g, rv, cur.x, cur.y, cur.z, kx, ky, kz);
      if (cell != kNoCell) {
        // p - centre is exact in float32 for points inside the voxel
        ox = voxel_offset_fix(cur.x, voxel_centre(g.ox, kx, g.voxel), fix_scale);
        oy = voxel_offset_fix(cur.y, voxel_centre(g.oy, ky, g.voxel), fix_scale);
        oz = voxel_offset_fix(cur.z, voxel_centre(g.oz, kz, g.voxel), fix_scale);
      }
    }
    const bool valid = cell != kNoCell;
    uint32_t peers = __match_any_sync(0xffffffffu, cell);
    if (!valid) peers = 0;
    const bool leader = valid && (__ffs(peers) - 1) == lane;
    const int iters = __reduce_max_sync(0xffffffffu, __popc(peers));
    int sx = 0, sy = 0, sz = 0;
    uint32_t srg = 0, sb = 0;
    uint32_t rest = peers;
#pragma unroll 1
    for (int k = 0; k < iters; ++k) {
      const bool has = rest != 0;
      const int src = has ? __ffs(rest) - 1 : lane;
      rest &= rest - 1;
      const int ax = __shfl_sync(0xffffffffu, ox, src), ay = __shfl_sync(0xffffffffu, oy, src);
      const int az = __shfl_sync(0xffffffffu, oz, src);
      const uint32_t arg = __shfl_sync(0xffffffffu, rg, src), ab = __shfl_sync(0xffffffffu, bb, src);
      if (has) sx += ax, sy += ay, sz += az, srg += arg, sb += ab;
    }
    if (leader) {
      unsigned long long* a = accum + (size_t)slot_of_cell(cell, units) * stride;
      atomicAdd(a + 0, (unsigned long long)(long long)sx);
      atomicAdd(a + 1, (unsigned long long)(long long)sy);
      atomicAdd(a + 2, (unsigned long long)(long long)sz);
      atomicAdd(a + 3, ((unsigned long long)(srg >> 16) << 32) | (srg & 0xffffu));
      atomicAdd(a + 4, ((unsigned long long)sb << 32) | (unsigned)__popc(peers));
    }
  }
}

__global__ void __launch_bounds__(256)
accumulate_records_kernel(GridDev g, int64_t n, const unsigned long long* __restrict__ records, const uint4* __restrict__ units,
                          uint64_t cell_begin, uint64_t cell_end, unsigned long long* __restrict__ accum) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const unsigned long long* r = records + i * kRecWords;
  const uint64_t cell = cell_of_key(g, __ldg(r));
  if (cell == kNoCell || cell < cell_begin || cell >= cell_end) return;
  unsigned long long* a = accum + (size_t)slot_of_cell(cell, units) * kAccWords;
#pragma unroll
  for (int q = 0; q < kAccWords; ++q) atomicAdd(a + q, __ldg(r + 1 + q));
}

// One thread per voxel.  The colour fields are 32 bits wide: a voxel with 2^24 or more points could
// have overflowed them, which is reported (counts_out[0] = -1) instead of returned as a wrong colour.
__global__ void __launch_bounds__(256)
finalize_kernel(GridDev g, const unsigned long long* __restrict__ accum, const uint64_t* __restrict__ keys,
                int64_t* __restrict__ counts, float* __restrict__ out_xyz, uint8_t* __restrict__ out_rgb,
                int32_t* __restrict__ out_count) {
  const int64_t mv = counts[1];
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < mv; r += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long* a = accum + (size_t)r * kAccWords;
    const long long sx = (long long)a[0], sy = (long long)a[1], sz = (long long)a[2];
    const unsigned long long rg = a[3], bn = a[4];
    const uint32_t cnt = (uint32_t)bn;
    if (cnt >= (1u << 24)) counts[0] = -1;
    out_count[r] = (int32_t)cnt;
    const uint64_t key = keys[r];
    const uint32_t kx = (uint32_t)(key & 0x1fffff), ky = (uint32_t)((key >> 21) & 0x1fffff), kz = (uint32_t)((key >> 42) & 0x1fffff);
    finalize_voxel(g, voxel_centre(g.ox, kx, g.voxel), voxel_centre(g.oy, ky, g.voxel), voxel_centre(g.oz, kz, g.voxel), sx, sy, sz,
                   rg >> 32, rg & 0xffffffffull, bn >> 32, (long long)cnt, out_xyz + r * 3, out_rgb + r * 3);
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
  uint64_t cells, n_units, tiles, own_tiles;
  size_t units, units_bytes, tile_sums, accum, total;
};

static uint64_t grid_cells(const GridDev& g) { return (uint64_t)g.nx * (uint64_t)g.ny * (uint64_t)g.nz; }
static bool use_dense(const GridDev& g) { return grid_cells(g) <= kDenseMaxCells; }

static void dense_layout(const GridDev& g, int64_t n, DenseLayout* L) {  // sized for the non-partial form
  L->cells = grid_cells(g);
  L->n_units = (L->cells + kUnitBits - 1) / kUnitBits;
  L->tiles = (L->n_units + kTileUnits - 1) / kTileUnits;
  L->own_tiles = (L->n_units + kOwnUnits - 1) / kOwnUnits;
  const uint64_t max_vox = (uint64_t)n < L->cells ? (uint64_t)n : L->cells;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (size_t)align_up((int64_t)bytes, 256);
    return o;
  };
  L->units_bytes = (size_t)L->n_units * 16;
  L->units = take(L->units_bytes);
  L->tile_sums = take((size_t)(L->tiles + 1) * 4);
  L->accum = take((size_t)max_vox * kAccWords * 8 + 16);
  L->total = off + 256;
}

struct DenseSource {
  // points
  const float* xyz = nullptr;
  const uint8_t* rgb = nullptr;
  const uint8_t* votes = nullptr;
  int thr = 0;
  int64_t row_len = 0;
  // records
  const unsigned long long* records = nullptr;
};

// records_out != nullptr: partial mode (output = records, no finalisation); else final voxels.
// tile range [tile_begin, tile_end) (records source only; 0, 0 = whole grid): only that slice of the occupancy
// array is cleared / scanned and records outside it are ignored - an owner rank merges its key range at a
// cost proportional to its share of the grid.  tile_prefix_out (optional): [tiles + 1] exclusive prefix of
// the voxel count per tile, i.e. where each tile's records start in the sorted output.
static int dense_fuse(const GridDev& g, int64_t n, const DenseSource& src, uint64_t* out_keys, float* out_xyz, uint8_t* out_rgb,
                      int32_t* out_count, unsigned long long* records_out, int64_t* counts_out, void* workspace,
                      int64_t workspace_bytes, cudaStream_t st, int64_t tile_begin = 0, int64_t tile_end = 0,
                      uint32_t* tile_prefix_out = nullptr) {
  DenseLayout L;
  dense_layout(g, n, &L);
  // ownership tiles -> owned cell range and the scan tiles that enclose it
  if (tile_end <= 0) tile_begin = 0, tile_end = (int64_t)L.own_tiles;
  DDN_REQUIRE(tile_begin >= 0 && tile_begin <= tile_end && tile_end <= (int64_t)L.own_tiles, "tile range");
  const uint64_t cell_begin = (uint64_t)tile_begin * kOwnUnits * kUnitBits;
  const uint64_t cell_end = (uint64_t)tile_end * kOwnUnits * kUnitBits;
  const int64_t own_begin = tile_begin, own_end = tile_end;
  tile_begin = own_begin / kOwnPerScanTile;
  tile_end = (own_end + kOwnPerScanTile - 1) / kOwnPerScanTile;
  const uint32_t n_tiles = own_end > own_begin ? (uint32_t)(tile_end - tile_begin) : 0u;
  if ((int64_t)L.total > workspace_bytes) {
    set_error("fuse workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)L.total);
    return DDN_ERR_WORKSPACE_TOO_SMALL;
  }
  char* base = reinterpret_cast<char*>(align_up((int64_t)(uintptr_t)workspace, 256));
  uint4* units = (uint4*)(base + L.units);
  uint32_t* tile_sums = (uint32_t*)(base + L.tile_sums);
  const bool partial = records_out != nullptr;
  // accumulators: words 1..5 of the output records (partial mode) or a workspace array
  unsigned long long* accum = partial ? records_out + 1 : (unsigned long long*)(base + L.accum);
  const int stride = partial ? kRecWords : kAccWords;
  unsigned long long* zero_base = partial ? records_out : accum;
  uint64_t* keys = partial ? (uint64_t*)records_out : out_keys;
  const float rv = 1.0f / g.voxel;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  const bool points = src.records == nullptr;
  const bool vec = points && ((uintptr_t)src.xyz % 16 == 0) && (src.votes == nullptr || (uintptr_t)src.votes % 4 == 0);
  DDN_REQUIRE((uintptr_t)zero_base % 16 == 0, "record / accumulator buffer must be 16-byte aligned");

  DDN_TRY(check_cuda(cudaMemsetAsync(counts_out, 0, 16, st), "memset counts"));
  if (n_tiles == 0) return DDN_OK;
  {
    const size_t ub = (size_t)tile_begin * kTileUnits, ue = std::min((size_t)tile_end * kTileUnits, (size_t)L.n_units);
    DDN_TRY(check_cuda(cudaMemsetAsync(units + ub, 0, (ue - ub) * 16, st), "memset occupancy"));
  }
  if (points) {
    const unsigned mblocks = (unsigned)((n + 256 * kMarkPX - 1) / (256 * kMarkPX));
    if (vec)
      mark_points_kernel<true><<<mblocks, 256, 0, st>>>(g, rv, n, src.xyz, src.votes, src.thr, (uint32_t*)units,
                                                        (unsigned long long*)counts_out);
    else
      mark_points_kernel<false><<<mblocks, 256, 0, st>>>(g, rv, n, src.xyz, src.votes, src.thr, (uint32_t*)units,
                                                         (unsigned long long*)counts_out);
  } else {
    mark_records_kernel<<<blocks, 256, 0, st>>>(g, n, (const uint64_t*)src.records, (uint32_t*)units, cell_begin, cell_end,
                                                (unsigned long long*)counts_out);
  }
  DDN_TRY(after_launch("mark_kernel"));
  tile_count_kernel<<<n_tiles, kScanThreads, 0, st>>>(units, (uint32_t)L.n_units, (uint32_t)tile_begin, tile_sums);
  DDN_TRY(after_launch("tile_count_kernel"));
  tile_scan_kernel<<<1, 1024, 0, st>>>(tile_sums, (int)n_tiles, counts_out);
  DDN_TRY(after_launch("tile_scan_kernel"));
  zero_accum_kernel<<<kNumSMs * 8, 256, 0, st>>>((ulonglong2*)zero_base, counts_out, stride);
  DDN_TRY(after_launch("zero_accum_kernel"));
  unit_prefix_kernel<<<n_tiles, kScanThreads, 0, st>>>(g, units, (uint32_t)L.n_units, (uint32_t)tile_begin, tile_sums, keys,
                                                       partial ? kRecWords : 1, tile_prefix_out, (uint32_t)L.own_tiles);
  DDN_TRY(after_launch("unit_prefix_kernel"));
  if (points) {
    const bool tiled = src.row_len >= 32 && src.row_len < (1 << 30);
    const int tw = tiled ? kAccTileW : 0;
    const int th = tw ? 32 / tw : 1, twe = tw ? tw : 32;
    const int64_t n_tiles = tiled ? ((src.row_len + twe - 1) / twe) * (((n + src.row_len - 1) / src.row_len + th - 1) / th) : (n + 31) / 32;
    const unsigned ablocks = (unsigned)((n_tiles + 8 * kAccTilesPerWarp - 1) / (8 * kAccTilesPerWarp));
#define DDN_ACC(TW) accumulate_points_kernel<TW><<<ablocks, 256, 0, st>>>(g, rv, n, (int)src.row_len, src.xyz, src.rgb, src.votes, src.thr, units, accum, stride)
    if (tw == kAccTileW) DDN_ACC(kAccTileW);
    else DDN_ACC(0);
#undef DDN_ACC
  } else {
    accumulate_records_kernel<<<blocks, 256, 0, st>>>(g, n, src.records, units, cell_begin, cell_end, accum);
  }
  DDN_TRY(after_launch("accumulate_kernel"));
  if (partial) return DDN_OK;
  finalize_kernel<<<kNumSMs * 8, 256, 0, st>>>(g, accum, out_keys, counts_out, out_xyz, out_rgb, out_count);
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

int ddn_voxel_fuse(const ddn_voxel_grid* grid_host, int64_t n_points, int64_t row_len, const float* xyz, const uint8_t* rgb,
                   const uint8_t* votes, int32_t vote_threshold, uint64_t* out_keys, float* out_xyz,
                   uint8_t* out_rgb, int32_t* out_count, int64_t* counts_out, void* workspace,
                   int64_t workspace_bytes, void* stream) {
  using namespace ddn;
  GridDev g;
  DDN_TRY(grid_from_host(grid_host, &g));
  DDN_REQUIRE(n_points >= 0 && n_points < (1ll << 31) - 1024, "n_points");
  DDN_REQUIRE(row_len >= 0, "row_len");
  DDN_REQUIRE(counts_out != nullptr, "null counts_out");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_points == 0) return check_cuda(cudaMemsetAsync(counts_out, 0, 16, st), "memset counts");
  DDN_REQUIRE(xyz && rgb && out_keys && out_xyz && out_rgb && out_count && workspace, "null pointer");
  if (!use_dense(g))
    return sort_fuse_points(g, n_points, xyz, rgb, votes, vote_threshold, out_keys, out_xyz, out_rgb, out_count, counts_out,
                            workspace, workspace_bytes, st, nullptr);
  DenseSource src;
  src.xyz = xyz, src.rgb = rgb, src.votes = votes, src.thr = vote_threshold, src.row_len = row_len;
  return dense_fuse(g, n_points, src, out_keys, out_xyz, out_rgb, out_count, nullptr, counts_out, workspace, workspace_bytes, st);
}

int ddn_fuse_tile_info(const ddn_voxel_grid* grid_host, int64_t* n_tiles, int64_t* cells_per_tile) {
  using namespace ddn;
  DDN_REQUIRE(n_tiles != nullptr && cells_per_tile != nullptr, "null output");
  GridDev g;
  DDN_TRY(grid_from_host(grid_host, &g));
  *n_tiles = 0;
  *cells_per_tile = (int64_t)kOwnUnits * kUnitBits;
  if (use_dense(g)) {
    DenseLayout L;
    dense_layout(g, 1, &L);
    *n_tiles = (int64_t)L.own_tiles;
  }
  return DDN_OK;
}

int ddn_voxel_partials(const ddn_voxel_grid* grid_host, int64_t n_points, int64_t row_len, const float* xyz, const uint8_t* rgb,
                       const uint8_t* votes, int32_t vote_threshold, uint64_t* records, uint32_t* tile_prefix,
                       int64_t* counts_out, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace ddn;
  GridDev g;
  DDN_TRY(grid_from_host(grid_host, &g));
  DDN_REQUIRE(n_points >= 0 && n_points < (1ll << 31) - 1024, "n_points");
  DDN_REQUIRE(row_len >= 0, "row_len");
  DDN_REQUIRE(counts_out != nullptr, "null counts_out");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_points == 0) {
    if (tile_prefix != nullptr && use_dense(g)) {
      DenseLayout L;
      dense_layout(g, 1, &L);
      DDN_TRY(check_cuda(cudaMemsetAsync(tile_prefix, 0, (size_t)(L.own_tiles + 1) * 4, st), "memset tile prefix"));
    }
    return check_cuda(cudaMemsetAsync(counts_out, 0, 16, st), "memset counts");
  }
  DDN_REQUIRE(xyz && rgb && records && workspace, "null pointer");
  if (!use_dense(g))
    return sort_fuse_points(g, n_points, xyz, rgb, votes, vote_threshold, nullptr, nullptr, nullptr, nullptr, counts_out, workspace,
                            workspace_bytes, st, (unsigned long long*)records);
  DenseSource src;
  src.xyz = xyz, src.rgb = rgb, src.votes = votes, src.thr = vote_threshold, src.row_len = row_len;
  return dense_fuse(g, n_points, src, nullptr, nullptr, nullptr, nullptr, (unsigned long long*)records, counts_out, workspace,
                    workspace_bytes, st, 0, 0, tile_prefix);
}

int ddn_voxel_merge(const ddn_voxel_grid* grid_host, int64_t n_records, const uint64_t* records, int64_t tile_begin,
                    int64_t tile_end, uint64_t* out_keys, float* out_xyz, uint8_t* out_rgb, int32_t* out_count,
                    int64_t* counts_out, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace ddn;
  GridDev g;
  DDN_TRY(grid_from_host(grid_host, &g));
  DDN_REQUIRE(n_records >= 0 && n_records < (1ll << 31) - 1024, "n_records");
  DDN_REQUIRE(counts_out != nullptr, "null counts_out");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_records == 0) return check_cuda(cudaMemsetAsync(counts_out, 0, 16, st), "memset counts");
  DDN_REQUIRE(records && out_keys && out_xyz && out_rgb && out_count && workspace, "null pointer");
  if (!use_dense(g))
    return sort_merge_records(g, n_records, (const unsigned long long*)records, out_keys, out_xyz, out_rgb, out_count, counts_out,
                              workspace, workspace_bytes, st);
  DenseSource src;
  src.records = (const unsigned long long*)records;
  return dense_fuse(g, n_records, src, out_keys, out_xyz, out_rgb, out_count, nullptr, counts_out, workspace, workspace_bytes, st,
                    tile_begin, tile_end);
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
