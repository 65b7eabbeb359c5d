// Shared definitions of stage 4 (voxel-grid fusion): grid description, IEEE-exact voxel coordinates,
// the occupancy-unit layout, fixed-point offsets and the finalisation of one voxel.  Used by the
// dense-rank path (fuse.cu), by K4's fused occupancy marking (filter.cu) and by the sort path kept for
// grids too large for a dense occupancy bitmap (fuse_sort.cu).
#pragma once

#include "common.cuh"

namespace ddn {

// Same layout as ddn_grid_state (include/ddn_b200.h): the grid lives in DEVICE memory, so a step needs no
// host round trip between the bounding box and the fusion passes.  Host-grid entry points fill one on
// the host and pass it by value to a one-thread kernel that stores it.
struct GridDev {
  float voxel, ox, oy, oz;
  int bx, by, bz;     // significant bits per axis (sort path key packing)
  int nx, ny, nz;     // cells per axis; a point outside [0, n) on any axis does not participate
  long long n_units;  // ceil(cells / 96); 0 when status != DDN_GRID_OK
  int status, reserved;
  long long cells;
};
static_assert(sizeof(GridDev) == sizeof(ddn_grid_state), "GridDev mirrors ddn_grid_state");

inline int grid_from_host(const ddn_voxel_grid* h, GridDev* g) {
  DDN_REQUIRE(h != nullptr, "null grid");
  DDN_REQUIRE(h->voxel > 0.f, "voxel size");
  for (int i = 0; i < 3; ++i) DDN_REQUIRE(h->bits[i] >= 1 && h->bits[i] <= 21, "bits per axis must be in [1,21]");
  for (int i = 0; i < 3; ++i)
    DDN_REQUIRE(h->dims[i] >= 0 && h->dims[i] <= (1 << h->bits[i]), "dims per axis must be in [0, 2^bits]");
  g->voxel = h->voxel;
  g->ox = h->origin[0];
  g->oy = h->origin[1];
  g->oz = h->origin[2];
  g->bx = h->bits[0];
  g->by = h->bits[1];
  g->bz = h->bits[2];
  g->nx = h->dims[0] > 0 ? h->dims[0] : (1 << h->bits[0]);
  g->ny = h->dims[1] > 0 ? h->dims[1] : (1 << h->bits[1]);
  g->nz = h->dims[2] > 0 ? h->dims[2] : (1 << h->bits[2]);
  g->cells = (long long)g->nx * (long long)g->ny * (long long)g->nz;
  g->n_units = (g->cells + 95) / 96;
  g->status = DDN_GRID_OK;
  g->reserved = 0;
  return DDN_OK;
}

// k = floor((p - o) / voxel) with IEEE float32 sub and div (bit-exact with the numpy definition,
// x/v and x*(1/v) round differently).  The quotient is first estimated with one multiply
// by the rounded reciprocal; the two can only floor differently when the estimate sits within a few
// ulp of an integer, and only then is the IEEE division issued.
__device__ __forceinline__ float voxel_coord(float p, float o, float voxel, float rvoxel) {
  const float a = __fsub_rn(p, o);
  const float t = __fmul_rn(a, rvoxel);
  const float k = floorf(t);
  const float r = t - k;
  const float eps = fmaxf(fabsf(t), 1.f) * 1e-6f;
  if (r < eps || r > 1.f - eps || !(fabsf(t) < 4194304.f)) return floorf(__fdiv_rn(a, voxel));
  return k;
}

constexpr int kFixShift = 20;  // offsets are stored in units of voxel * 2^-20

__device__ __forceinline__ float voxel_centre(float o, uint32_t k, float voxel) {
  return __fadd_rn(o, __fmul_rn((float)k + 0.5f, voxel));
}

// (p - voxel centre) * fix_scale rounded to nearest, fix_scale = fl(1/voxel) * 2^20: the integer a point
// adds to its voxel's sum (|result| <= 2^19 for a point inside the voxel).  One multiply, no division:
// the value only has to be the SAME everywhere (kernels, ranks, the numpy statement), not a quotient.
__host__ __device__ __forceinline__ float voxel_fix_scale(float voxel) { return (1.0f / voxel) * (float)(1 << kFixShift); }
__device__ __forceinline__ int voxel_offset_fix(float p, float centre, float fix_scale) {
  return __float2int_rn(__fmul_rn(__fsub_rn(p, centre), fix_scale));
}

// mean position = centre + (sum / count) * voxel * 2^-20 (float64, once per voxel); colour = round-half-up
__device__ __forceinline__ void finalize_voxel(const GridDev& g, float cx, float cy, float cz, long long sx, long long sy,
                                               long long sz, unsigned long long sr, unsigned long long sg,
                                               unsigned long long sb, long long cnt, float* __restrict__ oxyz,
                                               uint8_t* __restrict__ orgb) {
  const double inv = (double)g.voxel / ((double)cnt * (double)(1 << kFixShift));
  oxyz[0] = (float)((double)cx + (double)sx * inv);
  oxyz[1] = (float)((double)cy + (double)sy * inv);
  oxyz[2] = (float)((double)cz + (double)sz * inv);
  const unsigned long long c2 = 2ull * (unsigned long long)cnt;
  orgb[0] = (uint8_t)((2 * sr + cnt) / c2);
  orgb[1] = (uint8_t)((2 * sg + cnt) / c2);
  orgb[2] = (uint8_t)((2 * sb + cnt) / c2);
}

// ---- occupancy units (dense-rank path) --------------------------------------------------------------
// Occupancy + rank live in ONE array of 16-byte units: words x, y, z = 96 occupancy bits, word w = the
// exclusive rank prefix of the unit (written by the rank pass).  A slot lookup is a single LDG.128.
constexpr int kUnitBits = 96;
constexpr int kUnitsPerThread = 8;
constexpr int kScanThreads = 256;
constexpr int kTileUnits = kScanThreads * kUnitsPerThread;  // units per "scan tile" (32 KB of units)
// Ownership granularity of the multi-GPU form: a "tile" of the C ABI is kOwnUnits consecutive units
// (24,576 cells in key order), fine enough to cut a dense z-layer of the grid into balanced shares.
constexpr int kOwnUnits = kScanThreads;
constexpr int kOwnPerScanTile = kTileUnits / kOwnUnits;
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

// Sets the occupancy bit of `cell`.  `dirty` (optional): one byte per scan tile, set when a tile receives a
// bit - the rank passes and the clean-up then only touch the tiles a step has actually used.  The flag is
// looked at only when the cell lies in another scan tile than `left` (the cell the caller marked just before,
// kNoCell if none), through the L1-cached path (a stale 0 only costs one redundant store per SM), and written
// only while it still reads 0.  Measured alternatives (cfg 2, on top of a 2.23 ms K4): an uncached (volatile)
// check on every mark +1.0 ms, an unconditional store on every tile change +3.5 ms (same-address stores
// serialise in L2).
__device__ __forceinline__ uint32_t unit_of_cell(uint64_t cell) { return (uint32_t)(cell >> 5) / 3u; }
__device__ __forceinline__ void set_cell_bit(uint32_t* __restrict__ units, uint8_t* __restrict__ dirty, uint64_t cell,
                                             uint64_t left = kNoCell) {
  const uint32_t w32 = (uint32_t)(cell >> 5);  // word index in a plain bitmap
  const uint32_t unit = w32 / 3u;
  atomicOr(units + (size_t)unit * 4 + (w32 - unit * 3u), 1u << (cell & 31));
  if (dirty != nullptr) {
    const uint32_t tile = unit / (uint32_t)kTileUnits;
    if (left == kNoCell || unit_of_cell(left) / (uint32_t)kTileUnits != tile) {
      if (__ldg(dirty + tile) == 0) dirty[tile] = 1;
    }
  }
}

// The device buffers of one fusion session, as the kernels see them (host struct: ddn_fuse_session).
struct FuseDev {
  const GridDev* grid;
  uint32_t* units;
  uint8_t* dirty;                   // may be nullptr
  unsigned long long* counts;       // [0] participating points, [1] voxels
};

// ---- sort path (fuse_sort.cu) -------------------------------------------------------------------
// records != nullptr: partial mode, output = records [.,DDN_RECORD_WORDS] (see include/ddn_b200.h)
int sort_fuse_workspace_bytes(int64_t n_points, int64_t* bytes_out);
int sort_fuse_points(const GridDev& g, int64_t n, const float* xyz, const uint8_t* rgb, const uint8_t* votes, int thr,
                     uint64_t* out_keys, float* out_xyz, uint8_t* out_rgb, int32_t* out_count, int64_t* counts_out,
                     void* workspace, int64_t workspace_bytes, cudaStream_t st, unsigned long long* records);
int sort_merge_records(const GridDev& g, int64_t n, const unsigned long long* records, uint64_t* out_keys, float* out_xyz,
                       uint8_t* out_rgb, int32_t* out_count, int64_t* counts_out, void* workspace, int64_t workspace_bytes,
                       cudaStream_t st);

}  // namespace ddn
