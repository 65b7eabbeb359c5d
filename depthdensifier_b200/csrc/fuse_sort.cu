// Stage 4, SORT PATH: used only when the voxel grid is too large for the dense occupancy bitmap of
// fuse.cu (more than 2^35 cells).  New capability; the reference only concatenates
// points, scripts/test.py:353-359.
//
//   K5 voxel_key_kernel     quantise kept points to voxel coordinates with IEEE float32
//                           sub/div/floor (bit-exact with the numpy definition) and pack them into a
//                           COMPACT key that only spends the bits the bounding box needs, so the
//                           radix sort runs 4 passes over 32-bit keys instead of 8 over 64-bit ones.
//   K6 sort                 (key, point index) pairs, least-significant-digit radix sort.
//   K7 segment_mean_kernel  one run of equal keys = one voxel: integer fixed-point sums of
//                           (p - voxel centre) and of colours, so the result does not depend on the
//                           order of points inside a voxel (deterministic across rank counts).
//
// K6 and the run-length/scan steps call CUB device primitives (header-only, shipped with the CUDA
// toolkit); keys, sums and outputs are this library's own kernels.
#include <cub/cub.cuh>

#include "fuse_common.cuh"

namespace ddn {

constexpr int kKeyThreads = 256;

template <typename KeyT>
__global__ void __launch_bounds__(kKeyThreads)
voxel_key_kernel(GridDev g, int64_t n, const float* __restrict__ xyz, const uint8_t* __restrict__ votes, int thr,
                 KeyT* __restrict__ keys, uint32_t* __restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * kKeyThreads + threadIdx.x;
  if (i >= n) return;
  const KeyT sentinel = (KeyT)1 << (g.bx + g.by + g.bz);
  bool take = votes == nullptr || (int)votes[i] < thr;
  KeyT key = sentinel;
  if (take) {
    const float x = xyz[i * 3 + 0], y = xyz[i * 3 + 1], z = xyz[i * 3 + 2];
    const float fx = floorf(__fdiv_rn(__fsub_rn(x, g.ox), g.voxel));
    const float fy = floorf(__fdiv_rn(__fsub_rn(y, g.oy), g.voxel));
    const float fz = floorf(__fdiv_rn(__fsub_rn(z, g.oz), g.voxel));
    const bool inside = fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)g.nx && fy < (float)g.ny && fz < (float)g.nz;
    if (inside) key = (KeyT)(uint32_t)fx | ((KeyT)(uint32_t)fy << g.bx) | ((KeyT)(uint32_t)fz << (g.bx + g.by));
  }
  keys[i] = key;
  idx[i] = (uint32_t)i;
}

template <typename KeyT>
__global__ void fuse_finalize_kernel(GridDev g, int64_t n, const KeyT* __restrict__ unique_keys,
                                     const int* __restrict__ run_counts, const int* __restrict__ num_runs,
                                     int64_t* __restrict__ counts_out) {
  const int r = *num_runs;
  const KeyT sentinel = (KeyT)1 << (g.bx + g.by + g.bz);
  int64_t m = n, mv = r;
  if (r > 0 && unique_keys[r - 1] == sentinel) {
    mv = r - 1;
    m = n - run_counts[r - 1];
  }
  counts_out[0] = m;
  counts_out[1] = mv;
}

// One thread per run of equal keys.  kPartialOut = false: write the voxel mean (single GPU).
// kPartialOut = true: write the raw integer sums so that another rank can merge them.
template <typename KeyT, bool kPartialOut>
__global__ void __launch_bounds__(256)
segment_mean_kernel(GridDev g, const KeyT* __restrict__ unique_keys, const int* __restrict__ run_counts,
                    const int* __restrict__ run_starts, const int64_t* __restrict__ counts,
                    const uint32_t* __restrict__ sorted_idx, const float* __restrict__ xyz,
                    const uint8_t* __restrict__ rgb, uint64_t* __restrict__ out_keys, float* __restrict__ out_xyz,
                    uint8_t* __restrict__ out_rgb, int32_t* __restrict__ out_count, unsigned long long* __restrict__ records) {
  const int64_t mv = counts[1];
  const float scale = voxel_fix_scale(g.voxel);
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < mv; r += (int64_t)gridDim.x * blockDim.x) {
    const KeyT key = unique_keys[r];
    const uint32_t kx = (uint32_t)(key & (((KeyT)1 << g.bx) - 1));
    const uint32_t ky = (uint32_t)((key >> g.bx) & (((KeyT)1 << g.by) - 1));
    const uint32_t kz = (uint32_t)((key >> (g.bx + g.by)) & (((KeyT)1 << g.bz) - 1));
    // voxel centre in float32; p - centre is exact in float32 for points inside the voxel
    const float cx = voxel_centre(g.ox, kx, g.voxel), cy = voxel_centre(g.oy, ky, g.voxel), cz = voxel_centre(g.oz, kz, g.voxel);
    const int start = run_starts[r], cnt = run_counts[r];
    long long sx = 0, sy = 0, sz = 0;
    unsigned long long sr = 0, sg = 0, sb = 0;
    for (int j = 0; j < cnt; ++j) {
      const size_t i = sorted_idx[start + j];
      const float x = __ldg(xyz + i * 3 + 0), y = __ldg(xyz + i * 3 + 1), z = __ldg(xyz + i * 3 + 2);
      sx += voxel_offset_fix(x, cx, scale);
      sy += voxel_offset_fix(y, cy, scale);
      sz += voxel_offset_fix(z, cz, scale);
      sr += __ldg(rgb + i * 3 + 0);
      sg += __ldg(rgb + i * 3 + 1);
      sb += __ldg(rgb + i * 3 + 2);
    }
    const uint64_t canon = (uint64_t)kx | ((uint64_t)ky << 21) | ((uint64_t)kz << 42);
    if (kPartialOut) {
      unsigned long long* o = records + (size_t)r * DDN_RECORD_WORDS;
      o[0] = canon;
      o[1] = (unsigned long long)sx;
      o[2] = (unsigned long long)sy;
      o[3] = (unsigned long long)sz;
      o[4] = (sr << 32) | sg;
      o[5] = (sb << 32) | (unsigned long long)(unsigned)cnt;
    } else {
      out_keys[r] = canon;
      out_count[r] = cnt;
      finalize_voxel(g, cx, cy, cz, sx, sy, sz, sr, sg, sb, (long long)cnt, out_xyz + r * 3, out_rgb + r * 3);
    }
  }
}

// ---- merge of partial records (multi-GPU owner side) ----------------------------------------------
template <typename KeyT>
__global__ void compact_key_kernel(GridDev g, int64_t n, const unsigned long long* __restrict__ records, KeyT* __restrict__ keys,
                                   uint32_t* __restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t c = records[i * DDN_RECORD_WORDS];
  const uint64_t kx = c & 0x1fffff, ky = (c >> 21) & 0x1fffff, kz = (c >> 42) & 0x1fffff;
  keys[i] = (KeyT)kx | ((KeyT)ky << g.bx) | ((KeyT)kz << (g.bx + g.by));
  idx[i] = (uint32_t)i;
}

template <typename KeyT>
__global__ void merge_finalize_counts_kernel(int64_t n, const int* __restrict__ num_runs, int64_t* __restrict__ counts_out) {
  counts_out[0] = n;
  counts_out[1] = *num_runs;
}

template <typename KeyT>
__global__ void __launch_bounds__(256)
merge_segments_kernel(GridDev g, const KeyT* __restrict__ unique_keys, const int* __restrict__ run_counts,
                      const int* __restrict__ run_starts, const int64_t* __restrict__ counts,
                      const uint32_t* __restrict__ sorted_idx, const unsigned long long* __restrict__ records,
                      uint64_t* __restrict__ out_keys, float* __restrict__ out_xyz, uint8_t* __restrict__ out_rgb,
                      int32_t* __restrict__ out_count) {
  const int64_t mv = counts[1];
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < mv; r += (int64_t)gridDim.x * blockDim.x) {
    const KeyT key = unique_keys[r];
    const uint32_t kx = (uint32_t)(key & (((KeyT)1 << g.bx) - 1));
    const uint32_t ky = (uint32_t)((key >> g.bx) & (((KeyT)1 << g.by) - 1));
    const uint32_t kz = (uint32_t)((key >> (g.bx + g.by)) & (((KeyT)1 << g.bz) - 1));
    const float cx = voxel_centre(g.ox, kx, g.voxel), cy = voxel_centre(g.oy, ky, g.voxel), cz = voxel_centre(g.oz, kz, g.voxel);
    const int start = run_starts[r], nrec = run_counts[r];
    long long sx = 0, sy = 0, sz = 0, cnt = 0;
    unsigned long long sr = 0, sg = 0, sb = 0;
    for (int j = 0; j < nrec; ++j) {
      const size_t i = sorted_idx[start + j];
      const unsigned long long* q = records + i * DDN_RECORD_WORDS;
      sx += (long long)q[1];
      sy += (long long)q[2];
      sz += (long long)q[3];
      sr += q[4] >> 32;
      sg += q[4] & 0xffffffffull;
      sb += q[5] >> 32;
      cnt += (long long)(q[5] & 0xffffffffull);
    }
    out_keys[r] = (uint64_t)kx | ((uint64_t)ky << 21) | ((uint64_t)kz << 42);
    out_count[r] = (int32_t)cnt;
    finalize_voxel(g, cx, cy, cz, sx, sy, sz, sr, sg, sb, cnt, out_xyz + r * 3, out_rgb + r * 3);
  }
}

struct FuseLayout {
  size_t keys_a, keys_b, idx_a, idx_b, uniq, run_counts, run_starts, num_runs, cub_temp, total;
  size_t cub_temp_bytes;
};

template <typename KeyT>
static int fuse_layout(int64_t n, FuseLayout* L) {
  size_t t_sort = 0, t_rle = 0, t_scan = 0;
  cub::DoubleBuffer<KeyT> dk(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> dv(nullptr, nullptr);
  DDN_TRY(check_cuda(cub::DeviceRadixSort::SortPairs(nullptr, t_sort, dk, dv, (int)n, 0, (int)sizeof(KeyT) * 8),
                     "cub sort size"));
  DDN_TRY(check_cuda(cub::DeviceRunLengthEncode::Encode(nullptr, t_rle, (KeyT*)nullptr, (KeyT*)nullptr, (int*)nullptr,
                                                        (int*)nullptr, (int)n),
                     "cub rle size"));
  DDN_TRY(check_cuda(cub::DeviceScan::ExclusiveSum(nullptr, t_scan, (int*)nullptr, (int*)nullptr, (int)n), "cub scan size"));
  size_t temp = t_sort > t_rle ? t_sort : t_rle;
  temp = temp > t_scan ? temp : t_scan;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (size_t)align_up((int64_t)bytes, 256);
    return o;
  };
  L->keys_a = take(n * sizeof(KeyT));
  L->keys_b = take(n * sizeof(KeyT));
  L->idx_a = take(n * 4);
  L->idx_b = take(n * 4);
  L->uniq = take(n * sizeof(KeyT));
  L->run_counts = take(n * 4);
  L->run_starts = take((n + 1) * 4);
  L->num_runs = take(16);
  L->cub_temp = take(temp);
  L->cub_temp_bytes = temp;
  L->total = off + 256;
  return DDN_OK;
}

template <typename KeyT>
static int fuse_impl(const GridDev& g, int64_t n, const float* xyz, const uint8_t* rgb, const uint8_t* votes, int thr,
                     uint64_t* out_keys, float* out_xyz, uint8_t* out_rgb, int32_t* out_count, int64_t* counts_out,
                     void* workspace, int64_t workspace_bytes, cudaStream_t st, unsigned long long* records = nullptr) {
  FuseLayout L;
  DDN_TRY(fuse_layout<KeyT>(n, &L));
  if ((int64_t)L.total > workspace_bytes) {
    set_error("fuse workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)L.total);
    return DDN_ERR_WORKSPACE_TOO_SMALL;
  }
  char* base = reinterpret_cast<char*>(align_up((int64_t)(uintptr_t)workspace, 256));
  KeyT* keys_a = (KeyT*)(base + L.keys_a);
  KeyT* keys_b = (KeyT*)(base + L.keys_b);
  uint32_t* idx_a = (uint32_t*)(base + L.idx_a);
  uint32_t* idx_b = (uint32_t*)(base + L.idx_b);
  KeyT* uniq = (KeyT*)(base + L.uniq);
  int* run_counts = (int*)(base + L.run_counts);
  int* run_starts = (int*)(base + L.run_starts);
  int* num_runs = (int*)(base + L.num_runs);
  void* temp = base + L.cub_temp;
  size_t temp_bytes = L.cub_temp_bytes;

  const unsigned blocks = (unsigned)((n + kKeyThreads - 1) / kKeyThreads);
  voxel_key_kernel<KeyT><<<blocks, kKeyThreads, 0, st>>>(g, n, xyz, votes, thr, keys_a, idx_a);
  DDN_TRY(after_launch("voxel_key_kernel"));

  cub::DoubleBuffer<KeyT> dk(keys_a, keys_b);
  cub::DoubleBuffer<uint32_t> dv(idx_a, idx_b);
  const int end_bit = g.bx + g.by + g.bz + 1;  // + the sentinel bit
  DDN_TRY(check_cuda(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, dk, dv, (int)n, 0, end_bit, st), "cub sort"));
  g_launches.fetch_add((end_bit + 7) / 8 + 1, std::memory_order_relaxed);
  temp_bytes = L.cub_temp_bytes;
  DDN_TRY(check_cuda(cub::DeviceRunLengthEncode::Encode(temp, temp_bytes, dk.Current(), uniq, run_counts, num_runs, (int)n, st),
                     "cub rle"));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  temp_bytes = L.cub_temp_bytes;
  DDN_TRY(check_cuda(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, run_counts, run_starts, (int)n, st), "cub scan"));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  fuse_finalize_kernel<KeyT><<<1, 1, 0, st>>>(g, n, uniq, run_counts, num_runs, counts_out);
  DDN_TRY(after_launch("fuse_finalize_kernel"));
  if (records != nullptr)
    segment_mean_kernel<KeyT, true><<<kNumSMs * 8, 256, 0, st>>>(g, uniq, run_counts, run_starts, counts_out, dv.Current(),
                                                                xyz, rgb, out_keys, out_xyz, out_rgb, out_count, records);
  else
    segment_mean_kernel<KeyT, false><<<kNumSMs * 8, 256, 0, st>>>(g, uniq, run_counts, run_starts, counts_out, dv.Current(),
                                                                 xyz, rgb, out_keys, out_xyz, out_rgb, out_count, nullptr);
  return after_launch("segment_mean_kernel");
}

template <typename KeyT>
static int merge_impl(const GridDev& g, int64_t n, const unsigned long long* records, uint64_t* out_keys, float* out_xyz, uint8_t* out_rgb, int32_t* out_count,
                      int64_t* counts_out, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  FuseLayout L;
  DDN_TRY(fuse_layout<KeyT>(n, &L));
  if ((int64_t)L.total > workspace_bytes) {
    set_error("merge workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)L.total);
    return DDN_ERR_WORKSPACE_TOO_SMALL;
  }
  char* base = reinterpret_cast<char*>(align_up((int64_t)(uintptr_t)workspace, 256));
  KeyT* keys_a = (KeyT*)(base + L.keys_a);
  KeyT* keys_b = (KeyT*)(base + L.keys_b);
  uint32_t* idx_a = (uint32_t*)(base + L.idx_a);
  uint32_t* idx_b = (uint32_t*)(base + L.idx_b);
  KeyT* uniq = (KeyT*)(base + L.uniq);
  int* run_counts = (int*)(base + L.run_counts);
  int* run_starts = (int*)(base + L.run_starts);
  int* num_runs = (int*)(base + L.num_runs);
  void* temp = base + L.cub_temp;
  size_t temp_bytes = L.cub_temp_bytes;
  compact_key_kernel<KeyT><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, n, records, keys_a, idx_a);
  DDN_TRY(after_launch("compact_key_kernel"));
  cub::DoubleBuffer<KeyT> dk(keys_a, keys_b);
  cub::DoubleBuffer<uint32_t> dv(idx_a, idx_b);
  const int end_bit = g.bx + g.by + g.bz;
  DDN_TRY(check_cuda(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, dk, dv, (int)n, 0, end_bit, st), "cub sort"));
  g_launches.fetch_add((end_bit + 7) / 8 + 1, std::memory_order_relaxed);
  temp_bytes = L.cub_temp_bytes;
  DDN_TRY(check_cuda(cub::DeviceRunLengthEncode::Encode(temp, temp_bytes, dk.Current(), uniq, run_counts, num_runs, (int)n, st),
                     "cub rle"));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  temp_bytes = L.cub_temp_bytes;
  DDN_TRY(check_cuda(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, run_counts, run_starts, (int)n, st), "cub scan"));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  merge_finalize_counts_kernel<KeyT><<<1, 1, 0, st>>>(n, num_runs, counts_out);
  DDN_TRY(after_launch("merge_finalize_counts_kernel"));
  merge_segments_kernel<KeyT><<<kNumSMs * 8, 256, 0, st>>>(g, uniq, run_counts, run_starts, counts_out, dv.Current(), records,
                                                          out_keys, out_xyz, out_rgb, out_count);
  return after_launch("merge_segments_kernel");
}

int sort_fuse_workspace_bytes(int64_t n_points, int64_t* bytes_out) {
  FuseLayout L;
  DDN_TRY(fuse_layout<uint64_t>(n_points > 0 ? n_points : 1, &L));  // worst case (64-bit keys)
  *bytes_out = (int64_t)L.total;
  return DDN_OK;
}

int sort_fuse_points(const GridDev& g, int64_t n, const float* xyz, const uint8_t* rgb, const uint8_t* votes, int thr,
                     uint64_t* out_keys, float* out_xyz, uint8_t* out_rgb, int32_t* out_count, int64_t* counts_out,
                     void* workspace, int64_t workspace_bytes, cudaStream_t st, unsigned long long* records) {
  if (g.bx + g.by + g.bz <= 31)
    return fuse_impl<uint32_t>(g, n, xyz, rgb, votes, thr, out_keys, out_xyz, out_rgb, out_count, counts_out, workspace,
                               workspace_bytes, st, records);
  return fuse_impl<uint64_t>(g, n, xyz, rgb, votes, thr, out_keys, out_xyz, out_rgb, out_count, counts_out, workspace,
                             workspace_bytes, st, records);
}

int sort_merge_records(const GridDev& g, int64_t n, const unsigned long long* records, uint64_t* out_keys, float* out_xyz,
                       uint8_t* out_rgb, int32_t* out_count, int64_t* counts_out, void* workspace, int64_t workspace_bytes,
                       cudaStream_t st) {
  if (g.bx + g.by + g.bz <= 32)
    return merge_impl<uint32_t>(g, n, records, out_keys, out_xyz, out_rgb, out_count, counts_out, workspace, workspace_bytes, st);
  return merge_impl<uint64_t>(g, n, records, out_keys, out_xyz, out_rgb, out_count, counts_out, workspace, workspace_bytes, st);
}

}  // namespace ddn
