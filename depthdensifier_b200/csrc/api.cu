// Library-level entry points of libddn_b200.so: version, error string, launch counter, defaults.
#include <stdarg.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"

namespace ddn {

std::atomic<int> g_profile{0};
static std::vector<std::pair<std::string, cudaEvent_t>> g_marks;

void profile_mark(const char* name, cudaStream_t st) {
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, st);
  g_marks.emplace_back(name, ev);
}

static thread_local char t_error[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

// Diagnostic: read `n16` 16-byte words from `src` (local or NVLink peer memory) with one of three load flavours
// and fold them into out[0] - measures what SM-issued loads get out of a link (scripts/experiments/peer_read_probe.py).
template <int kMode>
__global__ void __launch_bounds__(256) peer_read_kernel(const uint4* __restrict__ src, long long n16, int per_thread, unsigned* out) {
  unsigned acc = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n16; i0 += stride * per_thread) {
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long i = i0 + k * stride;
      v[k] = make_uint4(0, 0, 0, 0);
      if (k < per_thread && i < n16) {
        if (kMode == 0) v[k] = __ldcv(src + i);
        else if (kMode == 1) v[k] = src[i];
        else v[k] = __ldg(src + i);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
  }
  if (acc == 0x12345679u) out[0] = acc;
}

}  // namespace ddn

extern "C" {

int ddn_debug_peer_read(const void* src, int64_t bytes, int32_t mode, int32_t per_thread, int32_t ctas, void* out, void* stream) {
  using namespace ddn;
  DDN_REQUIRE(src && out && bytes >= 16 && mode >= 0 && mode <= 2 && per_thread >= 1 && per_thread <= 8 && ctas >= 1, "arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n16 = bytes / 16;
  if (mode == 0) peer_read_kernel<0><<<ctas, 256, 0, st>>>((const uint4*)src, n16, per_thread, (unsigned*)out);
  else if (mode == 1) peer_read_kernel<1><<<ctas, 256, 0, st>>>((const uint4*)src, n16, per_thread, (unsigned*)out);
  else peer_read_kernel<2><<<ctas, 256, 0, st>>>((const uint4*)src, n16, per_thread, (unsigned*)out);
  return after_launch("peer_read_kernel");
}

int ddn_version(void) { return DDN_VERSION; }

const char* ddn_last_error_string(void) { return ddn::t_error; }

int64_t ddn_launch_count(void) { return ddn::g_launches.load(std::memory_order_relaxed); }

void ddn_profile_enable(int on) {
  ddn::g_profile.store(on ? 1 : 0);
  if (on) ddn::profile_mark("(start)", nullptr);
}

int ddn_profile_report(char* buf, int64_t size) {
  using namespace ddn;
  if (buf == nullptr || size <= 0) return DDN_ERR_INVALID_ARGUMENT;
  cudaDeviceSynchronize();
  std::string out;
  for (size_t i = 1; i < g_marks.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_marks[i - 1].second, g_marks[i].second);
    char line[160];
    snprintf(line, sizeof(line), "%s %.4f\n", g_marks[i].first.c_str(), ms);
    out += line;
  }
  for (auto& m : g_marks) cudaEventDestroy(m.second);
  g_marks.clear();
  snprintf(buf, (size_t)size, "%s", out.c_str());
  return DDN_OK;
}

void ddn_align_config_default(ddn_align_config* cfg) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->min_correspondences = 50;
  cfg->edge_margin = 10;
  cfg->robust = 1;
  cfg->outlier_threshold = 2.5f;
  cfg->skip_smoothing = 0;
  cfg->adaptive_correspondences = 1;
  cfg->max_pairs = 500;
  cfg->mode = 0;
  cfg->subsample_seed = 0;
  cfg->zero_unmasked_passthrough = 0;
  cfg->mask_packed = 0;
  cfg->use_tma = 0;
}

void ddn_filter_config_default(ddn_filter_config* cfg) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->depth_threshold = 0.7f;
  cfg->grazing_cos = 0.087f;
  cfg->sample_mode = 0;
  cfg->two_sided_tau = 0.0f;
  cfg->stride = 1;
  cfg->normals_in_world = 0;
  cfg->pixel_layout = 0;
}

}  // extern "C"
