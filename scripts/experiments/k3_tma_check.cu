// Stage-1 alignment (ddn_align_views) with the remap kernel's depth tile loaded by TMA (use_tma = 1) against the default
// LDG form, through the C ABI and WITHOUT Python (a fresh GPU box spends up to a minute importing torch; this program
// starts in about two seconds): bit identity of the refined maps and view stats on small scenes with partial tiles,
// then per-kernel times at 24 views of 1920 x 1080 from the library's own profiling marks.
//   k3_tma_check <path to libddn_b200.so> [out file]
// The scene is synthetic: identity poses, a smooth positive depth map with hash noise, a random mask with 10 % holes and
// 2048 sparse points per view placed on z = 1.7 * depth + 0.3, so that every view comes out DDN_VIEW_REFINED.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/ddn_b200.h"

static FILE* g_out = nullptr;
static void say(const char* fmt, ...) {
  char buf[4096];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  fputs(buf, stdout);
  fflush(stdout);
  if (g_out) { fputs(buf, g_out); fflush(g_out); }
}

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) { say("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(3); } \
  } while (0)

__host__ __device__ inline unsigned mix(unsigned h) {
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
__host__ __device__ inline float scene_depth(int v, float x, float y) { return 3.f + 1.5f * sinf(0.011f * x + (float)v) * cosf(0.013f * y); }

__global__ void fill_scene(float* depth, uint8_t* mask, int V, int H, int W) {
  const size_t n = (size_t)V * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), v = (int)(i / ((size_t)W * H));
    const unsigned h = mix((unsigned)i * 2654435761u + 17u);
    depth[i] = scene_depth(v, (float)x, (float)y) + 0.05f * ((float)(h & 0xffff) / 65536.f - 0.5f);
    mask[i] = (mix(h) % 10u) != 0u;
  }
}

typedef int (*fn_ws)(int64_t, int64_t, int64_t*);
typedef void (*fn_def)(ddn_align_config*);
typedef int (*fn_align)(const ddn_align_config*, int64_t, int64_t, int64_t, const float*, const uint8_t*, const double*, const double*,
                        const double*, const int64_t*, int64_t, float*, ddn_view_stats*, void*, int64_t, const float*, float*, void*);
typedef const char* (*fn_err)(void);
typedef void (*fn_prof)(int);
typedef int (*fn_rep)(char*, int64_t);

struct Lib {
  fn_ws ws; fn_def def; fn_align align; fn_err err; fn_prof prof; fn_rep rep;
} L;

struct Scene {
  int V, H, W, C;
  float *depth, *out0, *out1;
  uint8_t* mask;
  double *poses, *kmat, *sparse;
  int64_t* offs;
  ddn_view_stats *st0, *st1;
  void* ws;
  int64_t ws_bytes;
};

static Scene make_scene(int V, int H, int W, int C) {
  Scene s{};
  s.V = V, s.H = H, s.W = W, s.C = C;
  const size_t n = (size_t)V * H * W;
  CK(cudaMalloc(&s.depth, n * 4));
  CK(cudaMalloc(&s.out0, n * 4));
  CK(cudaMalloc(&s.out1, n * 4));
  CK(cudaMalloc(&s.mask, n));
  fill_scene<<<148 * 8, 256>>>(s.depth, s.mask, V, H, W);
  CK(cudaGetLastError());
  std::vector<double> poses((size_t)V * 12, 0.0), kmat((size_t)V * 9, 0.0), sp((size_t)V * C * 3);
  std::vector<int64_t> offs(V + 1);
  const double f = (double)W, cx = W * 0.5, cy = H * 0.5;
  for (int v = 0; v < V; ++v) {
    poses[v * 12 + 0] = poses[v * 12 + 5] = poses[v * 12 + 10] = 1.0;
    kmat[v * 9 + 0] = kmat[v * 9 + 4] = f;
    kmat[v * 9 + 2] = cx, kmat[v * 9 + 5] = cy, kmat[v * 9 + 8] = 1.0;
    offs[v] = (int64_t)v * C;
    for (int i = 0; i < C; ++i) {
      const unsigned h = mix((unsigned)(v * C + i) * 747796405u + 3u), h2 = mix(h);
      const double u = 12.0 + (h % (unsigned)(W - 24)) + 0.37, w = 12.0 + (h2 % (unsigned)(H - 24)) + 0.61;
      const double z = 1.7 * scene_depth(v, (float)u, (float)w) + 0.3 + 0.01 * ((mix(h2) & 0xff) / 255.0 - 0.5);
      double* p = &sp[((size_t)v * C + i) * 3];
      p[0] = (u - cx) * z / f, p[1] = (w - cy) * z / f, p[2] = z;
    }
  }
  offs[V] = (int64_t)V * C;
  CK(cudaMalloc(&s.poses, poses.size() * 8));
  CK(cudaMalloc(&s.kmat, kmat.size() * 8));
  CK(cudaMalloc(&s.sparse, sp.size() * 8));
  CK(cudaMalloc(&s.offs, offs.size() * 8));
  CK(cudaMemcpy(s.poses, poses.data(), poses.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(s.kmat, kmat.data(), kmat.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(s.sparse, sp.data(), sp.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(s.offs, offs.data(), offs.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&s.st0, sizeof(ddn_view_stats) * V));
  CK(cudaMalloc(&s.st1, sizeof(ddn_view_stats) * V));
  if (L.ws(V, C, &s.ws_bytes) != DDN_OK) { say("workspace_bytes: %s\n", L.err()); exit(2); }
  CK(cudaMalloc(&s.ws, (size_t)s.ws_bytes));
  return s;
}

static void free_scene(Scene& s) {
  cudaFree(s.depth), cudaFree(s.out0), cudaFree(s.out1), cudaFree(s.mask), cudaFree(s.poses), cudaFree(s.kmat), cudaFree(s.sparse);
  cudaFree(s.offs), cudaFree(s.st0), cudaFree(s.st1), cudaFree(s.ws);
}

static int run_align(const Scene& s, int use_tma, bool with_mask) {
  ddn_align_config cfg;
  L.def(&cfg);
  cfg.zero_unmasked_passthrough = 1;
  cfg.use_tma = use_tma;
  return L.align(&cfg, s.V, s.H, s.W, s.depth, with_mask ? s.mask : nullptr, s.poses, s.kmat, s.sparse, s.offs, s.C, use_tma ? s.out1 : s.out0,
                 use_tma ? s.st1 : s.st0, s.ws, s.ws_bytes, nullptr, nullptr, nullptr);
}

// median of the per-call times of `kernel` in a profile report ("name ms" lines)
static double median_of(const std::string& rep, const char* kernel) {
  std::vector<double> t;
  size_t pos = 0;
  while (pos < rep.size()) {
    const size_t e = rep.find('\n', pos);
    const std::string line = rep.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
    if (line.rfind(kernel, 0) == 0) t.push_back(atof(line.c_str() + strlen(kernel)));
    if (e == std::string::npos) break;
    pos = e + 1;
  }
  if (t.size() > 2) t.erase(t.begin());  // first call: cold
  std::sort(t.begin(), t.end());
  return t.empty() ? -1.0 : t[t.size() / 2];
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: k3_tma_check <libddn_b200.so> [out]\n"); return 1; }
  if (argc > 2) g_out = fopen(argv[2], "w");
  void* lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { say("dlopen: %s\n", dlerror()); return 2; }
  L.ws = (fn_ws)dlsym(lib, "ddn_align_workspace_bytes");
  L.def = (fn_def)dlsym(lib, "ddn_align_config_default");
  L.align = (fn_align)dlsym(lib, "ddn_align_views");
  L.err = (fn_err)dlsym(lib, "ddn_last_error_string");
  L.prof = (fn_prof)dlsym(lib, "ddn_profile_enable");
  L.rep = (fn_rep)dlsym(lib, "ddn_profile_report");
  if (!L.ws || !L.def || !L.align || !L.err || !L.prof || !L.rep) { say("missing symbol\n"); return 2; }
  CK(cudaSetDevice(0));

  // timing first (the number asked for), identity checks after it
  const int shapes[][4] = {{24, 1080, 1920, 2048}, {5, 120, 160, 600}, {4, 152, 200, 600}, {3, 384, 512, 600}, {2, 34, 132, 600}};
  int all_ok = 1;
  for (int si = 0; si < 5; ++si) {
    const int V = shapes[si][0], H = shapes[si][1], W = shapes[si][2], C = shapes[si][3];
    Scene s = make_scene(V, H, W, C);
    const size_t n = (size_t)V * H * W;
    std::vector<float> h0(n), h1(n);
    std::vector<ddn_view_stats> s0(V), s1(V);
    for (int with_mask = 1; with_mask >= (si == 0 ? 1 : 0); --with_mask) {
      CK(cudaMemset(s.out0, 0xff, n * 4));
      CK(cudaMemset(s.out1, 0xee, n * 4));
      int rc = run_align(s, 0, with_mask);
      if (rc != DDN_OK) { say("ldg align failed rc=%d: %s\n", rc, L.err()); return 2; }
      CK(cudaDeviceSynchronize());
      rc = run_align(s, 1, with_mask);
      if (rc != DDN_OK) { say("tma align failed rc=%d: %s\n", rc, L.err()); return 2; }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { say("%dx%dx%d mask=%d: TMA variant faulted: %s\n", V, W, H, with_mask, cudaGetErrorString(e)); return 3; }
      CK(cudaMemcpy(h0.data(), s.out0, n * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(h1.data(), s.out1, n * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(s0.data(), s.st0, sizeof(ddn_view_stats) * V, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(s1.data(), s.st1, sizeof(ddn_view_stats) * V, cudaMemcpyDeviceToHost));
      size_t diff = 0, first = n, nonzero = 0;
      for (size_t i = 0; i < n; ++i) {
        if (memcmp(&h0[i], &h1[i], 4) != 0) { if (first == n) first = i; ++diff; }
        nonzero += h0[i] > 0.f;
      }
      int refined = 0;
      for (int v = 0; v < V; ++v) refined += s0[v].status == DDN_VIEW_REFINED;
      const int same_stats = memcmp(s0.data(), s1.data(), sizeof(ddn_view_stats) * V) == 0;
      say("identity %dx%dx%d mask=%d: %zu of %zu pixels differ, stats %s, %d/%d views refined, %.1f %% of pixels > 0", V, W, H, with_mask, diff, n,
          same_stats ? "equal" : "DIFFER", refined, V, 100.0 * nonzero / n);
      if (diff) say(" (first at view %zu y %zu x %zu: ldg %g tma %g)", first / ((size_t)H * W), (first / W) % H, first % W, h0[first], h1[first]);
      say("\n");
      all_ok &= diff == 0 && same_stats && refined == V;
    }
    if (si == 0) {  // per-kernel times, 9 calls each, interleaved
      for (int use_tma = 0; use_tma < 2; ++use_tma) {
        L.prof(1);
        for (int r = 0; r < 9; ++r)
          if (run_align(s, use_tma, true) != DDN_OK) { say("align failed: %s\n", L.err()); return 2; }
        std::vector<char> buf(1 << 16);
        L.rep(buf.data(), (int64_t)buf.size());
        L.prof(0);
        const std::string rep(buf.data());
        const double k3 = median_of(rep, "remap_median_kernel"), k12 = median_of(rep, "align_stats_kernel");
        const double bytes = (double)n * 9.0;  // depth 4 + mask 1 + refined 4 bytes per pixel
        say("timing 24x1920x1080 %s: remap_median_kernel %.4f ms (%.0f GB/s algorithmic, 9 B/pixel), align_stats_kernel %.4f ms\n",
            use_tma ? "tma" : "ldg", k3, bytes / (k3 * 1e6), k12);
      }
    }
    free_scene(s);
  }
  say(all_ok ? "ALL IDENTICAL\n" : "MISMATCH\n");
  return all_ok ? 0 : 4;
}
