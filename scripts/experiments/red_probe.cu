// Probe: throughput of scattered global reductions, 32-bit vs 64-bit, 1..5 consecutive words per "voxel".
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a red_probe.cu -o red_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <typename T, int kWords>
__global__ void red_kernel(T* acc, const uint32_t* slots, int64_t n, int stride) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T* a = acc + (size_t)slots[i] * stride;
#pragma unroll
  for (int w = 0; w < kWords; ++w) atomicAdd(a + w, (T)(i + w));
}
// vector float reductions (sm_90+): one op adds 4 (or 2) floats
template <int kOps, int kVec>
__global__ void redv_kernel(float* acc, const uint32_t* slots, int64_t n, int stride) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* a = acc + (size_t)slots[i] * stride;
  const float v = (float)(i & 1023);
#pragma unroll
  for (int w = 0; w < kOps; ++w) {
    if (kVec == 4) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a + 4 * w), "f"(v), "f"(v + 1.f), "f"(v + 2.f), "f"(1.f) : "memory");
    else asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(a + 2 * w), "f"(v), "f"(1.f) : "memory");
  }
}
template <int kOps, int kVec>
void runv(const char* name, float* acc, uint32_t* slots, int64_t n, int stride) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int it = 0; it < 2; ++it) redv_kernel<kOps, kVec><<<(unsigned)((n + 255) / 256), 256>>>(acc, slots, n, stride);
  cudaEventRecord(a);
  for (int it = 0; it < 5; ++it) redv_kernel<kOps, kVec><<<(unsigned)((n + 255) / 256), 256>>>(acc, slots, n, stride);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  printf("%-28s %6.3f ms  %7.1f G ops/s (%s)\n", name, ms, n * (double)kOps / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
__global__ void fill(uint32_t* s, int64_t n, uint32_t m, uint32_t window) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // a moving window of `window` voxels: like consecutive pixels of a depth map falling into neighbouring voxels
  if (i < n) { uint32_t h = (uint32_t)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; s[i] = (h % window + (uint32_t)(i / 6)) % m; }
}
template <typename T, int kWords>
void run(const char* name, T* acc, uint32_t* slots, int64_t n, int stride) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int it = 0; it < 2; ++it) red_kernel<T, kWords><<<(unsigned)((n + 255) / 256), 256>>>(acc, slots, n, stride);
  cudaEventRecord(a);
  for (int it = 0; it < 5; ++it) red_kernel<T, kWords><<<(unsigned)((n + 255) / 256), 256>>>(acc, slots, n, stride);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  printf("%-28s %6.3f ms  %7.1f G lane-REDs/s\n", name, ms, n * (double)kWords / ms / 1e6);
}
int main() {
  const int64_t n = 86000000; const uint32_t m = 14700000;
  uint32_t* slots; cudaMalloc(&slots, n * 4);
  void* acc; cudaMalloc(&acc, (size_t)m * 64); cudaMemset(acc, 0, (size_t)m * 64);
  for (uint32_t window : {64u, 4096u, 14700000u}) {
  printf("-- window %u voxels\n", window);
  fill<<<(unsigned)((n + 255) / 256), 256>>>(slots, n, m, window);
  run<unsigned long long, 5>("5 x u64 (40 B, stride 5)", (unsigned long long*)acc, slots, n, 5);
  run<unsigned long long, 4>("4 x u64 (32 B, stride 4)", (unsigned long long*)acc, slots, n, 4);
  run<unsigned long long, 3>("3 x u64 (24 B, stride 3)", (unsigned long long*)acc, slots, n, 3);
  run<unsigned long long, 1>("1 x u64", (unsigned long long*)acc, slots, n, 1);
  run<unsigned int, 7>("7 x u32 (28 B, stride 7)", (unsigned int*)acc, slots, n, 7);
  run<unsigned int, 8>("8 x u32 (32 B, stride 8)", (unsigned int*)acc, slots, n, 8);
  run<unsigned int, 5>("5 x u32 (20 B, stride 5)", (unsigned int*)acc, slots, n, 5);
  run<unsigned int, 1>("1 x u32", (unsigned int*)acc, slots, n, 1);
  runv<2, 4>("2 x v4.f32 (32 B, stride 8)", (float*)acc, slots, n, 8);
  runv<1, 4>("1 x v4.f32 (16 B, stride 4)", (float*)acc, slots, n, 4);
  runv<4, 2>("4 x v2.f32 (32 B, stride 8)", (float*)acc, slots, n, 8);
  }
  return 0;
}
