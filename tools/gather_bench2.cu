// gather_bench2.cu -- round-2 follow-up to gather_bench.cu (evidence tool, not product): can anything beat the LSU's
// one-line-per-lane-and-cycle gather rate (290 G/s on this chip) for an L2-resident 32 MiB table slice?
//   (a) plain 8-byte __ldg gathers (the round-1 ceiling, for reference)
//   (b) the same gathers through the TEXTURE path (tex1Dfetch<int2> on a linear texture object)
//   (c) 8-byte gathers with L1::no_allocate
//   (d) probe-like shapes: streamed keys in, gather, 16 B out -- LDG vs TEX
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_bench2 tools/gather_bench2.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__host__ __device__ inline uint64_t mm(uint64_t x) { x ^= x >> 32; x *= 0xd6e8feb86659fd93ULL; x ^= x >> 32; x *= 0xd6e8feb86659fd93ULL; x ^= x >> 32; return x; }

template <int KPT, int PATH>  // PATH 0: __ldg, 1: tex1Dfetch, 2: ld.global.nc.L1::no_allocate, 3: alternating __ldg / tex1Dfetch
__global__ void gather_kernel(const uint64_t *__restrict__ table, cudaTextureObject_t tex, uint32_t mask, size_t n, uint64_t *sink) {
  size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t) gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (size_t i = tid; i < n; i += stride * KPT) {
    uint64_t v[KPT];
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
      uint32_t s = (uint32_t) mm(i + j * stride + 12345) & mask;
      if (PATH == 0) v[j] = __ldg((const unsigned long long *) table + s);
      else if (PATH == 1) { int2 t = tex1Dfetch<int2>(tex, (int) s); v[j] = ((uint64_t) (uint32_t) t.y << 32) | (uint32_t) t.x; }
      else if (PATH == 2) asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v[j]) : "l"(table + s));
      else if (j & 1) { int2 t = tex1Dfetch<int2>(tex, (int) s); v[j] = ((uint64_t) (uint32_t) t.y << 32) | (uint32_t) t.x; }
      else v[j] = __ldg((const unsigned long long *) table + s);
    }
#pragma unroll
    for (int j = 0; j < KPT; ++j) acc += v[j];
  }
  if (acc == 0x1234567) sink[0] = acc;
}

template <int KPT, int PATH>
__global__ void probe_like_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ table, cudaTextureObject_t tex, uint32_t mask, size_t n,
                                  uint64_t *out_a, uint64_t *out_b) {
  size_t tile = (size_t) blockDim.x * KPT;
  for (size_t base = (size_t) blockIdx.x * tile; base < n; base += (size_t) gridDim.x * tile) {
    uint64_t k[KPT], v[KPT];
#pragma unroll
    for (int j = 0; j < KPT; ++j) { size_t i = base + j * blockDim.x + threadIdx.x; k[j] = i < n ? __ldg((const unsigned long long *) keys + i) : 0; }
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
      uint32_t s = (uint32_t) mm(k[j]) & mask;
      if (PATH == 0) v[j] = __ldg((const unsigned long long *) table + s);
      else { int2 t = tex1Dfetch<int2>(tex, (int) s); v[j] = ((uint64_t) (uint32_t) t.y << 32) | (uint32_t) t.x; }
    }
#pragma unroll
    for (int j = 0; j < KPT; ++j) { size_t i = base + j * blockDim.x + threadIdx.x; if (i < n) { out_a[i] = k[j]; out_b[i] = v[j]; } }
  }
}

__global__ void fill_kernel(uint64_t *p, size_t n) {
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = mm(i);
}

template <class F>
static float time_ms(F f, int reps = 3) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms; }
  return best;
}

int main() {
  size_t n = (size_t) 1 << 29;
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("# %s, %d SMs, L2 %d MB; gathers over an L2-resident table slice, G accesses/s\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20);
  uint64_t *table, *sink, *keys, *oa, *ob;
  const int max_log2 = 23;  // 64 MiB
  CK(cudaMalloc(&table, ((size_t) 1 << max_log2) * 8)); CK(cudaMalloc(&sink, 64));
  fill_kernel<<<148 * 8, 256>>>(table, (size_t) 1 << max_log2);
  CK(cudaMalloc(&keys, n * 8)); CK(cudaMalloc(&oa, n * 8)); CK(cudaMalloc(&ob, n * 8));
  fill_kernel<<<148 * 8, 256>>>(keys, n);
  CK(cudaDeviceSynchronize());
  printf("footprint_MiB, ldg_kpt4, tex_kpt4, tex_kpt8, ldg_noalloc_kpt4, probe_like_ldg_kpt4, probe_like_tex_kpt4, mixed_ldg_tex_kpt8\n");
  for (int lg : {20, 22, 23}) {
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = table; rd.res.linear.desc = cudaCreateChannelDesc<int2>();
    rd.res.linear.sizeInBytes = ((size_t) 1 << lg) * 8;
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    uint32_t mask = (1u << lg) - 1;
    int grid = 148 * 8;
    float a = time_ms([&] { gather_kernel<4, 0><<<grid, 256>>>(table, tex, mask, n, sink); });
    float b = time_ms([&] { gather_kernel<4, 1><<<grid, 256>>>(table, tex, mask, n, sink); });
    float c = time_ms([&] { gather_kernel<8, 1><<<grid, 256>>>(table, tex, mask, n, sink); });
    float d = time_ms([&] { gather_kernel<4, 2><<<grid, 256>>>(table, tex, mask, n, sink); });
    float g = time_ms([&] { gather_kernel<8, 3><<<grid, 256>>>(table, tex, mask, n, sink); });
    float e = time_ms([&] { probe_like_kernel<4, 0><<<grid, 256>>>(keys, table, tex, mask, n, oa, ob); });
    float f = time_ms([&] { probe_like_kernel<4, 1><<<grid, 256>>>(keys, table, tex, mask, n, oa, ob); });
    printf("%zu, %.1f, %.1f, %.1f, %.1f, %.1f, %.1f, %.1f\n", (((size_t) 1 << lg) * 8) >> 20, n / a / 1e6, n / b / 1e6, n / c / 1e6, n / d / 1e6, n / e / 1e6, n / f / 1e6, n / g / 1e6);
    fflush(stdout);
    CK(cudaDestroyTextureObject(tex));
  }
  // mixed: half of a thread's gathers through LDG, half through TEX (do the two paths add up?)
  return 0;
}
