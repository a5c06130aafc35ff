// gather_bench.cu -- hardware ceilings for the probe path on this GPU (evidence tool, not product):
//   * random 8-byte / 32-byte gathers per second as a function of table footprint
//     (L2-resident .. 8 GiB: TLB + HBM sector behaviour), with and without a streamed key input/output
//   * pinned H2D / D2H copy bandwidth (the e2e bound)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_bench tools/gather_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__host__ __device__ inline uint64_t mm(uint64_t x) { x ^= x >> 32; x *= 0xd6e8feb86659fd93ULL; x ^= x >> 32; x *= 0xd6e8feb86659fd93ULL; x ^= x >> 32; return x; }

// each thread issues KPT independent random loads per iteration (addresses from a counter hash)
template <int KPT, int WIDTH>  // WIDTH: 8 or 32 bytes per access
__global__ void gather_kernel(const uint64_t *__restrict__ table, uint64_t mask, size_t n, uint64_t *sink) {
  size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t) gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (size_t i = tid; i < n; i += stride * KPT) {
    uint64_t v[KPT][WIDTH / 8];
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
      uint64_t s = mm(i + j * stride + 12345) & mask;
      if (WIDTH == 8) {
        v[j][0] = __ldg((const unsigned long long *) table + s);
      } else {
        const uint64_t *p = table + (s & ~3ull);
        asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[j][0]), "=l"(v[j][1]), "=l"(v[j][2]), "=l"(v[j][3]) : "l"(p));
      }
    }
#pragma unroll
    for (int j = 0; j < KPT; ++j)
      for (int q = 0; q < WIDTH / 8; ++q) acc += v[j][q];
  }
  if (acc == 0x1234567) sink[0] = acc;
}

// the probe's real shape: streamed key in, random gather, streamed 16 B out
template <int KPT>
__global__ void probe_like_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ table, uint64_t mask, size_t n,
                                  uint64_t *out_a, uint64_t *out_b) {
  size_t tile = (size_t) blockDim.x * KPT;
  for (size_t base = (size_t) blockIdx.x * tile; base < n; base += (size_t) gridDim.x * tile) {
    uint64_t k[KPT], v[KPT];
#pragma unroll
    for (int j = 0; j < KPT; ++j) { size_t i = base + j * blockDim.x + threadIdx.x; k[j] = i < n ? __ldg((const unsigned long long *) keys + i) : 0; }
#pragma unroll
    for (int j = 0; j < KPT; ++j) v[j] = __ldg((const unsigned long long *) table + (mm(k[j]) & mask));
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

int main(int argc, char **argv) {
  int max_log2 = argc > 1 ? atoi(argv[1]) : 30;  // table slots (8 B each)
  size_t n = (size_t) 1 << 29;                   // accesses per measurement
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("# %s, %d SMs, L2 %d MB\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20);
  uint64_t *table, *sink, *keys, *oa, *ob;
  CK(cudaMalloc(&table, ((size_t) 1 << max_log2) * 8)); CK(cudaMalloc(&sink, 64));
  fill_kernel<<<148 * 8, 256>>>(table, (size_t) 1 << max_log2);
  CK(cudaMalloc(&keys, n * 8)); CK(cudaMalloc(&oa, n * 8)); CK(cudaMalloc(&ob, n * 8));
  fill_kernel<<<148 * 8, 256>>>(keys, n);
  CK(cudaDeviceSynchronize());
  printf("footprint_MiB, gather8_kpt4_G/s, gather8_kpt8_G/s, gather32_kpt4_G/s, gather32_kpt8_G/s, probe_like_kpt4_G/s, probe_like_kpt8_G/s\n");
  for (int lg = 20; lg <= max_log2; lg += 1) {
    if (lg > 24 && lg < max_log2 && (lg & 1)) continue;
    uint64_t mask = ((uint64_t) 1 << lg) - 1;
    int grid = 148 * 8;
    float a = time_ms([&] { gather_kernel<4, 8><<<grid, 256>>>(table, mask, n, sink); });
    float b = time_ms([&] { gather_kernel<8, 8><<<grid, 256>>>(table, mask, n, sink); });
    float c = time_ms([&] { gather_kernel<4, 32><<<grid, 256>>>(table, mask, n, sink); });
    float d = time_ms([&] { gather_kernel<8, 32><<<grid, 256>>>(table, mask, n, sink); });
    float e = time_ms([&] { probe_like_kernel<4><<<grid, 256>>>(keys, table, mask, n, oa, ob); });
    float f = time_ms([&] { probe_like_kernel<8><<<grid, 256>>>(keys, table, mask, n, oa, ob); });
    printf("%zu, %.1f, %.1f, %.1f, %.1f, %.1f, %.1f\n", (((size_t) 1 << lg) * 8) >> 20, n / a / 1e6, n / b / 1e6, n / c / 1e6, n / d / 1e6, n / e / 1e6, n / f / 1e6);
    fflush(stdout);
  }
  // occupancy sweep at the largest footprint
  {
    uint64_t mask = ((uint64_t) 1 << max_log2) - 1;
    printf("# blocks/SM sweep at %zu MiB (gather8 kpt8, 256 thr): ", (((size_t) 1 << max_log2) * 8) >> 20);
    for (int bps : {2, 4, 6, 8}) { float t = time_ms([&] { gather_kernel<8, 8><<<148 * bps, 256>>>(table, mask, n, sink); }); printf("%d:%.1f ", bps, n / t / 1e6); }
    printf("G/s\n");
  }
  // L2 fetch granularity (cudaLimitMaxL2FetchGranularity): does a 32-byte granularity lift the big-table gather rate?
  for (size_t gran : {(size_t) 32, (size_t) 64, (size_t) 128}) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    size_t got = 0;
    cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    uint64_t mask = ((uint64_t) 1 << max_log2) - 1;
    float t = time_ms([&] { gather_kernel<8, 8><<<148 * 8, 256>>>(table, mask, n, sink); });
    float t2 = time_ms([&] { probe_like_kernel<4><<<148 * 8, 256>>>(keys, table, mask, n, oa, ob); });
    printf("# L2 fetch granularity request %zu -> %s, limit now %zu: %zu MiB gather8 %.1f G/s, probe_like %.1f G/s\n", gran, cudaGetErrorString(e), got,
           (((size_t) 1 << max_log2) * 8) >> 20, n / t / 1e6, n / t2 / 1e6);
  }
  // does the shared-memory carve-out (smaller L1) change the L2 sector traffic / rate of a gather?
  {
    uint64_t mask = ((uint64_t) 1 << 21) - 1;  // 16 MiB footprint
    for (int smem_kb : {0, 16, 40, 100}) {
      CK(cudaFuncSetAttribute(gather_kernel<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      CK(cudaFuncSetAttribute(probe_like_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      float t = time_ms([&] { gather_kernel<4, 8><<<148 * 4, 256, smem_kb * 1024>>>(table, mask, n, sink); });
      float t2 = time_ms([&] { probe_like_kernel<4><<<148 * 4, 256, smem_kb * 1024>>>(keys, table, mask, n, oa, ob); });
      printf("# 16 MiB footprint, %3d KB dyn smem/CTA (4 CTAs/SM): gather8 %.1f G/s, probe_like %.1f G/s\n", smem_kb, n / t / 1e6, n / t2 / 1e6);
    }
  }
  // PCIe
  {
    size_t bytes = (size_t) 1 << 30;
    void *h; CK(cudaMallocHost(&h, bytes));
    float t1 = time_ms([&] { CK(cudaMemcpyAsync(keys, h, bytes, cudaMemcpyHostToDevice)); });
    float t2 = time_ms([&] { CK(cudaMemcpyAsync(h, keys, bytes, cudaMemcpyDeviceToHost)); });
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    void *h2; CK(cudaMallocHost(&h2, bytes));
    float t3 = time_ms([&] { CK(cudaMemcpyAsync(keys, h, bytes, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(h2, oa, bytes, cudaMemcpyDeviceToHost, s2)); CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2)); });
    printf("# PCIe pinned 1 GiB: H2D %.1f GB/s, D2H %.1f GB/s, both directions at once %.1f GB/s total\n", bytes / t1 / 1e6, bytes / t2 / 1e6, 2.0 * bytes / t3 / 1e6);
  }
  return 0;
}
