// runtime.cu -- device bring-up, error reporting, memory/stream helpers, small
// utility kernels (hash, generators).  No CPU fallback: every compute entry
// point goes through require_device().
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace ccb {

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};
static int g_sm_count = 0;
static uint64_t g_scratch_keep = ~0ull;  // release threshold of the stream-ordered scratch pool (cc_scratch_set_retention)

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
  if (g_sm_count > 0) return g_sm_count;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
    g_sm_count = n;
  else
    g_sm_count = kSmCountFallback;
  return g_sm_count;
}

int require_device() {
  static thread_local int checked_dev = -1;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); libccb200 has no CPU fallback", cudaGetErrorString(e));
    return CC_ERR_NO_DEVICE;
  }
  if (dev == checked_dev) return CC_OK;
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); libccb200 has no CPU fallback", cudaGetErrorString(e));
    return CC_ERR_NO_DEVICE;
  }
  if (major != 10) {
    set_error("device %d has compute capability %d.x; libccb200 is built for sm_100a (B200) only", dev, major);
    return CC_ERR_NO_DEVICE;
  }
  checked_dev = dev;
  return CC_OK;
}

// ---- utility kernels --------------------------------------------------------------
__global__ void hash_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, size_t n) {
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = murmurhash64(in[i]);
}

// chaining_ht.cpp:15-26: key of build row r is (r / cf) * step with step = n / num_unique.
__global__ void gen_build_keys_kernel(int64_t *__restrict__ keys, size_t n, size_t cf, size_t step, size_t first = 0) {
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) keys[i] = (int64_t) (((first + i) / cf) * step);
}

__global__ void gen_keys_counter_kernel(int64_t *__restrict__ keys, size_t n, uint64_t seed, uint64_t first, uint64_t mask) {
  size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    keys[i] = (int64_t) (murmurhash64(seed + first + i) & mask);
}

static int grid_for(size_t n, int threads, int per_sm = 8) {
  size_t blocks = (n + threads - 1) / threads;
  size_t cap = (size_t) sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  return (int) blocks;
}

}  // namespace ccb

using namespace ccb;

extern "C" {

int cc_api_version(void) { return CC_API_VERSION; }
const char *cc_last_error(void) { return g_err; }
uint64_t cc_launch_count(void) { return g_launches.load(); }

int cc_device_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); libccb200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return CC_ERR_NO_DEVICE;
  }
  CC_REQUIRE(device >= 0 && device < n, "device %d out of range [0,%d)", device, n);
  CC_CUDA(cudaSetDevice(device));
  CC_TRY(require_device());
  g_sm_count = 0;
  sm_count();
  // keep stream-ordered scratch (cudaMallocAsync in the partitioned probe) cached between calls; cc_scratch_set_retention
  // bounds what stays cached and cc_scratch_release hands it back (a host that shares the GPU with another allocator)
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t keep = g_scratch_keep;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  cudaGetLastError();
  return CC_OK;
}

int cc_scratch_set_retention(uint64_t bytes) {
  CC_TRY(require_device());
  g_scratch_keep = bytes;
  int dev = 0;
  CC_CUDA(cudaGetDevice(&dev));
  cudaMemPool_t pool;
  CC_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
  uint64_t keep = bytes;
  CC_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  return CC_OK;
}

int cc_scratch_release(void) {
  CC_TRY(require_device());
  int dev = 0;
  CC_CUDA(cudaGetDevice(&dev));
  cudaMemPool_t pool;
  CC_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
  CC_CUDA(cudaDeviceSynchronize());
  CC_CUDA(cudaMemPoolTrimTo(pool, 0));
  return CC_OK;
}

int cc_device_get_info(cc_device_info *info) {
  CC_REQUIRE(info, "info is NULL");
  CC_TRY(require_device());
  int dev = 0;
  CC_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  CC_CUDA(cudaGetDeviceProperties(&p, dev));
  memset(info, 0, sizeof(*info));
  info->device = dev;
  info->sm_major = p.major;
  info->sm_minor = p.minor;
  info->sm_count = p.multiProcessorCount;
  info->l2_bytes = (size_t) p.l2CacheSize;
  size_t fr = 0, tot = 0;
  CC_CUDA(cudaMemGetInfo(&fr, &tot));
  info->total_mem = tot;
  info->free_mem = fr;
  strncpy(info->name, p.name, sizeof(info->name) - 1);
  return CC_OK;
}

int cc_malloc(void **d_ptr, size_t bytes) {
  CC_REQUIRE(d_ptr, "d_ptr is NULL");
  CC_TRY(require_device());
  *d_ptr = nullptr;
  if (bytes == 0) bytes = 16;
  CC_CUDA(cudaMalloc(d_ptr, bytes));
  return CC_OK;
}
int cc_free(void *d_ptr) {
  if (!d_ptr) return CC_OK;
  CC_CUDA(cudaFree(d_ptr));
  return CC_OK;
}
int cc_host_alloc(void **h_ptr, size_t bytes) {
  CC_REQUIRE(h_ptr, "h_ptr is NULL");
  CC_TRY(require_device());
  CC_CUDA(cudaMallocHost(h_ptr, bytes ? bytes : 16));
  return CC_OK;
}
int cc_host_free(void *h_ptr) {
  if (!h_ptr) return CC_OK;
  CC_CUDA(cudaFreeHost(h_ptr));
  return CC_OK;
}
int cc_memcpy_h2d(void *d, const void *h, size_t bytes, cc_stream_t s) {
  CC_TRY(require_device());
  if (bytes) CC_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, as_stream(s)));
  return CC_OK;
}
int cc_memcpy_d2h(void *h, const void *d, size_t bytes, cc_stream_t s) {
  CC_TRY(require_device());
  if (bytes) CC_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, as_stream(s)));
  return CC_OK;
}
int cc_memcpy_d2d(void *dd, const void *ds, size_t bytes, cc_stream_t s) {
  CC_TRY(require_device());
  if (bytes) CC_CUDA(cudaMemcpyAsync(dd, ds, bytes, cudaMemcpyDeviceToDevice, as_stream(s)));
  return CC_OK;
}
int cc_memset(void *d, int byte, size_t bytes, cc_stream_t s) {
  CC_TRY(require_device());
  if (bytes) CC_CUDA(cudaMemsetAsync(d, byte, bytes, as_stream(s)));
  return CC_OK;
}
int cc_stream_create(cc_stream_t *s) {
  CC_REQUIRE(s, "stream is NULL");
  CC_TRY(require_device());
  cudaStream_t st;
  CC_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  *s = st;
  return CC_OK;
}
int cc_stream_destroy(cc_stream_t s) {
  if (s) CC_CUDA(cudaStreamDestroy(as_stream(s)));
  return CC_OK;
}
int cc_stream_sync(cc_stream_t s) {
  CC_TRY(require_device());
  CC_CUDA(cudaStreamSynchronize(as_stream(s)));
  return CC_OK;
}

int cc_hash_u64(const uint64_t *d_in, uint64_t *d_out, size_t n, cc_stream_t s) {
  CC_TRY(require_device());
  if (n == 0) return CC_OK;
  CC_REQUIRE(d_in && d_out, "NULL buffer");
  hash_kernel<<<grid_for(n, 256), 256, 0, as_stream(s)>>>(d_in, d_out, n);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_gen_build_keys(int64_t *d_keys, size_t n, size_t cf, cc_stream_t s) {
  CC_TRY(require_device());
  if (n == 0) return CC_OK;
  CC_REQUIRE(d_keys && cf > 0, "NULL buffer or chunk_factor == 0");
  size_t num_unique = n / cf + (n % cf != 0);
  size_t step = n / num_unique;
  gen_build_keys_kernel<<<grid_for(n, 256), 256, 0, as_stream(s)>>>(d_keys, n, cf, step);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_gen_build_keys_range(int64_t *d_keys, size_t first, size_t count, size_t n_total, size_t cf, cc_stream_t s) {
  CC_TRY(require_device());
  if (count == 0) return CC_OK;
  CC_REQUIRE(d_keys && cf > 0, "NULL buffer or chunk_factor == 0");
  CC_REQUIRE(first + count <= n_total, "rows [%zu, %zu) exceed the %zu build rows", first, first + count, n_total);
  size_t num_unique = n_total / cf + (n_total % cf != 0);
  size_t step = n_total / num_unique;
  gen_build_keys_kernel<<<grid_for(count, 256), 256, 0, as_stream(s)>>>(d_keys, count, cf, step, first);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

int cc_gen_keys_counter(int64_t *d_keys, size_t n, uint64_t seed, uint64_t first, uint64_t mask, cc_stream_t s) {
  CC_TRY(require_device());
  if (n == 0) return CC_OK;
  CC_REQUIRE(d_keys, "NULL buffer");
  gen_keys_counter_kernel<<<grid_for(n, 256), 256, 0, as_stream(s)>>>(d_keys, n, seed, first, mask);
  CC_CHECK_LAUNCH();
  return CC_OK;
}

}  // extern "C"
