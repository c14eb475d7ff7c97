// Read-only streaming ceilings on one GPU: what a matrix pass can reach before any arithmetic is attached.
//   A: LDG.128 (ld.global.nc.L1::no_allocate) grid-stride, U independent loads in flight per thread
//   B: cp.async.bulk (TMA 1-D) into a shared-memory ring, mbarrier full/empty, consumers read LDS.128
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o stream_probe stream_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

template <int U>
__global__ void __launch_bounds__(256) k_ldg(const uint4* __restrict__ src, size_t n, uint32_t* out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ldg_stream(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < n; i += stride) { uint4 v = ldg_stream(src + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678u) out[0] = acc;
}

// ---- TMA ring -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// tiles of TILE bytes handed to CTAs round-robin; warp 0 lane 0 = producer, all CONS threads consume
template <int TILE, int STAGES, int CONS>
__global__ void __launch_bounds__(CONS + 32) k_tma(const uint8_t* __restrict__ src, size_t ntiles, uint32_t* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if (warp == CONS / 32) {   // producer warp
    if ((threadIdx.x & 31) == 0) {
      int s = 0; uint32_t ph = 0;
      for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], TILE);
        tma_load_1d(smem + (size_t)s * TILE, src + t * TILE, TILE, &full[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    uint32_t acc = 0;
    int s = 0; uint32_t ph = 0;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      mbar_wait(&full[s], ph);
      const uint4* p = reinterpret_cast<const uint4*>(smem + (size_t)s * TILE);
#pragma unroll 4
      for (int i = threadIdx.x; i < TILE / 16; i += CONS) { uint4 v = p[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    if (acc == 0x12345678u) out[0] = acc;
  }
}

// row-list variant: a stage = 16 row segments of SEG bytes (one bulk copy each), rows taken from a list; column tile x = blockIdx.x
template <int SEG, int STAGES, int CONS, bool PAR>
__global__ void __launch_bounds__(CONS + 32) k_tma_rows(const uint8_t* __restrict__ src, size_t row_bytes, const int* __restrict__ rows, int nrows_per_cta,
                                                        uint32_t* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  constexpr int TILE = 16 * SEG;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int* myrows = rows + (size_t)blockIdx.y * nrows_per_cta;
  const uint8_t* base = src + (size_t)blockIdx.x * SEG;
  const int nb = nrows_per_cta / 16;
  if (warp == CONS / 32) {
    int s = 0; uint32_t ph = 0;
    int rnext = (lane < 16) ? myrows[lane] : 0;
    for (int q = 0; q < nb; ++q) {
      const int r = rnext;
      if (q + 1 < nb && lane < 16) rnext = myrows[(q + 1) * 16 + lane];
      if (PAR) {
        if (lane == 0) { mbar_wait(&empty[s], ph ^ 1); mbar_expect_tx(&full[s], TILE); }
        __syncwarp();
        if (lane < 16) tma_load_1d(smem + (size_t)s * TILE + lane * SEG, base + (size_t)r * row_bytes, SEG, &full[s]);
      } else {
        if (lane == 0) { mbar_wait(&empty[s], ph ^ 1); mbar_expect_tx(&full[s], TILE); }
        for (int j = 0; j < 16; ++j) {
          const int rj = __shfl_sync(0xffffffffu, r, j);
          if (lane == 0) tma_load_1d(smem + (size_t)s * TILE + j * SEG, base + (size_t)rj * row_bytes, SEG, &full[s]);
        }
      }
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else {
    uint32_t acc = 0;
    int s = 0; uint32_t ph = 0;
    for (int q = 0; q < nb; ++q) {
      mbar_wait(&full[s], ph);
      const uint2* p = reinterpret_cast<const uint2*>(smem + (size_t)s * TILE);
#pragma unroll
      for (int j = 0; j < 16; ++j) if (threadIdx.x * 8 < SEG) { uint2 v = p[j * (SEG / 8) + threadIdx.x]; acc ^= v.x ^ v.y; }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    if (acc == 0x12345678u) out[0] = acc;
  }
}

template <typename F>
static float time_ms(F f, int reps = 20) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  std::vector<float> ts;
  for (int i = 0; i < reps; ++i) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); ts.push_back(ms); }
  std::sort(ts.begin(), ts.end());
  CK(cudaGetLastError());
  return ts[ts.size() / 2];
}

template <int U> static void run_ldg(const uint4* d, size_t n, uint32_t* out, int sms) {
  for (int per : {2, 4, 6, 8}) {
    float ms = time_ms([&] { k_ldg<U><<<sms * per, 256>>>(d, n, out); });
    printf("ldg U=%d ctas/sm=%d (256 thr): %.3f ms  %.0f GB/s\n", U, per, ms, n * 16.0 / ms / 1e6);
  }
}

template <int TILE, int STAGES, int CONS> static void run_tma(const uint8_t* d, size_t bytes, uint32_t* out, int sms) {
  size_t ntiles = bytes / TILE;
  size_t sm = (size_t)TILE * STAGES;
  CK(cudaFuncSetAttribute(k_tma<TILE, STAGES, CONS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  int maxper = (int)std::min<size_t>(8, (220 * 1024) / sm);
  for (int per = 1; per <= maxper; per *= 2) {
    float ms = time_ms([&] { k_tma<TILE, STAGES, CONS><<<sms * per, CONS + 32, sm>>>(d, ntiles, out); });
    printf("tma tile=%d stages=%d cons=%d ctas/sm=%d (%zu KB in flight/SM): %.3f ms  %.0f GB/s\n", TILE, STAGES, CONS, per, sm * per / 1024, ms,
           ntiles * (double)TILE / ms / 1e6);
  }
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  const size_t bytes = (size_t)200000 * 6256;   // the bench matrix: 200 000 rows x 1564 words
  uint8_t* d; uint32_t* out;
  CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&out, 4)); CK(cudaMemset(d, 0x5a, bytes));
  uint8_t* d2; CK(cudaMalloc(&d2, bytes));
  float ms = time_ms([&] { CK(cudaMemcpyAsync(d2, d, bytes, cudaMemcpyDeviceToDevice)); });
  printf("%s, %d SMs; memcpy d2d %.3f ms -> %.0f GB/s (read+write)\n", p.name, sms, ms, 2.0 * bytes / ms / 1e6);
  const size_t n = bytes / 16;
  run_ldg<1>((const uint4*)d, n, out, sms);
  run_ldg<2>((const uint4*)d, n, out, sms);
  run_ldg<4>((const uint4*)d, n, out, sms);
  run_ldg<8>((const uint4*)d, n, out, sms);
  run_tma<8192, 4, 128>(d, bytes, out, sms);
  run_tma<16384, 4, 128>(d, bytes, out, sms);
  run_tma<16384, 3, 256>(d, bytes, out, sms);
  run_tma<32768, 3, 256>(d, bytes, out, sms);
  run_tma<6256 * 4 / 4 * 4, 4, 128>(d, bytes, out, sms);   // 4 matrix rows per tile (25 024 B)
  {  // one-hot M-step access pattern: 200 000 rows x 2 column tiles of 3136 / 3120 B, 74 row slices
    const int N = 200000, S = 74, per = (N / S) / 16 * 16;
    std::vector<int> asc(N), rnd(N), chunk(N);
    for (int i = 0; i < N; ++i) asc[i] = rnd[i] = i;
    srand(1);
    for (int i = N - 1; i > 0; --i) std::swap(rnd[i], rnd[rand() % (i + 1)]);
    {  // 32-row chunks in random order (what a warp-aggregated scatter produces)
      std::vector<int> c(N / 32); for (int i = 0; i < N / 32; ++i) c[i] = i;
      for (int i = N / 32 - 1; i > 0; --i) std::swap(c[i], c[rand() % (i + 1)]);
      for (int i = 0; i < N / 32; ++i) for (int j = 0; j < 32; ++j) chunk[i * 32 + j] = c[i] * 32 + j;
    }
    int* drows; CK(cudaMalloc(&drows, N * sizeof(int)));
    constexpr int SEG = 3120, ST = 3, CONS = 416;
    size_t sm = (size_t)16 * SEG * ST;
    CK(cudaFuncSetAttribute(k_tma_rows<SEG, ST, CONS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    CK(cudaFuncSetAttribute(k_tma_rows<SEG, ST, CONS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const char* names[3] = {"ascending", "32-row chunks shuffled", "fully shuffled"};
    std::vector<int>* lists[3] = {&asc, &chunk, &rnd};
    for (int l = 0; l < 3; ++l) {
      CK(cudaMemcpy(drows, lists[l]->data(), N * sizeof(int), cudaMemcpyHostToDevice));
      const double bytes_moved = (double)S * per * 2 * SEG;
      float m0 = time_ms([&] { k_tma_rows<SEG, ST, CONS, false><<<dim3(2, S), CONS + 32, sm>>>(d, 6256, drows, per, out); });
      float m1 = time_ms([&] { k_tma_rows<SEG, ST, CONS, true><<<dim3(2, S), CONS + 32, sm>>>(d, 6256, drows, per, out); });
      printf("row list %-24s: one issuing lane %.3f ms %.0f GB/s | 16 issuing lanes %.3f ms %.0f GB/s\n", names[l], m0, bytes_moved / m0 / 1e6, m1,
             bytes_moved / m1 / 1e6);
    }
  }
  return 0;
}
