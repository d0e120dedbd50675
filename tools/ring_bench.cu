// Micro-benchmark: how fast can a B200 stream 335 MB through (a) a cp.async.bulk ring into shared memory,
// (b) plain 128-bit global loads?  Decides the structure of tile_topk_kernel.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <algorithm>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* d, const void* s, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d)), "l"(s), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(NT) ring_kernel(const float* __restrict__ in, long long n_chunks, int chunk_floats, int stages, float thr, int* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* ring = (float*)smem;
  uint64_t* bars = (uint64_t*)(smem + (size_t)stages * chunk_floats * 4);
  const int tid = threadIdx.x;
  if (tid == 0) { for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  long long p = blockIdx.x; int pslot = 0;
  auto issue = [&]() { if (p < n_chunks) { mbar_expect_tx(&bars[pslot], chunk_floats * 4); bulk_g2s(ring + (size_t)pslot * chunk_floats, in + p * chunk_floats, chunk_floats * 4, &bars[pslot]); p += gridDim.x; if (++pslot == stages) pslot = 0; } };
  if (tid == 0) for (int s = 0; s < stages; ++s) issue();
  int slot = 0; uint32_t phase = 0; int hits = 0;
  for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    mbar_wait(&bars[slot], phase);
    const float4* p4 = (const float4*)(ring + (size_t)slot * chunk_floats);
    for (int t = tid; t < chunk_floats / 4; t += NT) { float4 x = p4[t]; if (fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)) >= thr) ++hits; }
    __syncthreads();
    if (tid == 0) issue();
    if (++slot == stages) { slot = 0; phase ^= 1; }
  }
  if (hits) atomicAdd(out, hits);
}

template <int NT, int U>
__global__ void __launch_bounds__(NT) ldg_kernel(const float4* __restrict__ in, long long n4, float thr, int* out) {
  int hits = 0;
  const long long stride = (long long)gridDim.x * NT;
  long long i = blockIdx.x * (long long)NT + threadIdx.x;
  for (; i + (U - 1) * stride < n4; i += U * stride) {
    float4 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x[u].x), "=f"(x[u].y), "=f"(x[u].z), "=f"(x[u].w) : "l"(in + i + u * stride));
#pragma unroll
    for (int u = 0; u < U; ++u) if (fmaxf(fmaxf(x[u].x, x[u].y), fmaxf(x[u].z, x[u].w)) >= thr) ++hits;
  }
  for (; i < n4; i += stride) { float4 x = in[i]; if (fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)) >= thr) ++hits; }
  if (hits) atomicAdd(out, hits);
}

int main() {
  const long long n = 64LL * 80 * 128 * 128;
  float* d; int* out; float* flush;
  cudaMalloc(&d, n * 4); cudaMalloc(&out, 4); cudaMalloc(&flush, 256 << 20);
  cudaMemset(d, 0, n * 4); cudaMemset(out, 0, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](auto launch, const char* name) {
    std::vector<float> ts;
    for (int r = 0; r < 7; ++r) { cudaMemsetAsync(flush, r, 256 << 20); cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); ts.push_back(ms); }
    cudaError_t e = cudaGetLastError(); std::sort(ts.begin(), ts.end());
    printf("%-52s %8.1f us  %7.0f GB/s  %s\n", name, ts[3] * 1e3, n * 4 / (ts[3] * 1e-3) / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  int sms = 148;
  for (int chunk_kb : {8, 16, 32}) for (int stages : {2, 3, 4, 6}) for (int per_sm : {1, 2, 3, 4}) {
    int chunk_floats = chunk_kb * 256; size_t smem = (size_t)stages * chunk_kb * 1024 + 64;
    if (smem * per_sm > 220 * 1024) continue;
    long long n_chunks = n / chunk_floats;
    cudaFuncSetAttribute(ring_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    char name[128]; snprintf(name, 128, "ring  chunk=%2dKB stages=%d ctas/sm=%d (inflight %3d KB/SM)", chunk_kb, stages, per_sm, (stages - 1) * chunk_kb * per_sm);
    // force occupancy via smem padding is unreliable; use grid = per_sm * sms (persistent) and rely on scheduler
    timeit([&]() { ring_kernel<256><<<per_sm * sms, 256, smem>>>(d, n_chunks, chunk_floats, stages, 1.0f, out); }, name);
  }
  timeit([&]() { ldg_kernel<256, 4><<<sms * 8, 256>>>((const float4*)d, n / 4, 1.0f, out); }, "ldg   256thr U=4 grid=8/SM");
  timeit([&]() { ldg_kernel<256, 8><<<sms * 8, 256>>>((const float4*)d, n / 4, 1.0f, out); }, "ldg   256thr U=8 grid=8/SM");
  timeit([&]() { ldg_kernel<512, 4><<<sms * 4, 512>>>((const float4*)d, n / 4, 1.0f, out); }, "ldg   512thr U=4 grid=4/SM");
  timeit([&]() { ldg_kernel<256, 2><<<sms * 8, 256>>>((const float4*)d, n / 4, 1.0f, out); }, "ldg   256thr U=2 grid=8/SM");
  timeit([&]() { ldg_kernel<1024, 4><<<sms * 2, 1024>>>((const float4*)d, n / 4, 1.0f, out); }, "ldg  1024thr U=4 grid=2/SM");
  timeit([&]() { cudaMemcpyAsync(flush, d, n * 4, cudaMemcpyDeviceToDevice); }, "cudaMemcpy D2D (r+w, GB/s counts read only)");
  return 0;
}
