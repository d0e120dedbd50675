// common.cuh — shared host/device helpers for libtauv_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/tauv_b200.h"

namespace tauv {

// ----------------------------------------------------------------------------------------------
// Host-side error plumbing (thread-local message, C return codes)
// ----------------------------------------------------------------------------------------------
char* last_error_buf();  // 512-byte thread-local buffer (api.cu)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline int cuda_fail(cudaError_t e, const char* what) {
  snprintf(last_error_buf(), 512, "%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}

#define TAUV_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) return ::tauv::fail((code), __VA_ARGS__); \
  } while (0)

#define TAUV_CUDA(expr)                                         \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) return ::tauv::cuda_fail(_e, #expr); \
  } while (0)

#define TAUV_LAUNCH_CHECK(name)                                   \
  do {                                                            \
    cudaError_t _e = cudaGetLastError();                          \
    if (_e != cudaSuccess) return ::tauv::cuda_fail(_e, name);    \
  } while (0)

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int num_sms();  // cached per device (api.cu)
cudaError_t ensure_dynamic_smem(const void* func, size_t bytes);  // thread-safe, only ever raises (api.cu)

// Experiment switches (environment variables) exist only in -DTAUV_DEBUG builds: the default library reads no
// environment on the call path and keeps no mutable process state.
#ifdef TAUV_DEBUG
inline bool debug_env(const char* name) { return getenv(name) != nullptr; }
#else
inline bool debug_env(const char*) { return false; }
#endif

// ----------------------------------------------------------------------------------------------
// Device helpers
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__

#define TAUV_NEG_INF (__int_as_float(0xff800000))

// Order-preserving map float -> uint32 (larger float => larger key).  +-0 collapse to one key so
// that they tie like torch.topk treats them; +NaN sorts above +inf (torch.topk ranks NaN first).
__device__ __forceinline__ uint32_t float_to_key(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u << 1) == 0u) u = 0u;  // -0 -> +0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}
// Composite sort key: score key in the high word, complemented flat index in the low word, so a
// plain descending order on the 64-bit value is (score desc, index asc) and all keys are distinct.
__device__ __forceinline__ unsigned long long make_composite(uint32_t key, uint32_t idx) {
  return ((unsigned long long)key << 32) | (unsigned long long)(~idx);
}
__device__ __forceinline__ uint32_t composite_key(unsigned long long c) { return (uint32_t)(c >> 32); }
__device__ __forceinline__ uint32_t composite_idx(unsigned long long c) { return ~(uint32_t)c; }

// The reference's sigmoid on CPU is 1/(1+exp(-x)) with a <=1ulp exp and an exact divide
// (ATen UnaryOpsKernel sigmoid).  Same formula, IEEE divide, libdevice expf (<=2 ulp).
__device__ __forceinline__ float sigmoid_ref(float x) {
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

// ---- mbarrier / bulk-copy PTX (cp.async.bulk: 1-D TMA, no tensor map needed) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy, completion counted in bytes on `bar`.  16-byte aligned src/dst/size.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global bulk store (bulk_group completion)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// streaming (read-once / write-once) global accesses
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// Block-wide exact selection of the k-th largest of n DISTINCT 64-bit keys (MSB-first radix select,
// 11-bit digits, early exit as soon as the chosen bin is wholly needed).
//   load(i) -> key i (any memory);   hist: 2048 uint32 in shared;   ctl: 4 uint32 in shared.
// Returns T such that exactly min(k,n) keys are >= T.  All NT threads of the group must call it
// (it synchronises through Sync) and all get the same T.  NT must be a multiple of 32 and <= 1024.
// ----------------------------------------------------------------------------------------------
constexpr int kRadixBits = 11;
constexpr int kRadixBins = 1 << kRadixBits;

// Barrier policy: the whole CTA (default), or a named barrier over the first NT threads of a warp-specialised CTA
// (the other warps — e.g. a TMA producer — never take part).
struct SyncBlock {
  __device__ static __forceinline__ void sync() { __syncthreads(); }
};
template <int ID, int NT>
struct SyncNamed {
  __device__ static __forceinline__ void sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NT) : "memory"); }
};

template <int NT, class LoadFn, class Sync = SyncBlock>
__device__ unsigned long long block_kth_largest(LoadFn load, int n, int k, uint32_t* hist,
                                                uint32_t* ctl) {
  if (k >= n) return 0ull;
  const int tid = threadIdx.x;
  constexpr int BPT = kRadixBins / NT > 0 ? kRadixBins / NT : 1;  // bins per thread
  constexpr int NSCAN = kRadixBins / BPT;                         // threads taking part in the scan
  unsigned long long prefix = 0ull;  // selected high bits so far (right-aligned)
  int hi = 64;                       // bits [hi,64) of the threshold are fixed in `prefix`
  int k_rem = k;
  __shared__ uint32_t warp_tot[32];
  __shared__ unsigned long long s_and, s_or;
  // Skip the leading bits every key shares (scores of one frame sit in one or two binades, so the
  // first 11-bit digit would otherwise be wasted): one AND/OR sweep costs less than one radix pass.
  {
    if (tid == 0) {
      s_and = ~0ull;
      s_or = 0ull;
    }
    Sync::sync();
    unsigned long long a = ~0ull, o = 0ull;
    for (int i = tid; i < n; i += NT) {
      const unsigned long long c = load(i);
      a &= c;
      o |= c;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      a &= __shfl_xor_sync(0xffffffffu, a, d);
      o |= __shfl_xor_sync(0xffffffffu, o, d);
    }
    if ((tid & 31) == 0) {
      atomicAnd(&s_and, a);
      atomicOr(&s_or, o);
    }
    Sync::sync();
    const unsigned long long diff = s_and ^ s_or;
    hi = diff ? 64 - __clzll((long long)diff) : 0;  // highest differing bit + 1
    prefix = hi < 64 ? (s_or >> hi) : 0ull;
    if (hi == 0) return s_or;  // all keys equal (n == 1 after the k >= n test cannot happen; defensive)
  }
  while (hi > 0) {
    const int bits = hi >= kRadixBits ? kRadixBits : hi;
    const int lo = hi - bits;
    for (int i = tid; i < kRadixBins; i += NT) hist[i] = 0;
    Sync::sync();
    for (int i = tid; i < n; i += NT) {
      unsigned long long c = load(i);
      bool match = (hi == 64) ? true : ((c >> hi) == prefix);
      if (match) atomicAdd(&hist[(uint32_t)(c >> lo) & ((1u << bits) - 1u)], 1u);
    }
    Sync::sync();
    // suffix scan: thread t owns bins [t*BPT, (t+1)*BPT); find the largest bin g with
    // count(bins >= g) >= k_rem.
    uint32_t local[BPT];
    uint32_t mine = 0;
    if (tid < NSCAN) {
#pragma unroll
      for (int j = 0; j < BPT; ++j) {
        local[j] = hist[tid * BPT + j];
        mine += local[j];
      }
    }
    // inclusive suffix sum of `mine` across threads (higher tid = higher bins)
    uint32_t suf = mine;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t v = __shfl_down_sync(0xffffffffu, suf, o);
      if (lane + o < 32) suf += v;
    }
    if (lane == 0) warp_tot[warp] = suf;
    Sync::sync();
    uint32_t above_warps = 0;
    for (int w = warp + 1; w < NT / 32; ++w) above_warps += warp_tot[w];
    suf += above_warps;                 // count in bins >= my first bin
    const uint32_t above = suf - mine;  // count in bins >  my last bin
    if (tid < NSCAN && above < (uint32_t)k_rem && suf >= (uint32_t)k_rem) {
      uint32_t acc = above;
#pragma unroll
      for (int j = BPT - 1; j >= 0; --j) {
        if (acc + local[j] >= (uint32_t)k_rem) {
          ctl[0] = (uint32_t)(tid * BPT + j);  // selected digit
          ctl[1] = (uint32_t)k_rem - acc;      // still needed inside the bin
          ctl[2] = local[j];                   // bin population
          break;
        }
        acc += local[j];
      }
    }
    Sync::sync();
    const uint32_t digit = ctl[0];
    k_rem = (int)ctl[1];
    const uint32_t pop = ctl[2];
    prefix = (hi == 64) ? (unsigned long long)digit : ((prefix << bits) | digit);
    hi = lo;
    Sync::sync();  // ctl / hist reused next round
    if ((uint32_t)k_rem == pop) break;  // every key in the bin is needed
  }
  return hi == 0 ? prefix : (prefix << hi);
}

// In-place bitonic sort (descending) of n = power-of-two 64-bit keys in shared memory.
template <int NT>
__device__ void block_bitonic_sort_desc(unsigned long long* a, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += NT) {
        int i = 2 * t - (t & (stride - 1));
        int j = i + stride;
        bool desc = ((i & size) == 0);
        unsigned long long x = a[i], y = a[j];
        if ((x < y) == desc) {
          a[i] = y;
          a[j] = x;
        }
      }
    }
  }
  __syncthreads();
}

// The same for n <= NT: one key per thread, held in a register.  Compare-exchange steps with a partner inside the warp
// (stride < 32) are two shuffles; only the strides >= 32 go through shared memory and a barrier — 6 of the 36 steps for
// 256 keys (the all-shared-memory version spent ~290 cycles per step on its barrier: 10.4 k cycles of the NMS kernel).
template <int NT>
__device__ void block_bitonic_sort_desc_reg(unsigned long long* a, int n) {
  const int t = threadIdx.x;
  if (n > NT) {
    block_bitonic_sort_desc<NT>(a, n);
    return;
  }
  __syncthreads();
  unsigned long long x = t < n ? a[t] : 0ull;
  for (int size = 2; size <= n; size <<= 1) {
    const bool desc = (t & size) == 0;
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      unsigned long long y;
      if (stride >= 32) {
        if (t < n) a[t] = x;
        __syncthreads();
        y = t < n ? a[t ^ stride] : 0ull;
        __syncthreads();
      } else {
        y = __shfl_xor_sync(0xffffffffu, x, stride);
      }
      const bool lower = (t & stride) == 0;
      const bool take_max = lower == desc;
      x = (take_max == (y > x)) ? y : x;
    }
  }
  if (t < n) a[t] = x;
  __syncthreads();
}

#endif  // __CUDACC__

}  // namespace tauv
