// yolact_mask.cu — YOLACT mask assembly on sm_100a.
//
// Replaces (reference file:line under src/tauv_vision/yolact/model/):
//   masks.py:8-21     assemble_mask  (per detection: [P,H,W] broadcast multiply, sum over P, sigmoid, crop)
//   boxes.py:88-103   box_to_mask    (crop predicate, fused into the epilogue)
//
// Two implementations behind one entry point:
//   * mask_umma_kernel (yolact_mask_umma.cuh): the proto x coeff contraction on the 5th-gen tensor cores —
//     tcgen05.mma kind::f16 (bf16 hi/lo operand pairs, fp32 accumulate in TMEM), M = 128 pixels x N <= 256
//     detections per tile, prototype tiles by TMA tensor loads, sigmoid + crop applied on the TMEM -> register
//     read-back, rows written by TMA tensor stores.  Used for P == 32 with 16-byte aligned coefficients.
//   * mask_simt_kernel: exact-fp32 CUDA-core version for every other shape (P != 32, unaligned coefficients) and as
//     the on-device cross-check in the tests.
#include "common.cuh"
#include "yolact_common.cuh"

namespace tauv {

struct MaskArgs {
  const float* proto;        // [B,P,HW]
  const float* coeff;        // single: [n,P]; batched: [B,N,P]
  const int64_t* keep;       // batched: [B,top_k] prior indices; NULL in single mode
  const int32_t* n_keep;     // batched: [B]; NULL -> n_host
  const float4* box;         // [B,top_k,4] / [n,4] or NULL (no crop)
  int n_host;                // detections when n_keep == NULL
  int N;                     // priors per frame (batched)
  int P, H, W;
  int top_k;                 // output rows per frame
  float* out;                // [B,top_k,HW]
  float* logits;             // optional, same shape
  int precise;               // tensor-core epilogue: ex2 + rcp sigmoid (2 ulp) instead of tanh.approx (5e-4 absolute)
  long long* trace;          // debug (tools/mask_trace.py): per-unit role timestamps of CTA 0, or NULL
  // fused consumer (tauv_yolact_mask_depth*): when acc != NULL no mask is written; instead, per detection, the pooled
  // camera depth of the pixels that are "on" (inside the box, sigmoid > 0.5) is added up
  const uint2* pool;         // [B,HW] (sum of mm readings, number of valid readings) per prototype pixel
  unsigned long long* acc;   // [B,top_k,2] (sum, count), zeroed by the caller
};

constexpr int kSimtThreads = 256;
constexpr int kSimtDetChunk = 64;

__global__ void __launch_bounds__(kSimtThreads) mask_simt_kernel(MaskArgs a) {
  extern __shared__ float s_mem[];
  const int P = a.P;
  float* s_proto = s_mem;                              // [P][256]
  float* s_coeff = s_mem + (size_t)P * kSimtThreads;   // [chunk][P]
  float* s_box = s_coeff + (size_t)kSimtDetChunk * P;  // [chunk][4] crop bounds
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int HW = a.H * a.W;
  const int pix = blockIdx.x * kSimtThreads + tid;
  const int n = a.n_keep ? a.n_keep[b] : a.n_host;
  const float* proto = a.proto + (size_t)b * P * HW;
  for (int p = 0; p < P; ++p) s_proto[p * kSimtThreads + tid] = pix < HW ? proto[(size_t)p * HW + pix] : 0.f;
  const float py = (float)(pix / a.W), px = (float)(pix % a.W);
  for (int d0 = 0; d0 < n; d0 += kSimtDetChunk) {
    const int nd = min(kSimtDetChunk, n - d0);
    __syncthreads();
    for (int t = tid; t < nd * P; t += kSimtThreads) {
      const int d = t / P, p = t - d * P;
      const size_t row = a.keep ? ((size_t)b * a.N + (size_t)a.keep[(size_t)b * a.top_k + d0 + d]) : (size_t)(d0 + d);
      s_coeff[t] = a.coeff[row * P + p];
    }
    if (a.box) {
      for (int d = tid; d < nd; d += kSimtThreads) {
        const CropBounds c = crop_bounds(a.box[(size_t)b * a.top_k + d0 + d], a.H, a.W);
        s_box[4 * d] = c.left; s_box[4 * d + 1] = c.right; s_box[4 * d + 2] = c.top; s_box[4 * d + 3] = c.bottom;
      }
    }
    __syncthreads();
    if (a.acc) {
      // fused consumer: every lane takes part in the warp reductions (pixels beyond HW carry an empty pool entry)
      const uint2 pw = pix < HW ? a.pool[(size_t)b * HW + pix] : make_uint2(0u, 0u);
      for (int d = 0; d < nd; ++d) {
        float acc = 0.f;
        for (int p = 0; p < P; ++p) acc = __fadd_rn(acc, __fmul_rn(s_coeff[d * P + p], s_proto[p * kSimtThreads + tid]));
        bool on = sigmoid_ref(acc) > 0.5f;  // the reference's own test on its own fp32 sigmoid (yolact_node.py:178)
        if (a.box) on = on && px >= s_box[4 * d] && px <= s_box[4 * d + 1] && py >= s_box[4 * d + 2] && py <= s_box[4 * d + 3];
        const unsigned s = __reduce_add_sync(0xffffffffu, on ? pw.x : 0u);
        const unsigned c = __reduce_add_sync(0xffffffffu, on ? pw.y : 0u);
        if ((tid & 31) == 0 && (s | c)) {
          unsigned long long* dst = a.acc + ((size_t)b * a.top_k + d0 + d) * 2;
          atomicAdd(dst, (unsigned long long)s);
          atomicAdd(dst + 1, (unsigned long long)c);
        }
      }
    } else if (pix < HW) {
      for (int d = 0; d < nd; ++d) {
        float acc = 0.f;
        for (int p = 0; p < P; ++p) acc = __fadd_rn(acc, __fmul_rn(s_coeff[d * P + p], s_proto[p * kSimtThreads + tid]));
        float v = sigmoid_ref(acc);
        if (a.box) {
          const bool in = px >= s_box[4 * d] && px <= s_box[4 * d + 1] && py >= s_box[4 * d + 2] && py <= s_box[4 * d + 3];
          v = in ? v : 0.f;  // mask *= box_mask  (masks.py:19)
        }
        const size_t o = ((size_t)b * a.top_k + d0 + d) * HW + pix;
        a.out[o] = v;
        if (a.logits) a.logits[o] = acc;
      }
    }
  }
}

}  // namespace tauv

#include "yolact_mask_umma.cuh"

namespace tauv {

// Test hook (only in -DTAUV_DEBUG builds): TAUV_MASK_SIMT in the environment forces the CUDA-core kernel.
static int want_simt_env() { return debug_env("TAUV_MASK_SIMT") ? 1 : 0; }

#ifdef TAUV_DEBUG
static long long* g_mask_trace = nullptr;  // experiment hook (tools/mask_trace.py); only in -DTAUV_DEBUG builds
#endif

static int run_mask(const MaskArgs& a_in, int B, int max_rows, int force_simt, cudaStream_t st) {
  MaskArgs a = a_in;
#ifdef TAUV_DEBUG
  a.trace = g_mask_trace;
#else
  a.trace = nullptr;
#endif
  const int HW = a.H * a.W;
  if (!force_simt && umma_shape_ok(a)) return launch_mask_umma(a, B, max_rows, st);
  const size_t smem = ((size_t)a.P * kSimtThreads + (size_t)kSimtDetChunk * a.P + kSimtDetChunk * 4) * sizeof(float);
  TAUV_REQUIRE(smem <= 227 * 1024, TAUV_E_UNSUPPORTED, "P=%d needs %zu B shared memory", a.P, smem);
  TAUV_CUDA(ensure_dynamic_smem((const void*)(mask_simt_kernel), smem));
  dim3 grid((HW + kSimtThreads - 1) / kSimtThreads, B);
  mask_simt_kernel<<<grid, kSimtThreads, smem, st>>>(a);
  TAUV_LAUNCH_CHECK("mask_simt_kernel");
  return 0;
}

}  // namespace tauv

using namespace tauv;

static int want_simt() { return tauv::want_simt_env(); }

extern "C" int tauv_yolact_assemble_mask(const float* proto, const float* coeff, const float* box, int n, int P, int H,
                                         int W, float* out, float* logits_out, tauv_stream_t stream) {
  TAUV_REQUIRE(n >= 0, TAUV_E_SHAPE, "n must be >= 0");
  if (n == 0) return 0;
  TAUV_REQUIRE(proto && coeff && out, TAUV_E_NULL, "proto/coeff/out must not be NULL");
  TAUV_REQUIRE(P > 0 && H > 0 && W > 0, TAUV_E_SHAPE, "bad shape P=%d H=%d W=%d", P, H, W);
  TAUV_REQUIRE((uintptr_t)box % 16 == 0, TAUV_E_ALIGN, "box must be 16-byte aligned");
  MaskArgs a{};
  a.proto = proto; a.coeff = coeff; a.keep = nullptr; a.n_keep = nullptr; a.box = (const float4*)box;
  a.n_host = n; a.N = 0; a.P = P; a.H = H; a.W = W; a.top_k = n; a.out = out; a.logits = logits_out;
  return run_mask(a, 1, n, want_simt(), (cudaStream_t)stream);
}

extern "C" int tauv_yolact_assemble_mask_batched(const float* proto, const float* coeff_all, const int64_t* keep,
                                                 const int32_t* n_keep, const float* keep_box, int B, int N, int P,
                                                 int H, int W, int top_k, float* out, tauv_stream_t stream) {
  TAUV_REQUIRE(proto && coeff_all && keep && n_keep && out, TAUV_E_NULL, "pointers must not be NULL");
  TAUV_REQUIRE(B > 0 && N > 0 && P > 0 && H > 0 && W > 0 && top_k > 0, TAUV_E_SHAPE, "bad shape");
  TAUV_REQUIRE(B <= 65535, TAUV_E_UNSUPPORTED, "B=%d exceeds the built-in limit 65535", B);
  TAUV_REQUIRE((uintptr_t)keep_box % 16 == 0, TAUV_E_ALIGN, "keep_box must be 16-byte aligned");
  MaskArgs a{};
  a.proto = proto; a.coeff = coeff_all; a.keep = keep; a.n_keep = n_keep; a.box = (const float4*)keep_box;
  a.n_host = 0; a.N = N; a.P = P; a.H = H; a.W = W; a.top_k = top_k; a.out = out; a.logits = nullptr;
  return run_mask(a, B, top_k, want_simt(), (cudaStream_t)stream);
}

// ---- fused consumer: masked depth mean ---------------------------------------------------------
namespace tauv {

// F.interpolate(..., size) in its default 'nearest' mode maps output index i to input index
// min(floor(i * (float)in / out), in - 1), computed in fp32 (ATen UpSample.h: nearest_neighbor_compute_source_index).
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  return min((int)floorf((float)dst * scale), in_size - 1);
}
// first output index whose source index is >= r (the map is monotone): r / scale rounded up, then corrected by the
// map itself, so the result is exact whatever the rounding of the guess
__device__ __forceinline__ int nearest_first_dst(int r, float scale, int in_size, int out_size) {
  int g = min(max((int)ceilf((float)r / scale), 0), out_size);
  while (g > 0 && nearest_src(g - 1, scale, in_size) >= r) --g;
  while (g < out_size && nearest_src(g, scale, in_size) < r) ++g;
  return g;
}

// One thread per prototype pixel: the camera pixels whose nearest prototype pixel it is form a rectangle.
__global__ void __launch_bounds__(256) depth_pool_kernel(const uint16_t* __restrict__ depth, int Hi, int Wi, int H, int W,
                                                         uint2* __restrict__ pool) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= H * W) return;
  const int r = pix / W, c = pix - r * W;
  const float sy = (float)H / (float)Hi, sx = (float)W / (float)Wi;
  const int y0 = nearest_first_dst(r, sy, H, Hi), y1 = r + 1 < H ? nearest_first_dst(r + 1, sy, H, Hi) : Hi;
  const int x0 = nearest_first_dst(c, sx, W, Wi), x1 = c + 1 < W ? nearest_first_dst(c + 1, sx, W, Wi) : Wi;
  const uint16_t* d = depth + (size_t)b * Hi * Wi;
  unsigned sum = 0, cnt = 0;
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      const unsigned v = d[(size_t)y * Wi + x];
      sum += v;             // (a zero reading is "no reading": yolact_node.py:102; it adds nothing to the sum either)
      cnt += v != 0u;
    }
  pool[(size_t)b * H * W + pix] = make_uint2(sum, cnt);
}

// mean = nanmean over the selected camera pixels of depth_mm / 1000 (yolact_node.py:103,178); NaN when none / row unused
__global__ void depth_mean_kernel(const unsigned long long* __restrict__ acc, const int32_t* __restrict__ n_keep, int n_host,
                                  int top_k, int total, double* __restrict__ mean, int64_t* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = i / top_k, d = i - b * top_k;
  const int n = n_keep ? n_keep[b] : n_host;
  const unsigned long long s = acc[2 * (size_t)i], c = acc[2 * (size_t)i + 1];
  const bool used = d < n && c > 0;
  mean[i] = used ? ((double)s / 1000.0) / (double)c : __longlong_as_double(0x7ff8000000000000LL);
  if (count) count[i] = d < n ? (int64_t)c : 0;
}

static int run_mask_depth(MaskArgs a, int B, int max_rows, const uint16_t* depth, int Hi, int Wi, void* workspace,
                          size_t workspace_bytes, double* mean, int64_t* count, cudaStream_t st) {
  const size_t HW = (size_t)a.H * a.W;
  const size_t pool_bytes = (B * HW * sizeof(uint2) + 255) & ~(size_t)255, acc_bytes = (size_t)B * a.top_k * 16;
  TAUV_REQUIRE(workspace && workspace_bytes >= pool_bytes + acc_bytes, TAUV_E_WORKSPACE, "workspace too small: %zu < %zu",
               workspace_bytes, pool_bytes + acc_bytes);
  TAUV_REQUIRE((uintptr_t)workspace % 256 == 0, TAUV_E_ALIGN, "workspace must be 256-byte aligned");
  // a warp adds up 32 pool entries in 32 bits: 32 x (camera pixels per prototype pixel) x 65535 must fit
  const long long pre = (long long)((Hi + a.H - 1) / a.H + 1) * ((Wi + a.W - 1) / a.W + 1);
  TAUV_REQUIRE(pre * 65535LL * 32LL < (1LL << 32), TAUV_E_UNSUPPORTED, "camera image %dx%d too large for a %dx%d prototype map",
               Hi, Wi, a.H, a.W);
  uint2* pool = reinterpret_cast<uint2*>(workspace);
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(workspace) + pool_bytes);
  TAUV_CUDA(cudaMemsetAsync(acc, 0, acc_bytes, st));
  depth_pool_kernel<<<dim3((unsigned)((HW + 255) / 256), B), 256, 0, st>>>(depth, Hi, Wi, a.H, a.W, pool);
  TAUV_LAUNCH_CHECK("depth_pool_kernel");
  a.pool = pool;
  a.acc = acc;
  a.out = nullptr;
  a.logits = nullptr;
  const int rc = run_mask(a, B, max_rows, want_simt_env(), st);
  if (rc) return rc;
  const int total = B * a.top_k;
  depth_mean_kernel<<<(total + 255) / 256, 256, 0, st>>>(acc, a.n_keep, a.n_host, a.top_k, total, mean, count);
  TAUV_LAUNCH_CHECK("depth_mean_kernel");
  return 0;
}

}  // namespace tauv

extern "C" size_t tauv_yolact_mask_depth_workspace_bytes(int B, int H, int W, int top_k) {
  if (B <= 0 || H <= 0 || W <= 0 || top_k <= 0) return 0;
  return (((size_t)B * H * W * sizeof(uint2) + 255) & ~(size_t)255) + (size_t)B * top_k * 16;
}

extern "C" int tauv_yolact_mask_depth(const float* proto, const float* coeff, const float* box, int n, int P, int H, int W,
                                      const uint16_t* depth_mm, int Hi, int Wi, double* mean, int64_t* count,
                                      void* workspace, size_t workspace_bytes, tauv_stream_t stream) {
  TAUV_REQUIRE(n >= 0, TAUV_E_SHAPE, "n must be >= 0");
  if (n == 0) return 0;
  TAUV_REQUIRE(proto && coeff && depth_mm && mean, TAUV_E_NULL, "proto/coeff/depth_mm/mean must not be NULL");
  TAUV_REQUIRE(P > 0 && H > 0 && W > 0 && Hi > 0 && Wi > 0, TAUV_E_SHAPE, "bad shape P=%d H=%d W=%d Hi=%d Wi=%d", P, H, W, Hi, Wi);
  TAUV_REQUIRE((uintptr_t)box % 16 == 0, TAUV_E_ALIGN, "box must be 16-byte aligned");
  MaskArgs a{};
  a.proto = proto; a.coeff = coeff; a.keep = nullptr; a.n_keep = nullptr; a.box = (const float4*)box;
  a.n_host = n; a.N = 0; a.P = P; a.H = H; a.W = W; a.top_k = n;
  return run_mask_depth(a, 1, n, depth_mm, Hi, Wi, workspace, workspace_bytes, mean, count, (cudaStream_t)stream);
}

extern "C" int tauv_yolact_mask_depth_batched(const float* proto, const float* coeff_all, const int64_t* keep,
                                              const int32_t* n_keep, const float* keep_box, int B, int N, int P, int H,
                                              int W, int top_k, const uint16_t* depth_mm, int Hi, int Wi, double* mean,
                                              int64_t* count, void* workspace, size_t workspace_bytes,
                                              tauv_stream_t stream) {
  TAUV_REQUIRE(proto && coeff_all && keep && n_keep && depth_mm && mean, TAUV_E_NULL, "pointers must not be NULL");
  TAUV_REQUIRE(B > 0 && N > 0 && P > 0 && H > 0 && W > 0 && top_k > 0 && Hi > 0 && Wi > 0, TAUV_E_SHAPE, "bad shape");
  TAUV_REQUIRE(B <= 65535, TAUV_E_UNSUPPORTED, "B=%d exceeds the built-in limit 65535", B);
  TAUV_REQUIRE((uintptr_t)keep_box % 16 == 0, TAUV_E_ALIGN, "keep_box must be 16-byte aligned");
  MaskArgs a{};
  a.proto = proto; a.coeff = coeff_all; a.keep = keep; a.n_keep = n_keep; a.box = (const float4*)keep_box;
  a.n_host = 0; a.N = N; a.P = P; a.H = H; a.W = W; a.top_k = top_k;
  return run_mask_depth(a, B, top_k, depth_mm, Hi, Wi, workspace, workspace_bytes, mean, count, (cudaStream_t)stream);
}

// ---- binarised masks at the camera / network-input resolution ----------------------------------
namespace tauv {

// F.interpolate(..., mode="bilinear") with align_corners=False (ATen UpSample.h: area_pixel_compute_source_index):
// source coordinate = max(fp32(in/out) * (dst + 0.5) - 0.5, 0); the left/top tap is its integer part, the other tap one
// further unless that leaves the image; weights (1 - frac, frac).  The multiply-subtract is a single fused operation
// in ATen's builds (CPU and CUDA alike — it shows in the last bits of the weights, see oracle/ref_port.py), hence the
// explicit fmaf: this library is otherwise built without FMA contraction.
struct LinearTap {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ LinearTap linear_tap(int dst, float scale, int in_size) {
  float s = fmaf(scale, (float)dst + 0.5f, -0.5f);
  s = s < 0.0f ? 0.0f : s;
  LinearTap t;
  t.i0 = min((int)s, in_size - 1);
  t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
  t.w1 = s - (float)t.i0;
  t.w0 = 1.0f - t.w1;
  return t;
}

// low [B*top_k][H*W] fp32 masks -> out [B*top_k][Ho*Wo] bytes (1 where the resized mask is > 0.5).  One thread makes
// VEC horizontally adjacent bytes (VEC == 4 needs Wo % 4 == 0 and a 4-byte aligned output: one 128-byte store per warp).
// MODE 0: nearest (yolact_node.py:135, followed by the node's `mask_np > 0.5` at :178); 1: bilinear
// (evaluate_batch.py:101-102).  Rows >= n_keep[b] are not written.
template <int MODE, int VEC>
__global__ void __launch_bounds__(256) mask_binary_resize_kernel(const float* __restrict__ low, const float4* __restrict__ box,
                                                                 const int32_t* __restrict__ n_keep, int n_host, int top_k,
                                                                 int H, int W, int Ho, int Wo, uint8_t* __restrict__ out) {
  const int j = blockIdx.y, b = blockIdx.z;
  if (j >= (n_keep ? n_keep[b] : n_host)) return;
  const size_t row = (size_t)b * top_k + j;
  const float* __restrict__ src = low + row * (size_t)H * W;
  const long long p0 = ((long long)blockIdx.x * 256 + threadIdx.x) * VEC;
  if (p0 >= (long long)Ho * Wo) return;
  const int yo = (int)(p0 / Wo), xo = (int)(p0 - (long long)yo * Wo);
  const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
  unsigned bits[VEC >= 4 ? VEC / 4 : 1] = {};
  bool zero = false;
  if (box) {  // pixels whose source taps lie clear of the crop box (two-pixel margin) are zeros by construction
    const CropBounds cb = crop_bounds(box[row], H, W);
    const float ys = ((float)yo + 0.5f) * sy, x0s = ((float)xo + 0.5f) * sx, x1s = ((float)(xo + VEC - 1) + 0.5f) * sx;
    zero = ys < cb.top - 2.0f || ys > cb.bottom + 3.0f || x1s < cb.left - 2.0f || x0s > cb.right + 3.0f;
  }
  if (zero) {
  } else if (MODE == 0) {
    const float* r = src + (size_t)nearest_src(yo, sy, H) * W;
#pragma unroll
    for (int v = 0; v < VEC; ++v) bits[v >> 2] |= (r[nearest_src(xo + v, sx, W)] > 0.5f ? 1u : 0u) << (8 * (v & 3));
  } else {
    const LinearTap ty = linear_tap(yo, sy, H);
    const float* r0 = src + (size_t)ty.i0 * W;
    const float* r1 = src + (size_t)ty.i1 * W;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const LinearTap tx = linear_tap(xo + v, sx, W);
      const float val = ty.w0 * (tx.w0 * r0[tx.i0] + tx.w1 * r0[tx.i1]) + ty.w1 * (tx.w0 * r1[tx.i0] + tx.w1 * r1[tx.i1]);
      bits[v >> 2] |= (val > 0.5f ? 1u : 0u) << (8 * (v & 3));
    }
  }
  uint8_t* dst = out + row * (size_t)Ho * Wo + p0;
  if constexpr (VEC == 16) *reinterpret_cast<uint4*>(dst) = make_uint4(bits[0], bits[1], bits[2], bits[3]);
  else if constexpr (VEC == 4) *reinterpret_cast<uint32_t*>(dst) = bits[0];
  else *dst = (uint8_t)bits[0];
}

// Nearest mode, rows that are multiples of 16 bytes: one thread per (SOURCE row, 16 output pixels).  Nearest resizing
// repeats a source row in every output row that maps to it (2.6 of them at 276 -> 720), so the 16 bytes are made once
// and stored that many times; and a source row (or a column chunk) that lies clear of the detection's crop box is
// zeros by construction (the assembly has applied the same box), so it is stored without being read.
__global__ void __launch_bounds__(256) mask_binary_nearest16_kernel(const float* __restrict__ low, const float4* __restrict__ box,
                                                                    const int32_t* __restrict__ n_keep, int n_host, int top_k,
                                                                    int H, int W, int Ho, int Wo, uint8_t* __restrict__ out) {
  const int j = blockIdx.y, b = blockIdx.z;
  if (j >= (n_keep ? n_keep[b] : n_host)) return;
  const size_t row = (size_t)b * top_k + j;
  const int chunks = Wo >> 4;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= H * chunks) return;
  const int r = idx / chunks, xo = (idx - r * chunks) << 4;
  const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
  const int y0 = nearest_first_dst(r, sy, H, Ho), y1 = r + 1 < H ? nearest_first_dst(r + 1, sy, H, Ho) : Ho;
  if (y0 >= y1) return;
  unsigned bits[4] = {0u, 0u, 0u, 0u};
  bool zero = false;
  if (box) {  // (a one-pixel margin keeps this test independent of the rounding of the crop itself)
    const CropBounds cb = crop_bounds(box[row], H, W);
    const float x_lo = (float)nearest_src(xo, sx, W), x_hi = (float)nearest_src(xo + 15, sx, W);
    zero = (float)r < cb.top - 1.0f || (float)r > cb.bottom + 1.0f || x_hi < cb.left - 1.0f || x_lo > cb.right + 1.0f;
  }
  if (!zero) {
    const float* src = low + (row * (size_t)H + r) * W;
#pragma unroll
    for (int v = 0; v < 16; ++v) bits[v >> 2] |= (src[nearest_src(xo + v, sx, W)] > 0.5f ? 1u : 0u) << (8 * (v & 3));
  }
  const uint4 q = make_uint4(bits[0], bits[1], bits[2], bits[3]);
  uint8_t* dst = out + row * (size_t)Ho * Wo + xo;
  for (int yo = y0; yo < y1; ++yo) *reinterpret_cast<uint4*>(dst + (size_t)yo * Wo) = q;
}

static size_t mask_binary_ws_bytes(int B, int H, int W, int top_k) {
  return ((size_t)B * top_k * H * W * sizeof(float) + 255) & ~(size_t)255;
}

static int run_mask_binary(MaskArgs a, int B, int max_rows, int Ho, int Wo, int mode, uint8_t* out, void* workspace,
                           size_t workspace_bytes, cudaStream_t st) {
  TAUV_REQUIRE(mode == TAUV_RESIZE_NEAREST || mode == TAUV_RESIZE_BILINEAR, TAUV_E_UNSUPPORTED,
               "mode must be TAUV_RESIZE_NEAREST or TAUV_RESIZE_BILINEAR; got %d", mode);
  const size_t need = mask_binary_ws_bytes(B, a.H, a.W, a.top_k);
  TAUV_REQUIRE(workspace && workspace_bytes >= need, TAUV_E_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
  TAUV_REQUIRE((uintptr_t)workspace % 256 == 0, TAUV_E_ALIGN, "workspace must be 256-byte aligned");
  TAUV_REQUIRE(a.top_k <= 65535, TAUV_E_UNSUPPORTED, "%d masks per frame exceed the built-in limit 65535", a.top_k);
  a.out = reinterpret_cast<float*>(workspace);
  a.logits = nullptr;
  a.precise = mode == TAUV_RESIZE_BILINEAR;
  const int rc = run_mask(a, B, max_rows, want_simt_env(), st);
  if (rc) return rc;
  const long long total = (long long)Ho * Wo;
  const bool vec = Wo % 4 == 0 && (uintptr_t)out % 4 == 0;
  // 16 bytes per thread (one 128-bit store, 512 bytes per warp instruction) when the rows allow it: the pass writes one
  // byte per camera pixel and kept mask — 9.4 GB per 64 frames — and was bound by its 4-byte stores and per-pixel setup
  const bool vec16 = Wo % 16 == 0 && (uintptr_t)out % 16 == 0;
  const long long per_block = 256LL * (vec16 ? 16 : vec ? 4 : 1);
  const dim3 grid((unsigned)((total + per_block - 1) / per_block), (unsigned)a.top_k, (unsigned)B);
#define TAUV_RESIZE(MODE, VEC) \
  mask_binary_resize_kernel<MODE, VEC><<<grid, 256, 0, st>>>(a.out, reinterpret_cast<const float4*>(a.box), a.n_keep, \
                                                             a.n_host, a.top_k, a.H, a.W, Ho, Wo, out)
  if (mode == TAUV_RESIZE_NEAREST) {
    if (vec16) {
      const dim3 g16((unsigned)(((long long)a.H * (Wo >> 4) + 255) / 256), (unsigned)a.top_k, (unsigned)B);
      mask_binary_nearest16_kernel<<<g16, 256, 0, st>>>(a.out, reinterpret_cast<const float4*>(a.box), a.n_keep, a.n_host,
                                                        a.top_k, a.H, a.W, Ho, Wo, out);
    } else if (vec) TAUV_RESIZE(0, 4); else TAUV_RESIZE(0, 1);
  } else {
    if (vec16) TAUV_RESIZE(1, 16); else if (vec) TAUV_RESIZE(1, 4); else TAUV_RESIZE(1, 1);
  }
#undef TAUV_RESIZE
  TAUV_LAUNCH_CHECK("mask_binary_resize_kernel");
  return 0;
}

}  // namespace tauv

extern "C" size_t tauv_yolact_mask_binary_workspace_bytes(int B, int H, int W, int top_k) {
  if (B <= 0 || H <= 0 || W <= 0 || top_k <= 0) return 0;
  return tauv::mask_binary_ws_bytes(B, H, W, top_k);
}

extern "C" int tauv_yolact_mask_binary(const float* proto, const float* coeff, const float* box, int n, int P, int H, int W,
                                       int out_h, int out_w, int mode, uint8_t* out, void* workspace,
                                       size_t workspace_bytes, tauv_stream_t stream) {
  TAUV_REQUIRE(n >= 0, TAUV_E_SHAPE, "n must be >= 0");
  if (n == 0) return 0;
  TAUV_REQUIRE(proto && coeff && out, TAUV_E_NULL, "proto/coeff/out must not be NULL");
  TAUV_REQUIRE(P > 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0, TAUV_E_SHAPE, "bad shape P=%d H=%d W=%d out=%dx%d", P, H, W,
               out_h, out_w);
  TAUV_REQUIRE((uintptr_t)box % 16 == 0, TAUV_E_ALIGN, "box must be 16-byte aligned");
  MaskArgs a{};
  a.proto = proto; a.coeff = coeff; a.keep = nullptr; a.n_keep = nullptr; a.box = (const float4*)box;
  a.n_host = n; a.N = 0; a.P = P; a.H = H; a.W = W; a.top_k = n;
  return run_mask_binary(a, 1, n, out_h, out_w, mode, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int tauv_yolact_mask_binary_batched(const float* proto, const float* coeff_all, const int64_t* keep,
                                               const int32_t* n_keep, const float* keep_box, int B, int N, int P, int H,
                                               int W, int top_k, int out_h, int out_w, int mode, uint8_t* out,
                                               void* workspace, size_t workspace_bytes, tauv_stream_t stream) {
  TAUV_REQUIRE(proto && coeff_all && keep && n_keep && out, TAUV_E_NULL, "pointers must not be NULL");
  TAUV_REQUIRE(B > 0 && N > 0 && P > 0 && H > 0 && W > 0 && top_k > 0 && out_h > 0 && out_w > 0, TAUV_E_SHAPE, "bad shape");
  TAUV_REQUIRE(B <= 65535, TAUV_E_UNSUPPORTED, "B=%d exceeds the built-in limit 65535", B);
  TAUV_REQUIRE((uintptr_t)keep_box % 16 == 0, TAUV_E_ALIGN, "keep_box must be 16-byte aligned");
  MaskArgs a{};
  a.proto = proto; a.coeff = coeff_all; a.keep = keep; a.n_keep = n_keep; a.box = (const float4*)keep_box;
  a.n_host = 0; a.N = N; a.P = P; a.H = H; a.W = W; a.top_k = top_k;
  return run_mask_binary(a, B, top_k, out_h, out_w, mode, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

#ifdef TAUV_DEBUG
// Debug hook for tools/mask_trace.py (only in -DTAUV_DEBUG builds; process-global, not thread-safe).
extern "C" void tauv_debug_mask_trace(long long* buf) { tauv::g_mask_trace = buf; }
#endif
