// yolact_mask.cu — YOLACT mask assembly on sm_100a.
//
// Replaces (reference file:line under src/tauv_vision/yolact/model/):
//   masks.py:8-21     assemble_mask  (per detection: [P,H,W] broadcast multiply, sum over P, sigmoid, crop)
//   boxes.py:88-103   box_to_mask    (crop predicate, fused into the epilogue)
//
// Two implementations behind one entry point:
//   * mask_umma_kernel (yolact_mask_umma.cuh): the proto x coeff contraction on the 5th-gen tensor cores —
//     tcgen05.mma kind::f16 (bf16 operands, fp32 accumulate in TMEM), M = 128 detections x N = 256 pixels
//     per tile, sigmoid + crop applied on the TMEM -> register read-back.  Used whenever the shape fits.
//   * mask_simt_kernel: exact-fp32 CUDA-core version for shapes the tensor-core tiling does not take
//     (P not a multiple of 16 or > 64, odd H*W) and as the on-device cross-check in the tests.
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
  long long* trace;          // debug (tools/mask_trace.py): per-unit role timestamps of CTA 0, or NULL
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
    if (pix < HW) {
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

static long long* g_mask_trace = nullptr;  // experiment hook (tools/mask_trace.py); not part of the public ABI

static int run_mask(const MaskArgs& a_in, int B, int max_rows, int force_simt, cudaStream_t st) {
  MaskArgs a = a_in;
  a.trace = g_mask_trace;
  const int HW = a.H * a.W;
  if (!force_simt && umma_shape_ok(a)) return launch_mask_umma(a, B, max_rows, st);
  const size_t smem = ((size_t)a.P * kSimtThreads + (size_t)kSimtDetChunk * a.P + kSimtDetChunk * 4) * sizeof(float);
  TAUV_REQUIRE(smem <= 227 * 1024, TAUV_E_UNSUPPORTED, "P=%d needs %zu B shared memory", a.P, smem);
  TAUV_CUDA(cudaFuncSetAttribute(mask_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((HW + kSimtThreads - 1) / kSimtThreads, B);
  mask_simt_kernel<<<grid, kSimtThreads, smem, st>>>(a);
  TAUV_LAUNCH_CHECK("mask_simt_kernel");
  return 0;
}

}  // namespace tauv

using namespace tauv;

// Test hook: TAUV_MASK_SIMT=1 in the environment forces the CUDA-core kernel (read per call; no global state).
static int want_simt() {
  const char* e = getenv("TAUV_MASK_SIMT");
  return e && e[0] == '1';
}

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

// Debug hook for tools/mask_trace.py (process-global, not thread-safe, not in the public header).
extern "C" void tauv_debug_mask_trace(long long* buf) { tauv::g_mask_trace = buf; }
