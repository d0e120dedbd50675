// elementwise.cu — the small elementwise decoders of the CenterNet head (scope row a6).
//   decode.py:291-316  angle_decode   (two-bin softmax select + atan2 offset)
//   decode.py:319-324  depth_decode   (1/sigmoid(d) - 1)
#include "common.cuh"
#include <math.h>

namespace tauv {

// softmax over a pair, ATen order: subtract the max, exp, sum, divide; returns p[1]
__device__ __forceinline__ float softmax2_inside(float a, float b) {
  const float m = fmaxf(a, b);
  const float ea = expf(__fsub_rn(a, m)), eb = expf(__fsub_rn(b, m));
  return __fdiv_rn(eb, __fadd_rn(ea, eb));
}

// torch.remainder(x, b) for b > 0 (Python-style modulo on fmod)
__device__ __forceinline__ float py_mod(float x, float b) {
  float m = fmodf(x, b);
  if (m != 0.0f && m < 0.0f) m = __fadd_rn(m, b);
  return m;
}

__global__ void angle_decode_kernel(const float4* __restrict__ bin, const float4* __restrict__ off, long long n,
                                    float c0, float c1, float two_pi, float scale, float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 b = bin[i], o = off[i];
  const float s0 = softmax2_inside(b.x, b.y);
  const float s1 = softmax2_inside(b.z, b.w);
  const float a0 = __fadd_rn(atan2f(o.x, o.y), c0);
  const float a1 = __fadd_rn(atan2f(o.z, o.w), c1);
  float a = (s1 > s0) ? a1 : a0;
  a = py_mod(a, two_pi);
  out[i] = __fmul_rn(a, scale);
}

__global__ void depth_decode_kernel(const float* __restrict__ in, long long n, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __fsub_rn(__fdiv_rn(1.0f, sigmoid_ref(in[i])), 1.0f);
}

}  // namespace tauv

using namespace tauv;

extern "C" int tauv_angle_decode(const float* predicted_bin, const float* predicted_offset, int64_t n,
                                 double theta_range, float* out, tauv_stream_t stream) {
  TAUV_REQUIRE(n >= 0, TAUV_E_SHAPE, "n must be >= 0");
  if (n == 0) return 0;
  TAUV_REQUIRE(predicted_bin && predicted_offset && out, TAUV_E_NULL, "pointers must not be NULL");
  TAUV_REQUIRE((uintptr_t)predicted_bin % 16 == 0 && (uintptr_t)predicted_offset % 16 == 0, TAUV_E_ALIGN,
               "bin/offset rows must be 16-byte aligned");
  angle_decode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(predicted_bin), reinterpret_cast<const float4*>(predicted_offset), n,
      (float)(M_PI / 2), (float)(-M_PI / 2), (float)(2 * M_PI), (float)(theta_range / (2 * M_PI)), out);
  TAUV_LAUNCH_CHECK("angle_decode_kernel");
  return 0;
}

extern "C" int tauv_depth_decode(const float* in, int64_t n, float* out, tauv_stream_t stream) {
  TAUV_REQUIRE(n >= 0, TAUV_E_SHAPE, "n must be >= 0");
  if (n == 0) return 0;
  TAUV_REQUIRE(in && out, TAUV_E_NULL, "pointers must not be NULL");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  depth_decode_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, n, out);
  TAUV_LAUNCH_CHECK("depth_decode_kernel");
  return 0;
}
