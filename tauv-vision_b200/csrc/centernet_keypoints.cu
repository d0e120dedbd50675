// centernet_keypoints.cu — greedy keypoint -> object association of decode_keypoints on the device.
//
// Replaces the host Python double loop of the reference (src/tauv_vision/centernet/model/decode.py:98-135):
//   for every detected keypoint in rank order (stop at the first score below the threshold, :100-102 — a DOUBLE
//   compare: `float(score) < keypoint_score_threshold`), look up (object label, keypoint slot) of its channel
//   (:104-106), take the detections of that label whose slot is still free (:108-111), and give the keypoint to the one
//   whose direction from the object centre best matches the keypoint's affinity vector: min over candidates of
//   |atan2(a_y, a_x) - atan2(k_y - d_y, k_x - d_x)| in Python doubles, first minimum wins (:113-131).
// One warp per frame: the keypoints are inherently sequential (a slot taken by keypoint j is closed for j+1), the
// candidates of one keypoint are tested by the lanes in parallel.  The affinity vector is read straight from the
// strided [B,Kp,2,H,W] head tensor; nothing but the <= k x max_kp assigned keypoints is written.
#include "common.cuh"

namespace tauv {

constexpr int kKpMaxSlots = 4096;  // k * max_kp flags in shared memory

__global__ void __launch_bounds__(32) keypoint_assoc_kernel(
    const int64_t* __restrict__ label, const double* __restrict__ yx, const int32_t* __restrict__ count, int k,
    const int64_t* __restrict__ kp_index, const int64_t* __restrict__ kp_label, const float* __restrict__ kp_score, int kk,
    const float* __restrict__ affinity, long long ab, long long ak, long long ac, long long ay_, long long ax_,
    const int32_t* __restrict__ kp_map, int Kp, int max_kp, int out_h, int out_w, double kp_thr,
    uint8_t* __restrict__ kp_set, float* __restrict__ kp_yx, float* __restrict__ kp_score_out, float* __restrict__ kp_aff_out) {
  __shared__ uint8_t taken[kKpMaxSlots];
  const int b = blockIdx.x, lane = threadIdx.x;
  const int n = count[b];
  for (int i = lane; i < k * max_kp; i += 32) {
    taken[i] = 0;
    kp_set[(size_t)b * k * max_kp + i] = 0;
  }
  __syncwarp();
  for (int j = 0; j < kk; ++j) {
    const float s = kp_score[(size_t)b * kk + j];
    if ((double)s < kp_thr) break;  // decode.py:100-102
    const long long kl = kp_label[(size_t)b * kk + j];
    if (kl < 0 || kl >= Kp) continue;
    const int obj = kp_map[2 * kl], slot = kp_map[2 * kl + 1];
    if (slot < 0 || slot >= max_kp) continue;
    const long long iy = kp_index[((size_t)b * kk + j) * 2], ix = kp_index[((size_t)b * kk + j) * 2 + 1];
    // int64 tensor / int -> fp32 true divide, then float() (decode.py:115-118)
    const float kyf = __fdiv_rn((float)iy, (float)out_h), kxf = __fdiv_rn((float)ix, (float)out_w);
    const float ay = affinity[b * ab + kl * ak + iy * ay_ + ix * ax_];
    const float ax = affinity[b * ab + kl * ak + ac + iy * ay_ + ix * ax_];
    const double ang = atan2((double)ay, (double)ax);
    double best = 0.0;
    int best_d = -1;
    for (int d0 = 0; d0 < n; d0 += 32) {
      const int d = d0 + lane;
      double err = 0.0;
      bool cand = false;
      if (d < n && label[(size_t)b * k + d] == obj && !taken[d * max_kp + slot]) {
        cand = true;
        err = fabs(ang - atan2((double)kyf - yx[((size_t)b * k + d) * 2], (double)kxf - yx[((size_t)b * k + d) * 2 + 1]));
      }
      // lexicographic (err, d) minimum over the lanes, merged with the best of the earlier rounds (their d is lower)
      int md = cand ? d : 0x7fffffff;
      double me = err;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double oe = __shfl_xor_sync(0xffffffffu, me, o);
        const int od = __shfl_xor_sync(0xffffffffu, md, o);
        const bool take = od != 0x7fffffff && (md == 0x7fffffff || oe < me || (oe == me && od < md));
        if (take) {
          me = oe;
          md = od;
        }
      }
      if (md != 0x7fffffff && (best_d < 0 || me < best)) {  // (strictly smaller: an earlier detection keeps a tie)
        best = me;
        best_d = md;
      }
    }
    if (best_d >= 0) {
      if (lane == 0) {
        const size_t o = ((size_t)b * k + best_d) * max_kp + slot;
        taken[best_d * max_kp + slot] = 1;
        kp_set[o] = 1;
        kp_yx[o * 2] = kyf;
        kp_yx[o * 2 + 1] = kxf;
        kp_score_out[o] = s;
        kp_aff_out[o * 2] = ay;
        kp_aff_out[o * 2 + 1] = ax;
      }
      __syncwarp();
    }
  }
}

}  // namespace tauv

using namespace tauv;

extern "C" int tauv_centernet_keypoint_assoc(const int64_t* label, const double* yx, const int32_t* count, int B, int k,
                                             const int64_t* kp_index, const int64_t* kp_label, const float* kp_score,
                                             int kk, const float* affinity, const int64_t affinity_strides[5],
                                             const int32_t* kp_map, int Kp, int max_kp, int out_h, int out_w,
                                             double keypoint_score_threshold, uint8_t* kp_set, float* kp_yx,
                                             float* kp_score_out, float* kp_aff_out, tauv_stream_t stream) {
  TAUV_REQUIRE(label && yx && count && kp_index && kp_label && kp_score && affinity && affinity_strides && kp_map,
               TAUV_E_NULL, "inputs must not be NULL");
  TAUV_REQUIRE(kp_set && kp_yx && kp_score_out && kp_aff_out, TAUV_E_NULL, "outputs must not be NULL");
  TAUV_REQUIRE(B > 0 && k > 0 && kk > 0 && Kp > 0 && max_kp > 0 && out_h > 0 && out_w > 0, TAUV_E_SHAPE, "bad shape");
  TAUV_REQUIRE((long long)k * max_kp <= kKpMaxSlots, TAUV_E_UNSUPPORTED, "k * max_kp = %lld exceeds the built-in limit %d",
               (long long)k * max_kp, kKpMaxSlots);
  keypoint_assoc_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(
      label, yx, count, k, kp_index, kp_label, kp_score, kk, affinity, affinity_strides[0], affinity_strides[1],
      affinity_strides[2], affinity_strides[3], affinity_strides[4], kp_map, Kp, max_kp, out_h, out_w,
      keypoint_score_threshold, kp_set, kp_yx, kp_score_out, kp_aff_out);
  TAUV_LAUNCH_CHECK("keypoint_assoc_kernel");
  return 0;
}
