// yolact_boxes.cu — YOLACT box arithmetic on sm_100a.
//
// Replaces (reference file:line under src/tauv_vision/yolact/model/):
//   anchors.py:9-41    get_anchor            (CPU + H2D on every forward today, model.py:47-48)
//   boxes.py:45-52     box_encode
//   boxes.py:55-61     box_decode
//   boxes.py:64-85     iou_matrix            (~14 materialised temporaries of the output size)
//   boxes.py:88-103    box_to_mask
//   loss.py:16-22,62-66  anchor matching + regression targets ([B,N,M] temporaries)
//
// Everything here is elementwise fp32 with the reference's exact operation order; FMA contraction is
// off for the whole library (--fmad=false) and the divides / exp / log are IEEE or libdevice.
#include "common.cuh"
#include "yolact_common.cuh"

namespace tauv {

template <bool ENCODE>
__global__ void __launch_bounds__(256) box_codec_kernel(const float4* __restrict__ in, const float4* __restrict__ anchor,
                                                        long long total, int N, int anchor_batch, float v0, float v1,
                                                        float4* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 a = anchor[anchor_batch == 1 ? (i % N) : i];
    const float4 x = in[i];
    out[i] = ENCODE ? encode_one(x, a, v0, v1) : decode_one(x, a, v0, v1);
  }
}

__global__ void __launch_bounds__(256) iou_matrix_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                         int Ba, int Bb, int Na, int Nb, long long total,
                                                         float* __restrict__ out) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t % Nb);
    const long long r = t / Nb;
    const int i = (int)(r % Na);
    const long long bb = r / Na;
    const Corners ca = to_corners(a[(Ba == 1 ? 0 : bb) * Na + i]);
    const Corners cb = to_corners(b[(Bb == 1 ? 0 : bb) * Nb + j]);
    out[t] = iou_pair(ca, cb);
  }
}

struct AnchorHW {
  float h[8], w[8];
};

__global__ void anchors_level_kernel(int H, int W, int A, AnchorHW hw, float4* __restrict__ out) {
  const int n = A * H * W;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int a = t / (H * W);          // aspect-major inside the level (anchors.py:19-20,36-39)
  const int cell = t - a * (H * W);
  const int i = cell / W, j = cell - i * W;
  float4 v;
  v.x = __fdiv_rn(__fadd_rn((float)i, 0.5f), (float)H);  // anchors.py:14
  v.y = __fdiv_rn(__fadd_rn((float)j, 0.5f), (float)W);  // anchors.py:15
  v.z = hw.h[a];
  v.w = hw.w[a];
  out[t] = v;
}

__global__ void box_to_mask_kernel(const float* __restrict__ box, int H, int W, float* __restrict__ out) {
  const CropBounds c = crop_bounds(make_float4(box[0], box[1], box[2], box[3]), H, W);
  const int n = H * W;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const float y = (float)(t / W), x = (float)(t % W);
    out[t] = (x >= c.left && x <= c.right && y >= c.top && y <= c.bottom) ? 1.0f : 0.0f;
  }
}

// loss.py:16-22 + :62-66.  One thread per (frame, prior); the frame's M truths sit in shared memory.
constexpr int kMatchTiles = 1;  // (measured: 4 tiles per CTA — truths staged once per 1024 priors — is slower, 34.8 -> 40.9 us: the kernel wants CTAs in flight, not fewer prologues)
__global__ void __launch_bounds__(256) match_anchors_kernel(
    const float4* __restrict__ anchor, const float4* __restrict__ truth_box, const uint8_t* __restrict__ truth_valid,
    int N, int M, float pos_thr, float neg_thr, float v0, float v1, int64_t* __restrict__ match_index,
    float* __restrict__ match_iou, uint8_t* __restrict__ positive, uint8_t* __restrict__ negative,
    float4* __restrict__ target) {
  extern __shared__ float s_truth[];  // [M][8]: y0,x0,y1,x1,area,valid, "all five finite and not NaN" flag, pad
  const int b = blockIdx.y;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const Corners c = to_corners(truth_box[(size_t)b * M + m]);
    float* s = s_truth + 8 * m;
    s[0] = c.y0; s[1] = c.x0; s[2] = c.y1; s[3] = c.x1; s[4] = c.area;
    s[5] = truth_valid[(size_t)b * M + m] ? 1.0f : 0.0f;
    const float sum = c.y0 + c.x0 + c.y1 + c.x1 + c.area;  // NaN or inf in any of them poisons the sum
    s[6] = (fabsf(sum) <= 3.0e38f) ? 1.0f : 0.0f;
    s[7] = 0.0f;
  }
  __syncthreads();
  // (every CTA takes kMatchTiles tiles of 256 priors of its frame: the truths are staged once per CTA, and 64 frames x
  // 19 CTAs are one wave on 148 SMs)
#pragma unroll 1
  for (int tile = 0; tile < kMatchTiles; ++tile) {
  const int n = (blockIdx.x * kMatchTiles + tile) * blockDim.x + threadIdx.x;
  if ((n & ~31) >= N) break;    // (warp-uniform: the whole warp is past the end)
  const bool in_range = n < N;  // (no early return for single lanes: the warp votes below)
  const float4 av = anchor[in_range ? n : N - 1];
  const Corners ca = to_corners(av);
  const bool a_ok = fabsf(ca.y0 + ca.x0 + ca.y1 + ca.x1 + ca.area) <= 3.0e38f;
  // ---- cull per warp before computing.  A warp holds 32 consecutive priors — neighbours on one row of one FPN level —
  // so their hull is small and most truths miss it altogether.  Lane m tests truth m against the hull: if the hull's
  // intersection height or width with the truth is <= 0 then so is every lane's (rounded subtraction is monotone in
  // both operands), the unions of all lanes are positive and finite (checked with the smallest and largest prior
  // area), and every (prior, truth) pair of the warp is the exact zero the per-pair fast path would produce.  Only the
  // truths that survive — typically one or two of 16 — are looped over; a truth or a prior with a NaN / infinite
  // coordinate is never culled.
  const int lane = threadIdx.x & 31;
  float hy0 = ca.y0, hx0 = ca.x0, hy1 = ca.y1, hx1 = ca.x1, amin = ca.area, amax = ca.area;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    hy0 = fminf(hy0, __shfl_xor_sync(0xffffffffu, hy0, o));
    hx0 = fminf(hx0, __shfl_xor_sync(0xffffffffu, hx0, o));
    hy1 = fmaxf(hy1, __shfl_xor_sync(0xffffffffu, hy1, o));
    hx1 = fmaxf(hx1, __shfl_xor_sync(0xffffffffu, hx1, o));
    amin = fminf(amin, __shfl_xor_sync(0xffffffffu, amin, o));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  }
  const bool warp_ok = __all_sync(0xffffffffu, a_ok);
  float best = 0.f;
  int best_m = 0;
  for (int m0 = 0; m0 < M; m0 += 32) {
    bool cand = false;
    if (m0 + lane < M) {
      const float4 s0 = *reinterpret_cast<const float4*>(s_truth + 8 * (m0 + lane));      // y0 x0 y1 x1
      const float4 s1 = *reinterpret_cast<const float4*>(s_truth + 8 * (m0 + lane) + 4);  // area valid ok -
      const float ih = __fsub_rn(fminf(hy1, s0.z), fmaxf(hy0, s0.x));
      const float iw = __fsub_rn(fminf(hx1, s0.w), fmaxf(hx0, s0.y));
      const float ulo = __fadd_rn(amin, s1.x), uhi = __fadd_rn(amax, s1.x);
      const bool culled = warp_ok && s1.z != 0.0f && (ih <= 0.0f || iw <= 0.0f) && ulo > 0.0f && uhi <= 3.0e38f;
      cand = !culled;
    }
    unsigned todo = __ballot_sync(0xffffffffu, cand);
    while (todo) {
      const int m = m0 + __ffs(todo) - 1;
      todo &= todo - 1;
      const float4 s0 = *reinterpret_cast<const float4*>(s_truth + 8 * m);      // y0 x0 y1 x1
      const float4 s1 = *reinterpret_cast<const float4*>(s_truth + 8 * m + 4);  // area valid ok -
      float v;
      // Most remaining (prior, truth) pairs do not overlap either: the intersection is exactly 0 and, with a positive
      // finite union, so is the IoU (0/u = +0, times valid = +0) — no NaN-propagating min/max, no IEEE divide.
      // Everything else takes the reference arithmetic.
      const float ih = __fsub_rn(fminf(ca.y1, s0.z), fmaxf(ca.y0, s0.x));
      const float iw = __fsub_rn(fminf(ca.x1, s0.w), fmaxf(ca.x0, s0.y));
      const float uni0 = __fadd_rn(ca.area, s1.x);
      const bool finite = a_ok && s1.z != 0.0f;
      if (finite && (ih <= 0.0f || iw <= 0.0f) && uni0 > 0.0f && uni0 <= 3.0e38f) {
        v = 0.0f;
      } else if (finite) {
        // no NaN anywhere: fminf / fmaxf ARE torch.min / torch.max, and ih, iw, uni0 above are iou_pair's own terms
        // (boxes.py:68-83) — the NaN-propagating version below is 40 instructions that every lane of the warp would
        // step through for the sake of the few that overlap this truth
        const float inter = __fmul_rn(fmaxf(ih, 0.0f), fmaxf(iw, 0.0f));
        v = __fmul_rn(__fdiv_rn(inter, __fsub_rn(uni0, inter)), s1.y);  // iou * truth_valid.float()
      } else {
        Corners cb;
        cb.y0 = s0.x; cb.x0 = s0.y; cb.y1 = s0.z; cb.x1 = s0.w; cb.area = s1.x;
        v = __fmul_rn(iou_pair(ca, cb), s1.y);
      }
      // torch.max(dim): first maximum wins; a NaN wins over everything that came before it.  (A culled truth is an
      // exact +0: it can only ever be the maximum as truth 0, which is what best / best_m start from.)
      if (m == 0 || v > best || (v != v && best == best)) {
        best = v;
        best_m = m;
      }
    }
  }
  if (!in_range) continue;
  const size_t o = (size_t)b * N + n;
  match_index[o] = best_m;
  match_iou[o] = best;
  positive[o] = best >= pos_thr;
  negative[o] = best <= neg_thr;
  // the reference encodes the positives only (loss.py:62-66): everything else is written as zeros
  if (target)
    target[o] = (best >= pos_thr) ? encode_one(truth_box[(size_t)b * M + best_m], av, v0, v1) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

static unsigned grid_for(long long total, int threads) {
  long long blocks = (total + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace tauv

using namespace tauv;

static int codec(bool encode, const float* in, const float* anchor, int B, int N, int anchor_batch, float v0, float v1,
                 float* out, cudaStream_t st) {
  TAUV_REQUIRE(in && anchor && out, TAUV_E_NULL, "in/anchor/out must not be NULL");
  TAUV_REQUIRE(B > 0 && N > 0, TAUV_E_SHAPE, "bad shape B=%d N=%d", B, N);
  TAUV_REQUIRE(anchor_batch == 1 || anchor_batch == B, TAUV_E_SHAPE, "anchor batch %d must be 1 or %d", anchor_batch, B);
  TAUV_REQUIRE((uintptr_t)in % 16 == 0 && (uintptr_t)anchor % 16 == 0 && (uintptr_t)out % 16 == 0, TAUV_E_ALIGN,
               "box tensors must be 16-byte aligned");
  const long long total = (long long)B * N;
  if (encode)
    box_codec_kernel<true><<<grid_for(total, 256), 256, 0, st>>>((const float4*)in, (const float4*)anchor, total, N,
                                                                 anchor_batch, v0, v1, (float4*)out);
  else
    box_codec_kernel<false><<<grid_for(total, 256), 256, 0, st>>>((const float4*)in, (const float4*)anchor, total, N,
                                                                  anchor_batch, v0, v1, (float4*)out);
  TAUV_LAUNCH_CHECK("box_codec_kernel");
  return 0;
}

extern "C" int tauv_yolact_box_decode(const float* encoding, const float* anchor, int B, int N, int anchor_batch,
                                      float v0, float v1, float* out, tauv_stream_t stream) {
  return codec(false, encoding, anchor, B, N, anchor_batch, v0, v1, out, (cudaStream_t)stream);
}

extern "C" int tauv_yolact_box_encode(const float* box, const float* anchor, int B, int N, int anchor_batch, float v0,
                                      float v1, float* out, tauv_stream_t stream) {
  return codec(true, box, anchor, B, N, anchor_batch, v0, v1, out, (cudaStream_t)stream);
}

extern "C" int tauv_iou_matrix(const float* a, const float* b, int Ba, int Bb, int Na, int Nb, float* out,
                               tauv_stream_t stream) {
  TAUV_REQUIRE(a && b && out, TAUV_E_NULL, "a/b/out must not be NULL");
  TAUV_REQUIRE(Ba > 0 && Bb > 0 && Na > 0 && Nb > 0, TAUV_E_SHAPE, "bad shape");
  TAUV_REQUIRE(Ba == Bb || Ba == 1 || Bb == 1, TAUV_E_SHAPE, "batch dims %d and %d do not broadcast", Ba, Bb);
  TAUV_REQUIRE((uintptr_t)a % 16 == 0 && (uintptr_t)b % 16 == 0, TAUV_E_ALIGN, "box tensors must be 16-byte aligned");
  const int B = Ba > Bb ? Ba : Bb;
  const long long total = (long long)B * Na * Nb;
  iou_matrix_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)a, (const float4*)b, Ba, Bb,
                                                                            Na, Nb, total, out);
  TAUV_LAUNCH_CHECK("iou_matrix_kernel");
  return 0;
}

extern "C" int tauv_yolact_anchors(const int* fpn_h_host, const int* fpn_w_host, int n_levels, int n_aspect,
                                   const float* hw_host, float* out, tauv_stream_t stream) {
  TAUV_REQUIRE(fpn_h_host && fpn_w_host && hw_host && out, TAUV_E_NULL, "pointers must not be NULL");
  TAUV_REQUIRE(n_levels > 0 && n_aspect > 0, TAUV_E_SHAPE, "bad shape");
  TAUV_REQUIRE(n_aspect <= 8, TAUV_E_UNSUPPORTED, "n_aspect=%d exceeds the built-in limit 8", n_aspect);
  TAUV_REQUIRE((uintptr_t)out % 16 == 0, TAUV_E_ALIGN, "out must be 16-byte aligned");
  size_t off = 0;
  for (int l = 0; l < n_levels; ++l) {
    const int H = fpn_h_host[l], W = fpn_w_host[l];
    TAUV_REQUIRE(H > 0 && W > 0, TAUV_E_SHAPE, "level %d has size %dx%d", l, H, W);
    const int n = n_aspect * H * W;
    AnchorHW hw;
    for (int a = 0; a < n_aspect; ++a) {
      hw.h[a] = hw_host[(size_t)l * 2 * n_aspect + a];
      hw.w[a] = hw_host[(size_t)l * 2 * n_aspect + n_aspect + a];
    }
    anchors_level_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(H, W, n_aspect, hw, (float4*)out + off);
    TAUV_LAUNCH_CHECK("anchors_level_kernel");
    off += n;
  }
  return 0;
}

extern "C" int tauv_box_to_mask(const float* box, int H, int W, float* out, tauv_stream_t stream) {
  TAUV_REQUIRE(box && out, TAUV_E_NULL, "box/out must not be NULL");
  TAUV_REQUIRE(H > 0 && W > 0, TAUV_E_SHAPE, "bad shape");
  box_to_mask_kernel<<<grid_for((long long)H * W, 256), 256, 0, (cudaStream_t)stream>>>(box, H, W, out);
  TAUV_LAUNCH_CHECK("box_to_mask_kernel");
  return 0;
}

extern "C" int tauv_yolact_match_anchors(const float* anchor, const float* truth_box, const uint8_t* truth_valid, int B,
                                         int N, int M, float pos_thr, float neg_thr, float v0, float v1,
                                         int64_t* match_index, float* match_iou, uint8_t* positive, uint8_t* negative,
                                         float* target, tauv_stream_t stream) {
  TAUV_REQUIRE(anchor && truth_box && truth_valid && match_index && match_iou && positive && negative, TAUV_E_NULL,
               "pointers must not be NULL");
  TAUV_REQUIRE(B > 0 && N > 0 && M > 0, TAUV_E_SHAPE, "bad shape B=%d N=%d M=%d", B, N, M);
  TAUV_REQUIRE(M <= 8192, TAUV_E_UNSUPPORTED, "M=%d exceeds the built-in limit 8192", M);
  TAUV_REQUIRE(B <= 65535, TAUV_E_UNSUPPORTED, "B=%d exceeds the built-in limit 65535", B);
  TAUV_REQUIRE((uintptr_t)anchor % 16 == 0 && (uintptr_t)truth_box % 16 == 0 && (uintptr_t)target % 16 == 0, TAUV_E_ALIGN,
               "box tensors must be 16-byte aligned");
  dim3 grid((N + 256 * kMatchTiles - 1) / (256 * kMatchTiles), B);
  const size_t smem = (size_t)M * 8 * sizeof(float);
  if (smem > 48 * 1024)
    TAUV_CUDA(ensure_dynamic_smem((const void*)(match_anchors_kernel), smem));
  match_anchors_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const float4*)anchor, (const float4*)truth_box,
                                                                  truth_valid, N, M, pos_thr, neg_thr, v0, v1, match_index,
                                                                  match_iou, positive, negative, (float4*)target);
  TAUV_LAUNCH_CHECK("match_anchors_kernel");
  return 0;
}
