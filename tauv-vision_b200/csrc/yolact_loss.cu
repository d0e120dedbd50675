// yolact_loss.cu — the classification (hard-negative mining) and box terms of the YOLACT loss on sm_100a, fused with
// the target lookup, forward and backward (YOLACT half).
//
// Replaces (reference file:line under src/tauv_vision/yolact/model/):
//   loss.py:26-34   per-frame target classes (truth class of the matched truth, 0 = background where not positive) and
//                   F.cross_entropy(reduction="none") over the [N, C1] class logits
//   loss.py:35-46   hard-negative mining: torch.topk of -softmax(...)[:, 0] over the negative priors,
//                   k = negative_example_ratio * n_positive
//   loss.py:48-56   sum of the selected priors' cross entropies, normalised by (1 + ratio) * total positives
//   loss.py:58-73   smooth-L1 between the positive priors' box encodings and their regression targets (the targets come
//                   from tauv_yolact_match_anchors), normalised by the total positives
// The reference runs ~12 ATen kernels per frame in a Python loop over the batch and materialises the [N, C1] softmax
// twice; here the class logits are read once (forward) and once more by the backward, which writes the gradient.
//
//   ycls_rows_kernel     : the only pass over the class logits: a warp stages 32 consecutive rows in shared memory
//                          (the staging of scores_tile_kernel), lane = row: log-sum-exp, the cross entropy against the
//                          row's target class, and the background confidence as a sortable key (+inf where the prior
//                          is not a negative).  8 bytes out per 4 C1 bytes in.
//   yloss_frame_kernel   : one CTA per frame: positives counted and listed, smooth-L1 over them, the k-th smallest
//                          background key by an exact radix select (ties by prior index), the selected priors'
//                          cross entropies summed in fp64 in a fixed order (deterministic).
//   ycls_backward_kernel : grad = scale * selected * (softmax - onehot): blocks of 32 rows without a selected prior
//                          (almost all) are zero-filled without being read.
//   ybox_backward_kernel : grad = scale * positive * smooth_l1'(enc - target).
#include "common.cuh"
#include "yolact_common.cuh"

namespace tauv {

constexpr int kLossTileWarps = 4;
constexpr int kLossFrameThreads = 1024;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// a warp's block of 32 consecutive rows of C1 floats -> its shared-memory tile (row r at tile + r * C1)
__device__ __forceinline__ void loss_stage_rows(float* tile, const float* src, int n_el, int lane) {
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int n16 = n_el >> 2;
    for (int q = lane; q < n16; q += 32)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(tile + 4 * q)), "l"(src + 4 * q) : "memory");
    for (int e = (n16 << 2) + lane; e < n_el; e += 32)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(tile + e)), "l"(src + e) : "memory");
  } else {
    for (int e = lane; e < n_el; e += 32)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(tile + e)), "l"(src + e) : "memory");
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();
}

// max and sum of exp(x - max) of one row in shared memory (four independent chains; ex2.approx: <= 2 ulp per term)
__device__ __forceinline__ void loss_row_lse(const float* x, int C1, float* m_out, float* sum_out) {
  float m = x[0];
  for (int c = 1; c < C1; ++c) m = fmaxf(m, x[c]);
  const float L2E = 1.4426950408889634f;
  const float ml = m * L2E;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int c = 0;
  for (; c + 4 <= C1; c += 4) {
    s0 += ex2_approx(fmaf(x[c], L2E, -ml));
    s1 += ex2_approx(fmaf(x[c + 1], L2E, -ml));
    s2 += ex2_approx(fmaf(x[c + 2], L2E, -ml));
    s3 += ex2_approx(fmaf(x[c + 3], L2E, -ml));
  }
  for (; c < C1; ++c) s0 += ex2_approx(fmaf(x[c], L2E, -ml));
  *m_out = m;
  *sum_out = (s0 + s1) + (s2 + s3);
}

__device__ __forceinline__ int loss_target_class(const uint8_t* positive, const int64_t* match_index,
                                                 const int64_t* truth_cls, long long row, int N, int M, int C1) {
  if (!positive[row]) return 0;  // loss.py:28
  const long long b = row / N;
  long long j = match_index[row];
  j = j < 0 ? 0 : (j >= M ? M - 1 : j);
  const long long t = truth_cls[b * M + j];  // loss.py:27
  return (int)(t < 0 ? 0 : (t >= C1 ? C1 - 1 : t));
}

__global__ void __launch_bounds__(kLossTileWarps * 32) ycls_rows_kernel(
    const float* __restrict__ cls, long long rows, int C1, int N, int M, const uint8_t* __restrict__ positive,
    const uint8_t* __restrict__ negative, const int64_t* __restrict__ match_index, const int64_t* __restrict__ truth_cls,
    float* __restrict__ ce, uint32_t* __restrict__ bgkey) {
  extern __shared__ __align__(16) float s_loss_tiles[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* tile = s_loss_tiles + (size_t)wib * 32 * C1;
  const long long warp = (long long)blockIdx.x * kLossTileWarps + wib;
  const long long nwarps = (long long)gridDim.x * kLossTileWarps;
  for (long long blk = warp; blk * 32 < rows; blk += nwarps) {
    const long long row0 = blk * 32;
    const int nrows = (int)min(32LL, rows - row0);
    // (the row's flags and target: in flight while the tile arrives)
    int t = 0;
    bool neg = false;
    if (lane < nrows) {
      t = loss_target_class(positive, match_index, truth_cls, row0 + lane, N, M, C1);
      neg = negative[row0 + lane] != 0;
    }
    loss_stage_rows(tile, cls + row0 * C1, nrows * C1, lane);
    if (lane < nrows) {
      const float* x = tile + lane * C1;
      float m, sum;
      loss_row_lse(x, C1, &m, &sum);
      // -log_softmax(x)[t] = (m - x_t) + log(sum)   (loss.py:30-34)
      ce[row0 + lane] = (m - x[t]) + logf(sum);
      // softmax(x)[0]   (loss.py:38); the mining ranks the negatives by it, everything else sorts last (loss.py:40-43)
      const float bg = __fdiv_rn(expf(x[0] - m), sum);
      bgkey[row0 + lane] = neg ? __float_as_uint(fmaxf(bg, 0.0f)) : 0x7f800000u;
    }
    __syncwarp();
  }
}

struct FrameLossArgs {
  const float* ce;            // [B,N]
  const uint32_t* bgkey;      // [B,N]
  const uint8_t* positive;    // [B,N]
  const float4* enc;          // [B,N,4]
  const float4* target;       // [B,N,4]
  int N, ratio;
  uint8_t* selected;          // [B,N]
  int32_t* pos_list;          // [B,N] or NULL: the positives of each frame in prior order
  double* sums;               // [B,2]: sum of the selected cross entropies, sum of the positives' smooth-L1
  int64_t* n_pos;             // [B]
};

__device__ __forceinline__ float smooth_l1(float d) {  // F.smooth_l1_loss, beta = 1
  const float a = fabsf(d);
  return a < 1.0f ? 0.5f * d * d : a - 0.5f;
}

// sum over the CTA in a fixed order; the result is valid in every thread
__device__ __forceinline__ double loss_block_sum(double v, double* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kLossFrameThreads / 32; ++w) t += s_red[w];
  return t;
}

// exclusive prefix of the warps' counts (s_cnt[warp] holds this warp's count); returns this warp's offset, *total = sum
__device__ __forceinline__ int loss_warp_offsets(int* s_cnt, int count, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_cnt[warp] = count;
  __syncthreads();
  int off = 0, tot = 0;
  for (int w = 0; w < kLossFrameThreads / 32; ++w) {
    const int c = s_cnt[w];
    if (w < warp) off += c;
    tot += c;
  }
  *total = tot;
  return off;
}

__global__ void __launch_bounds__(kLossFrameThreads, 1) yloss_frame_kernel(const FrameLossArgs a) {
  __shared__ double s_red[kLossFrameThreads / 32];
  __shared__ int s_cnt[kLossFrameThreads / 32];
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_ctl[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x, N = a.N;
  const size_t base = (size_t)b * N;
  // warp w owns the priors [w * wper, (w + 1) * wper), lanes side by side: prior order = (warp, round, lane)
  const int wper = ((N + kLossFrameThreads - 1) / kLossFrameThreads) * 32;
  const int i0 = warp * wper, i1 = min(N, i0 + wper);

  // ---- positives: count, list, smooth-L1 (loss.py:58-68) ----
  int wcount = 0;
  for (int i = i0 + lane; i - lane < i1; i += 32)
    wcount += __popc(__ballot_sync(0xffffffffu, i < i1 && a.positive[base + i] != 0));
  int n_pos;
  int woff = loss_warp_offsets(s_cnt, wcount, &n_pos);
  double box_sum = 0.0;
  for (int i = i0 + lane; i - lane < i1; i += 32) {
    const bool p = i < i1 && a.positive[base + i] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, p);
    if (p) {
      if (a.pos_list) a.pos_list[base + woff + __popc(bal & ((1u << lane) - 1u))] = i;
      const float4 e = a.enc[base + i], t = a.target[base + i];
      box_sum += (double)smooth_l1(e.x - t.x) + (double)smooth_l1(e.y - t.y) + (double)smooth_l1(e.z - t.z) +
                 (double)smooth_l1(e.w - t.w);
    }
    woff += __popc(bal);
  }
  box_sum = loss_block_sum(box_sum, s_red);

  // ---- hard negatives: the k smallest background keys, ties by prior index (loss.py:35-46) ----
  const long long kk_ll = (long long)a.ratio * n_pos;
  const int kk = (int)(kk_ll < 0 ? 0 : (kk_ll > N ? N : kk_ll));
  uint32_t T = 0u;    // the kk-th smallest key
  int need = 0;       // how many priors with key == T are taken
  if (kk > 0) {
    uint32_t prefix = 0u, pmask = 0u;
    int remaining = kk;
    for (int shift = 24; shift >= 0; shift -= 8) {
      if (tid < 256) s_hist[tid] = 0u;
      __syncthreads();
      for (int i = i0 + lane; i < i1; i += 32) {
        const uint32_t key = a.bgkey[base + i];
        if ((key & pmask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (warp == 0) {  // the digit at which the ascending cumulative count reaches `remaining`
        uint32_t c[8], run = 0u;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          c[u] = s_hist[lane * 8 + u];
          run += c[u];
        }
        uint32_t incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        uint32_t before = incl - run;
        if (before < (uint32_t)remaining && (uint32_t)remaining <= incl) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (before < (uint32_t)remaining && (uint32_t)remaining <= before + c[u]) {
              s_ctl[0] = (uint32_t)(lane * 8 + u);
              s_ctl[1] = before;
            }
            before += c[u];
          }
        }
      }
      __syncthreads();
      prefix |= s_ctl[0] << shift;
      pmask |= 255u << shift;
      remaining -= (int)s_ctl[1];
      __syncthreads();
    }
    T = prefix;
    need = remaining;  // 1 <= need <= number of keys equal to T
  }
  // ties at T in prior order: this warp's offset among them
  int wties = 0;
  if (kk > 0)
    for (int i = i0 + lane; i - lane < i1; i += 32)
      wties += __popc(__ballot_sync(0xffffffffu, i < i1 && a.bgkey[base + i] == T));
  int n_ties;
  int toff = loss_warp_offsets(s_cnt, wties, &n_ties);

  // ---- the selected priors (loss.py:48-52) and their cross entropies, summed in a fixed order ----
  double cls_sum = 0.0;
  for (int i = i0 + lane; i - lane < i1; i += 32) {
    const bool in = i < i1;
    const uint32_t key = in ? a.bgkey[base + i] : 0xffffffffu;
    const bool tie = in && kk > 0 && key == T;
    const unsigned bal = __ballot_sync(0xffffffffu, tie);
    bool sel = in && a.positive[base + i] != 0;
    if (in && kk > 0 && key < T) sel = true;
    if (tie && toff + __popc(bal & ((1u << lane) - 1u)) < need) sel = true;
    toff += __popc(bal);
    if (in) {
      a.selected[base + i] = sel ? 1 : 0;
      if (sel) cls_sum += (double)a.ce[base + i];
    }
  }
  cls_sum = loss_block_sum(cls_sum, s_red);
  if (tid == 0) {
    a.sums[(size_t)b * 2 + 0] = cls_sum;
    a.sums[(size_t)b * 2 + 1] = box_sum;
    a.n_pos[b] = n_pos;
  }
}

// scale of the gradients: grad_out / ((1 + ratio) * P) for the class term, grad_out / P for the box term, with P the
// batch's positives; plain grad_out when there are none (loss.py:54-57, :70-73)
__device__ __forceinline__ float loss_scale(const float* grad_out, const int64_t* n_pos_total, int mult) {
  const long long P = *n_pos_total;
  const float g = *grad_out;
  return P > 0 ? __fdiv_rn(g, (float)(mult * P)) : g;
}

__global__ void __launch_bounds__(kLossTileWarps * 32) ycls_backward_kernel(
    const float* __restrict__ cls, long long rows, int C1, int N, int M, const uint8_t* __restrict__ positive,
    const uint8_t* __restrict__ selected, const int64_t* __restrict__ match_index, const int64_t* __restrict__ truth_cls,
    const float* __restrict__ grad_out, const int64_t* __restrict__ n_pos_total, int ratio, float* __restrict__ grad) {
  extern __shared__ __align__(16) float s_loss_tiles[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* tile = s_loss_tiles + (size_t)wib * 32 * C1;
  const long long warp = (long long)blockIdx.x * kLossTileWarps + wib;
  const long long nwarps = (long long)gridDim.x * kLossTileWarps;
  const float scale = loss_scale(grad_out, n_pos_total, 1 + ratio);
  for (long long blk = warp; blk * 32 < rows; blk += nwarps) {
    const long long row0 = blk * 32;
    const int nrows = (int)min(32LL, rows - row0);
    const int n_el = nrows * C1;
    const bool sel = lane < nrows && selected[row0 + lane] != 0;
    float* dst = grad + row0 * C1;
    if (!__any_sync(0xffffffffu, sel)) {  // (all but a few per cent of the blocks)
      if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {   // 128-bit stores: a quarter of the instructions
        const int n16 = n_el >> 2;
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int q = lane; q < n16; q += 32) d4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = (n16 << 2) + lane; e < n_el; e += 32) dst[e] = 0.0f;
      } else {
        for (int e = lane; e < n_el; e += 32) dst[e] = 0.0f;
      }
      continue;
    }
    const int t = sel ? loss_target_class(positive, match_index, truth_cls, row0 + lane, N, M, C1) : 0;
    loss_stage_rows(tile, cls + row0 * C1, n_el, lane);
    if (lane < nrows) {
      float* x = tile + lane * C1;
      if (sel) {
        float m, sum;
        loss_row_lse(x, C1, &m, &sum);
        const float inv = __fdiv_rn(1.0f, sum);
        for (int c = 0; c < C1; ++c) {
          const float p = expf(x[c] - m) * inv;
          x[c] = scale * (c == t ? p - 1.0f : p);
        }
      } else {
        for (int c = 0; c < C1; ++c) x[c] = 0.0f;
      }
    }
    __syncwarp();
    for (int e = lane; e < n_el; e += 32) dst[e] = tile[e];
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) ybox_backward_kernel(const float4* __restrict__ enc, const float4* __restrict__ target,
                                                            const uint8_t* __restrict__ positive, long long rows,
                                                            const float* __restrict__ grad_out,
                                                            const int64_t* __restrict__ n_pos_total,
                                                            float4* __restrict__ grad) {
  const float scale = loss_scale(grad_out, n_pos_total, 1);
  auto d1 = [scale](float d) { return scale * (fabsf(d) < 1.0f ? d : (d > 0.0f ? 1.0f : -1.0f)); };
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (positive[i]) {
      const float4 e = enc[i], t = target[i];
      g = make_float4(d1(e.x - t.x), d1(e.y - t.y), d1(e.z - t.z), d1(e.w - t.w));
    }
    grad[i] = g;
  }
}

static int loss_tile_grid(long long rows) {
  const long long blocks = (rows + 32 * kLossTileWarps - 1) / (32 * kLossTileWarps);
  const long long cap = (long long)num_sms() * 16;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}


// ---- mask term (loss.py:75-121) -----------------------------------------------------------------------------------
// Per positive prior i of frame b (matched truth j): mask = clamp(sigmoid(coeff_i . proto_b), 1e-4), truth mask =
// bilinear resize of (seg_b == j) to the prototype grid, BCE of the two (both clamped to [1e-4, 1 - 1e-4]) weighted by
// box_to_mask(truth_box_j) * nearest-resized img_valid_b, summed and divided by the resized truth mask's area; positives
// whose resized truth mask is empty are skipped.  The reference loops over the positives in Python (~25 ATen kernels
// and a [SH, SW] -> [PH, PW] F.interpolate each); here
//   ymask_area_kernel           : the areas of all truths' resized masks in one pass per frame (fixed-point integer
//                                 atomics: order-independent) — positives matched to one truth share them;
//   ymask_records_kernel        : crop box, area, prior and truth of every listed positive, gathered once per call;
//   ymask_positive_kernel<0>    : CTAs walk the frame's positives (pos_list of yloss_frame_kernel), threads the pixels
//                                 of the truth box (the weight is zero outside it);
//   ymask_positive_kernel<1>    : same walk, d/d coeff_i = sum over pixels of dlogit * proto (block reduction);
//   ymask_backward_proto_kernel : a thread per pixel walks the frame's positives (records staged in shared memory) with
//                                 the pixel's K prototype values in registers, d/d proto = sum over positives of
//                                 dlogit * coeff_i — no floating-point atomics, so both gradients are deterministic.
// Inside the clamps d BCE / d logit = sigmoid - truth (F.binary_cross_entropy's backward times sigmoid'), outside 0.
constexpr int kMaskLossThreads = 256;
constexpr int kMaskLossMaxK = 32;

// What the walks need to know about a positive, gathered once per call: reaching it through pos_list -> match_index ->
// truth_box / area is a chain of four dependent loads (~2.5 us) that every CTA paid per positive (the per-pixel proto
// backward: per chunk of positives, with two CTAs resident per SM — most of its time).
struct __align__(16) MaskRec {
  float left, right, top, bottom;   // crop_bounds of the matched truth's box on the prototype grid
  float area;                       // of the truth's resized mask (0: the positive is skipped, loss.py:93-94)
  int n, j;                         // prior, truth
  int pad;
};

// The segmentation map as the caller holds it: uint8 in the reference's dataset (segmentation_dataset.py:98-99), int32 or
// int64 elsewhere — read in place (a converted copy of 64 x 550 x 550 values is a pass of its own).
struct SegView {
  const void* p;
  int bytes;   // 1 (unsigned), 4 or 8 (signed)
  __device__ __forceinline__ int operator[](int i) const {
    if (bytes == 1) return (int)static_cast<const uint8_t*>(p)[i];
    if (bytes == 4) return static_cast<const int32_t*>(p)[i];
    const long long v = static_cast<const long long*>(p)[i];
    return v < -1 || v > 0x7fffffffLL ? -1 : (int)v;   // (no truth index lies out there)
  }
};

struct MaskLossArgs {
  const float* coeff;          // [B,N,K]
  const float* proto;          // [B,K,PH,PW]
  const int32_t* pos_list;     // [B,N]
  const int64_t* n_pos;        // [B]
  const int64_t* match_index;  // [B,N]
  const float* truth_box;      // [B,M,4]
  const void* seg;             // [B,SH,SW], seg_bytes per element
  int seg_bytes;
  const uint8_t* img_valid;    // [B,SH,SW]
  int N, K, M, PH, PW, SH, SW;
  float sy, sx;                // (float)SH / PH, (float)SW / PW: ATen's area_pixel_compute_scale without align_corners
  double* tsum;                // [B,M]: area of every truth's resized mask (u64 fixed point while ymask_area_kernel adds)
  struct MaskRec* recs;        // [B,N]: one record per listed positive (ymask_records_kernel), first n_pos[b] entries
  double* partial;             // [B,gridDim.x] (forward)
  const float* grad_out;       // [1] (backward)
  const int64_t* n_pos_total;  // [1] (backward)
  float* grad_coeff;           // [B,N,K], zero-filled by the caller
  float* grad_proto;           // [B,K,PH,PW]
};

struct MaskPx {
  int o00, o01, o10, o11;      // the four taps of the bilinear resize (offsets into the frame's seg map)
  float ly0, ly1, lx0, lx1;
  float valid;                 // nearest-resized img_valid
  float fy, fx;
};

__device__ __forceinline__ MaskPx mask_px(const MaskLossArgs& a, int b, int y, int x) {
  MaskPx g;
  // upsample_bilinear2d, align_corners = False: src = scale * (dst + 0.5) - 0.5, clamped at 0
  float ys = __fsub_rn(__fmul_rn(a.sy, __fadd_rn((float)y, 0.5f)), 0.5f);
  float xs = __fsub_rn(__fmul_rn(a.sx, __fadd_rn((float)x, 0.5f)), 0.5f);
  ys = ys < 0.0f ? 0.0f : ys;
  xs = xs < 0.0f ? 0.0f : xs;
  const int y0 = min((int)ys, a.SH - 1), x0 = min((int)xs, a.SW - 1);
  const int y1 = y0 + (y0 < a.SH - 1 ? 1 : 0), x1 = x0 + (x0 < a.SW - 1 ? 1 : 0);
  g.ly1 = __fsub_rn(ys, (float)y0);
  g.ly0 = __fsub_rn(1.0f, g.ly1);
  g.lx1 = __fsub_rn(xs, (float)x0);
  g.lx0 = __fsub_rn(1.0f, g.lx1);
  g.o00 = y0 * a.SW + x0;
  g.o01 = y0 * a.SW + x1;
  g.o10 = y1 * a.SW + x0;
  g.o11 = y1 * a.SW + x1;
  // upsample_nearest2d: src = min(floor(dst * scale), in - 1)
  const int yn = min((int)floorf(__fmul_rn((float)y, a.sy)), a.SH - 1), xn = min((int)floorf(__fmul_rn((float)x, a.sx)), a.SW - 1);
  g.valid = a.img_valid[(size_t)b * a.SH * a.SW + (size_t)yn * a.SW + xn] ? 1.0f : 0.0f;
  g.fy = (float)y;
  g.fx = (float)x;
  return g;
}

__device__ __forceinline__ float mask_truth(const MaskPx& g, const SegView& seg, int j) {
  const float v00 = seg[g.o00] == j ? 1.0f : 0.0f, v01 = seg[g.o01] == j ? 1.0f : 0.0f;
  const float v10 = seg[g.o10] == j ? 1.0f : 0.0f, v11 = seg[g.o11] == j ? 1.0f : 0.0f;
  return __fadd_rn(__fmul_rn(g.ly0, __fadd_rn(__fmul_rn(g.lx0, v00), __fmul_rn(g.lx1, v01))),
                   __fmul_rn(g.ly1, __fadd_rn(__fmul_rn(g.lx0, v10), __fmul_rn(g.lx1, v11))));
}

// the same from the four tap values (the taps of a pixel do not depend on the truth: loaded once, compared many times)
__device__ __forceinline__ float mask_truth_taps(const MaskPx& g, int s00, int s01, int s10, int s11, int j) {
  const float v00 = s00 == j ? 1.0f : 0.0f, v01 = s01 == j ? 1.0f : 0.0f;
  const float v10 = s10 == j ? 1.0f : 0.0f, v11 = s11 == j ? 1.0f : 0.0f;
  return __fadd_rn(__fmul_rn(g.ly0, __fadd_rn(__fmul_rn(g.lx0, v00), __fmul_rn(g.lx1, v01))),
                   __fmul_rn(g.ly1, __fadd_rn(__fmul_rn(g.lx0, v10), __fmul_rn(g.lx1, v11))));
}

__device__ __forceinline__ float mask_weight(const MaskPx& g, const CropBounds& c) {
  return (g.fx >= c.left && g.fx <= c.right && g.fy >= c.top && g.fy <= c.bottom) ? g.valid : 0.0f;
}

__device__ __forceinline__ float clamp_unit(float v) { return fminf(fmaxf(v, 1e-4f), 1.0f - 1e-4f); }

// d (w * BCE) / d logit: w * (s - t) where both clamps pass the gradient (s >= 1e-4 and max(s, 1e-4) <= 1 - 1e-4)
__device__ __forceinline__ float mask_dlogit(float logit, float t, float w) {
  const float s = sigmoid_ref(logit);
  if (!(s >= 1e-4f && s <= 1.0f - 1e-4f)) return 0.0f;
  return w * (s - clamp_unit(t));
}

__device__ __forceinline__ double mask_block_sum(double v, double* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kMaskLossThreads / 32; ++w) t += s_red[w];
  return t;
}

// the pixels of the prototype grid a crop box can cover (a superset: mask_weight decides pixel by pixel)
struct MaskBoxRange {
  int y0, x0, bh, bw;
};
__device__ __forceinline__ MaskBoxRange mask_box_range(const CropBounds& c, int PH, int PW) {
  auto lo = [](float v, int n) { return v > 0.0f ? (v < (float)n ? (int)ceilf(v) : n) : 0; };           // (NaN -> 0)
  auto hi = [](float v, int n) { return v < (float)(n - 1) ? (v >= 0.0f ? (int)floorf(v) : -1) : n - 1; };  // (NaN -> n - 1)
  MaskBoxRange r;
  r.y0 = lo(c.top, PH);
  r.x0 = lo(c.left, PW);
  r.bh = max(hi(c.bottom, PH) - r.y0 + 1, 0);
  r.bw = max(hi(c.right, PW) - r.x0 + 1, 0);
  return r;
}

// area[b][j] = sum over the prototype grid of the bilinearly resized mask of truth j (loss.py:86-93, :113), for all truths
// of a frame in ONE pass over the grid: a pixel's four taps each add their bilinear weight to the truth they show.
// The sums are kept in 2^-40 fixed point (64-bit integer atomics, shared memory first), so they do not depend on the
// order of the additions: deterministic, and within 1e-7 of the fp32 sum of the resized mask.
constexpr double kAreaScale = 1099511627776.0;  // 2^40
__global__ void __launch_bounds__(kMaskLossThreads) ymask_area_kernel(const MaskLossArgs a) {
  extern __shared__ unsigned long long s_area[];  // [M]
  const int b = blockIdx.y;
  const int HW = a.PH * a.PW;
  const SegView seg{static_cast<const char*>(a.seg) + (size_t)b * a.SH * a.SW * a.seg_bytes, a.seg_bytes};
  for (int m = threadIdx.x; m < a.M; m += kMaskLossThreads) s_area[m] = 0ull;
  __syncthreads();
  const int px = blockIdx.x * kMaskLossThreads + threadIdx.x;
  if (px < HW) {
    const int y = px / a.PW, x = px - y * a.PW;
    const MaskPx g = mask_px(a, b, y, x);
    const int sv[4] = {seg[g.o00], seg[g.o01], seg[g.o10], seg[g.o11]};
    const float wv[4] = {__fmul_rn(g.ly0, g.lx0), __fmul_rn(g.ly0, g.lx1), __fmul_rn(g.ly1, g.lx0), __fmul_rn(g.ly1, g.lx1)};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double w = 0.0;   // taps that show the same truth are added first: one atomic per distinct truth
      bool first = true;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (r < q && sv[r] == sv[q]) first = false;
        if (r >= q && sv[r] == sv[q]) w += (double)wv[r];
      }
      if (first && sv[q] >= 0 && sv[q] < a.M && w > 0.0) atomicAdd(&s_area[sv[q]], (unsigned long long)llrint(w * kAreaScale));
    }
  }
  __syncthreads();
  unsigned long long* fx = reinterpret_cast<unsigned long long*>(a.tsum) + (size_t)b * a.M;
  for (int m = threadIdx.x; m < a.M; m += kMaskLossThreads)
    if (s_area[m]) atomicAdd(&fx[m], s_area[m]);
}

__global__ void ymask_area_finish_kernel(double* tsum, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) tsum[t] = (double)reinterpret_cast<const unsigned long long*>(tsum)[t] / kAreaScale;
}

__global__ void __launch_bounds__(256) ymask_records_kernel(const MaskLossArgs a) {
  const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  if (i >= (int)a.n_pos[b]) return;
  const int n = a.pos_list[(size_t)b * a.N + i];
  long long jl = a.match_index[(size_t)b * a.N + n];
  const int j = (int)(jl < 0 ? 0 : (jl >= a.M ? a.M - 1 : jl));
  const float* tb = a.truth_box + ((size_t)b * a.M + j) * 4;
  const CropBounds c = crop_bounds(make_float4(tb[0], tb[1], tb[2], tb[3]), a.PH, a.PW);
  a.recs[(size_t)b * a.N + i] = MaskRec{c.left, c.right, c.top, c.bottom, (float)a.tsum[(size_t)b * a.M + j], n, j, 0};
}

// The K-term dot products and gradient accumulations use explicit fmaf: the library is built with --fmad=false (every
// expression that feeds a comparison rounds like ATen), which would make each term an FMUL + FADD pair — twice the
// instructions of these issue-bound loops; the reference's own `@` (loss.py:82) is a BLAS matmul whose rounding is not
// specified either.
// (FULLK: K == 32, the YOLACT head — the per-channel `k < K` tests of the unrolled loops cost an ISETP each per pixel,
// there are only seven predicate registers to keep them in)
template <bool BACKWARD, bool FULLK>
__global__ void __launch_bounds__(kMaskLossThreads, BACKWARD ? 2 : 3) ymask_positive_kernel(const MaskLossArgs a) {
  __shared__ float s_coeff2[2][kMaskLossMaxK];
  __shared__ double s_red[kMaskLossThreads / 32];
  __shared__ float s_gc[kMaskLossThreads / 32][kMaskLossMaxK];
  const int tid = threadIdx.x, b = blockIdx.y;
  const int npos = (int)a.n_pos[b];
  const int HW = a.PH * a.PW;
  const float* proto = a.proto + (size_t)b * a.K * HW;
  const SegView seg{static_cast<const char*>(a.seg) + (size_t)b * a.SH * a.SW * a.seg_bytes, a.seg_bytes};
  // forward: every thread keeps its own sum over the CTA's positives of (its pixels' BCE) / area — ONE block sum at the
  // end instead of one per positive (ten shuffles and two barriers each: 15 % of the stall samples), and the coefficient
  // row double-buffered: one barrier per positive instead of four.  The order of the additions is fixed: run-to-run identical.
  double thread_sum = 0.0;
  float gscale = 0.0f;
  if (BACKWARD) {
    const long long P = *a.n_pos_total;
    gscale = P > 0 ? __fdiv_rn(*a.grad_out, (float)P) : *a.grad_out;   // loss.py:117-120
  }
  int it = 0;
  for (int i = blockIdx.x; i < npos; i += gridDim.x, ++it) {
    const MaskRec rec = a.recs[(size_t)b * a.N + i];
    const int n = rec.n, j = rec.j;
    const CropBounds crop{rec.left, rec.right, rec.top, rec.bottom};
    const MaskBoxRange box = mask_box_range(crop, a.PH, a.PW);
    const int npx = box.bh * box.bw;
    const int bwd = max(box.bw, 1), step_y = kMaskLossThreads / bwd, step_x = kMaskLossThreads - step_y * bwd;
    const float area = rec.area;
    float* s_coeff = s_coeff2[BACKWARD ? 0 : (it & 1)];
    if (BACKWARD) __syncthreads();   // (forward: the buffer written here was last read two positives ago, a barrier in between)
    if (tid < kMaskLossMaxK) s_coeff[tid] = tid < a.K ? a.coeff[((size_t)b * a.N + n) * a.K + tid] : 0.0f;
    __syncthreads();
    if (!BACKWARD) {
      double num = 0.0;
      if (area > 0.0f) {   // loss.py:93-94
        const double inv_area = 1.0 / (double)area;   // (independent of the pixel loop: its latency hides under the loads)
        int ry = tid / bwd, rx = tid - ry * bwd;   // (the pixel's row and column inside the box, kept incrementally)
        for (int q = tid; q < npx; q += kMaskLossThreads, ry += step_y, rx += step_x) {
          if (rx >= bwd) { rx -= bwd; ++ry; }
          const int y = box.y0 + ry, x = box.x0 + rx, px = y * a.PW + x;
          const MaskPx g = mask_px(a, b, y, x);
          const float w = mask_weight(g, crop);
          if (w != 0.0f) {
            float pv[kMaskLossMaxK];   // (all K loads in flight before the first use: the loop is latency-bound)
#pragma unroll
            for (int k = 0; k < kMaskLossMaxK; ++k) pv[k] = (FULLK || k < a.K) ? proto[(size_t)k * HW + px] : 0.0f;
            float logit = 0.0f;
#pragma unroll
            for (int k = 0; k < kMaskLossMaxK; ++k) logit = fmaf(s_coeff[k], pv[k], logit);   // loss.py:82 (s_coeff and pv are zero beyond K)
            const float m = clamp_unit(fmaxf(sigmoid_ref(logit), 1e-4f)), tc = clamp_unit(mask_truth(g, seg, j));  // :83-84, :97-98
            num += (double)(w * -(tc * logf(m) + (1.0f - tc) * logf(1.0f - m)));               // :96-100, :113
          }
        }
        thread_sum += num * inv_area;   // :113
      }
    } else {
      float gc[kMaskLossMaxK];
#pragma unroll
      for (int k = 0; k < kMaskLossMaxK; ++k) gc[k] = 0.0f;
      if (area > 0.0f) {
        const float G = __fdiv_rn(gscale, area);
        int ry = tid / bwd, rx = tid - ry * bwd;
        for (int q = tid; q < npx; q += kMaskLossThreads, ry += step_y, rx += step_x) {
          if (rx >= bwd) { rx -= bwd; ++ry; }
          const int y = box.y0 + ry, x = box.x0 + rx, px = y * a.PW + x;
          const MaskPx g = mask_px(a, b, y, x);
          const float w = mask_weight(g, crop);
          if (w == 0.0f) continue;
          // sixteen prototype loads in flight at a time; the first sixteen are read again (from L1) for the
          // accumulation — holding all 32 next to the 32 accumulators took 219 registers: one CTA per SM
          float logit = 0.0f;
          float pv[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) pv[k] = (FULLK || k < a.K) ? proto[(size_t)k * HW + px] : 0.0f;
#pragma unroll
          for (int k = 0; k < 16; ++k) logit = fmaf(s_coeff[k], pv[k], logit);
#pragma unroll
          for (int k = 0; k < 16; ++k) pv[k] = (FULLK || 16 + k < a.K) ? proto[(size_t)(16 + k) * HW + px] : 0.0f;
#pragma unroll
          for (int k = 0; k < 16; ++k) logit = fmaf(s_coeff[16 + k], pv[k], logit);
          const float dl = G * mask_dlogit(logit, mask_truth(g, seg, j), w);
#pragma unroll
          for (int k = 0; k < 16; ++k) gc[16 + k] = fmaf(dl, pv[k], gc[16 + k]);
#pragma unroll
          for (int k = 0; k < 16; ++k) gc[k] = fmaf(dl, (FULLK || k < a.K) ? proto[(size_t)k * HW + px] : 0.0f, gc[k]);
        }
      }
      // K sums over the CTA: warp shuffles, then the eight warps' values in a fixed order
      // (a butterfly that halves the values a lane holds at every step — 31 shuffles instead of 32 x 5: lane l ends up
      // with the warp's sum for channel l; the order of the additions is fixed)
      const int lane = tid & 31, warp = tid >> 5;
      static_assert(kMaskLossMaxK == 32, "the butterfly below is written for 32 channels");
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < o; ++k) {
          const float send = upper ? gc[k] : gc[k + o], keep = upper ? gc[k + o] : gc[k];
          gc[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      s_gc[warp][lane] = gc[0];
      __syncthreads();
      if (tid < a.K) {
        float v = 0.0f;
        for (int w = 0; w < kMaskLossThreads / 32; ++w) v += s_gc[w][tid];
        a.grad_coeff[((size_t)b * a.N + n) * a.K + tid] = v;
      }
    }
  }
  if (!BACKWARD) {
    const double cta_sum = mask_block_sum(thread_sum, s_red);
    if (tid == 0) a.partial[(size_t)b * gridDim.x + blockIdx.x] = cta_sum;
  }
}

// one record per positive, staged in shared memory for the CTA's 256 pixels
struct MaskPosRec {
  float left, right, top, bottom;
  float G;      // grad_out / (P * area), 0 = skipped
  int n, j;
};
constexpr int kMaskRecChunk = 64;
constexpr int kMaskTileW = 32, kMaskTileH = kMaskLossThreads / 32;   // ymask_backward_proto_kernel's tile of pixels

template <bool FULLK>
__global__ void __launch_bounds__(kMaskLossThreads, 2) ymask_backward_proto_kernel(const MaskLossArgs a) {
  __shared__ MaskPosRec s_rec[kMaskRecChunk];
  __shared__ float s_cf[kMaskRecChunk][kMaskLossMaxK];   // the chunk's coefficient rows
  const int b = blockIdx.y;
  const int HW = a.PH * a.PW;
  // a CTA is a tile of 32 x 8 pixels (a warp = 32 consecutive pixels of a row: coalesced): a crop box covers most of a
  // tile or misses it, so the eight warps between two barriers have the same work, and fewer boxes touch a tile than touch
  // a whole row of the grid (256 consecutive pixels, the first layout: `barrier` was its top stall reason)
  const int tiles_x = (a.PW + kMaskTileW - 1) / kMaskTileW;
  const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x - tile_y * tiles_x;
  const int y = tile_y * kMaskTileH + (threadIdx.x >> 5), x = tile_x * kMaskTileW + (threadIdx.x & 31);
  const bool live = y < a.PH && x < a.PW;
  const int px = live ? y * a.PW + x : 0;
  const int npos = (int)a.n_pos[b];
  const float* proto = a.proto + (size_t)b * a.K * HW;
  const SegView seg{static_cast<const char*>(a.seg) + (size_t)b * a.SH * a.SW * a.seg_bytes, a.seg_bytes};
  const long long P = *a.n_pos_total;
  const float gscale = P > 0 ? __fdiv_rn(*a.grad_out, (float)P) : *a.grad_out;
  float pv[kMaskLossMaxK], gp[kMaskLossMaxK];
#pragma unroll
  for (int k = 0; k < kMaskLossMaxK; ++k) {
    pv[k] = (live && (FULLK || k < a.K)) ? proto[(size_t)k * HW + px] : 0.0f;
    gp[k] = 0.0f;
  }
  const MaskPx g = mask_px(a, b, live ? y : 0, live ? x : 0);
  // (the four taps of this pixel's bilinear resize do not depend on the positive: loaded once)
  const int s00 = live ? seg[g.o00] : -1, s01 = live ? seg[g.o01] : -1, s10 = live ? seg[g.o10] : -1, s11 = live ? seg[g.o11] : -1;
  // the rows and columns this CTA's tile spans: positives whose crop box misses it are not staged at all
  const float fy_lo = (float)(tile_y * kMaskTileH), fy_hi = (float)min(tile_y * kMaskTileH + kMaskTileH - 1, a.PH - 1);
  const float fx_lo = (float)(tile_x * kMaskTileW), fx_hi = (float)min(tile_x * kMaskTileW + kMaskTileW - 1, a.PW - 1);
  __shared__ int s_kept[2];
  static_assert(kMaskRecChunk == 64, "the ordered compaction below uses two warps");
  for (int i0 = 0; i0 < npos; i0 += kMaskRecChunk) {
    __syncthreads();
    if (threadIdx.x < kMaskRecChunk) {   // (warps 0 and 1: one candidate per lane, kept in list order)
      const int r = threadIdx.x, lane = r & 31;
      MaskPosRec rec{};
      bool keep = false;
      if (i0 + r < npos) {
        const MaskRec c = a.recs[(size_t)b * a.N + i0 + r];
        rec = MaskPosRec{c.left, c.right, c.top, c.bottom, c.area > 0.0f ? __fdiv_rn(gscale, c.area) : 0.0f, c.n, c.j};
        keep = rec.G != 0.0f && c.bottom >= fy_lo && c.top <= fy_hi && c.right >= fx_lo && c.left <= fx_hi;  // (false for NaN)
      }
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_kept[r >> 5] = __popc(bal);
      asm volatile("bar.sync 1, 64;" ::: "memory");
      const int slot = (r >= 32 ? s_kept[0] : 0) + __popc(bal & ((1u << lane) - 1u));
      if (keep) s_rec[slot] = rec;
    }
    __syncthreads();
    const int nr = s_kept[0] + s_kept[1];
    for (int e = threadIdx.x; e < nr * kMaskLossMaxK; e += kMaskLossThreads) {
      const int r = e / kMaskLossMaxK, k = e - r * kMaskLossMaxK;
      s_cf[r][k] = k < a.K ? a.coeff[((size_t)b * a.N + s_rec[r].n) * a.K + k] : 0.0f;   // (zero beyond K)
    }
    __syncthreads();
    for (int r = 0; r < nr && live; ++r) {
      const MaskPosRec& rec = s_rec[r];
      if (!(g.fx >= rec.left && g.fx <= rec.right && g.fy >= rec.top && g.fy <= rec.bottom) || g.valid == 0.0f) continue;
      const float* cf = s_cf[r];   // (the same address in every thread: shared-memory broadcasts)
      float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f;
#pragma unroll
      for (int k = 0; k < kMaskLossMaxK; k += 4) {
        l0 = fmaf(cf[k], pv[k], l0);   // (cf and pv are zero beyond K)
        l1 = fmaf(cf[k + 1], pv[k + 1], l1);
        l2 = fmaf(cf[k + 2], pv[k + 2], l2);
        l3 = fmaf(cf[k + 3], pv[k + 3], l3);
      }
      const float dl = rec.G * mask_dlogit((l0 + l1) + (l2 + l3), mask_truth_taps(g, s00, s01, s10, s11, rec.j), g.valid);
#pragma unroll
      for (int k = 0; k < kMaskLossMaxK; ++k)
        if (FULLK || k < a.K) gp[k] = fmaf(dl, cf[k], gp[k]);
    }
  }
  if (live) {
#pragma unroll
    for (int k = 0; k < kMaskLossMaxK; ++k)
      if (FULLK || k < a.K) a.grad_proto[((size_t)b * a.K + k) * HW + px] = gp[k];
  }
}

// The normalisation of the three terms (loss.py:54-57, :70-73, :117-120) in one small launch: P = the batch's positives;
// class term = sum / ((1 + ratio) P), box and mask terms = sum / P, undivided when P == 0.  Fixed summation order.
__global__ void __launch_bounds__(256) yloss_reduce_kernel(const double* __restrict__ sums, const int64_t* __restrict__ n_pos,
                                                           int B, int ratio, const double* __restrict__ mask_partial,
                                                           int n_partial, float* __restrict__ losses,
                                                           int64_t* __restrict__ n_pos_total) {
  __shared__ double s_v[3][256];
  __shared__ long long s_p[256];
  const int tid = threadIdx.x;
  double c = 0.0, bx = 0.0, mk = 0.0;
  long long P = 0;
  for (int i = tid; i < B; i += 256) {
    P += n_pos[i];
    if (sums) {
      c += sums[2 * i];
      bx += sums[2 * i + 1];
    }
  }
  if (mask_partial)
    for (int i = tid; i < n_partial; i += 256) mk += mask_partial[i];
  s_v[0][tid] = c;
  s_v[1][tid] = bx;
  s_v[2][tid] = mk;
  s_p[tid] = P;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      s_v[0][tid] += s_v[0][tid + o];
      s_v[1][tid] += s_v[1][tid + o];
      s_v[2][tid] += s_v[2][tid + o];
      s_p[tid] += s_p[tid + o];
    }
    __syncthreads();
  }
  if (tid == 0) {
    const long long Pt = s_p[0];
    const double d = Pt > 0 ? (double)Pt : 1.0;
    if (n_pos_total) *n_pos_total = Pt;
    if (sums) {
      losses[0] = (float)(Pt > 0 ? s_v[0][0] / ((1.0 + ratio) * d) : s_v[0][0]);
      losses[1] = (float)(Pt > 0 ? s_v[1][0] / d : s_v[1][0]);
    }
    if (mask_partial) losses[2] = (float)(Pt > 0 ? s_v[2][0] / d : s_v[2][0]);
  }
}

constexpr int kMaskLossWalkers = 32;  // CTAs that share a frame's positives

}  // namespace tauv

using namespace tauv;

extern "C" size_t tauv_yolact_class_box_loss_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return align_up((size_t)B * N * sizeof(float), 256) + align_up((size_t)B * N * sizeof(uint32_t), 256);
}

extern "C" int tauv_yolact_class_box_loss(const float* cls, const float* enc, const float* target,
                                          const uint8_t* positive, const uint8_t* negative, const int64_t* match_index,
                                          const int64_t* truth_cls, int B, int N, int C1, int M, int ratio,
                                          uint8_t* selected, int32_t* pos_list, double* sums, int64_t* n_pos,
                                          void* workspace, size_t workspace_bytes, tauv_stream_t stream) {
  TAUV_REQUIRE(cls && enc && target && positive && negative && match_index && truth_cls && selected && sums && n_pos &&
                   workspace,
               TAUV_E_NULL, "pointers must not be NULL");
  TAUV_REQUIRE(B > 0 && N > 0 && C1 > 0 && M > 0 && ratio >= 0, TAUV_E_SHAPE, "bad shape B=%d N=%d C1=%d M=%d ratio=%d", B, N,
               C1, M, ratio);
  TAUV_REQUIRE(C1 <= 256, TAUV_E_UNSUPPORTED, "C1=%d exceeds the built-in limit 256", C1);
  TAUV_REQUIRE(B <= 65535, TAUV_E_UNSUPPORTED, "B=%d exceeds the built-in limit 65535", B);
  TAUV_REQUIRE((uintptr_t)enc % 16 == 0 && (uintptr_t)target % 16 == 0 && (uintptr_t)cls % 4 == 0, TAUV_E_ALIGN,
               "box tensors must be 16-byte aligned");
  TAUV_REQUIRE(workspace_bytes >= tauv_yolact_class_box_loss_workspace_bytes(B, N) && (uintptr_t)workspace % 256 == 0,
               TAUV_E_WORKSPACE, "workspace too small or not 256-byte aligned");
  float* ce = (float*)workspace;
  uint32_t* bgkey = (uint32_t*)((char*)workspace + align_up((size_t)B * N * sizeof(float), 256));
  const long long rows = (long long)B * N;
  const size_t smem = (size_t)kLossTileWarps * 32 * C1 * sizeof(float);
  if (smem > 48 * 1024) TAUV_CUDA(ensure_dynamic_smem((const void*)ycls_rows_kernel, smem));
  ycls_rows_kernel<<<loss_tile_grid(rows), kLossTileWarps * 32, smem, (cudaStream_t)stream>>>(
      cls, rows, C1, N, M, positive, negative, match_index, truth_cls, ce, bgkey);
  TAUV_LAUNCH_CHECK("ycls_rows_kernel");
  FrameLossArgs a{ce, bgkey, positive, (const float4*)enc, (const float4*)target, N, ratio, selected, pos_list, sums, n_pos};
  yloss_frame_kernel<<<B, kLossFrameThreads, 0, (cudaStream_t)stream>>>(a);
  TAUV_LAUNCH_CHECK("yloss_frame_kernel");
  return 0;
}

extern "C" int tauv_yolact_class_box_loss_backward(const float* cls, const float* enc, const float* target,
                                                   const uint8_t* positive, const uint8_t* selected,
                                                   const int64_t* match_index, const int64_t* truth_cls, int B, int N,
                                                   int C1, int M, int ratio, const int64_t* n_pos_total,
                                                   const float* grad_cls_loss, const float* grad_box_loss,
                                                   float* grad_cls, float* grad_enc, tauv_stream_t stream) {
  TAUV_REQUIRE(cls && enc && target && positive && selected && match_index && truth_cls && n_pos_total, TAUV_E_NULL,
               "pointers must not be NULL");
  TAUV_REQUIRE((grad_cls == nullptr) == (grad_cls_loss == nullptr) && (grad_enc == nullptr) == (grad_box_loss == nullptr),
               TAUV_E_NULL, "each gradient output needs its incoming gradient");
  TAUV_REQUIRE(B > 0 && N > 0 && C1 > 0 && M > 0 && ratio >= 0, TAUV_E_SHAPE, "bad shape B=%d N=%d C1=%d M=%d ratio=%d", B, N,
               C1, M, ratio);
  TAUV_REQUIRE(C1 <= 256, TAUV_E_UNSUPPORTED, "C1=%d exceeds the built-in limit 256", C1);
  TAUV_REQUIRE((uintptr_t)enc % 16 == 0 && (uintptr_t)target % 16 == 0 && (uintptr_t)grad_enc % 16 == 0, TAUV_E_ALIGN,
               "box tensors must be 16-byte aligned");
  const long long rows = (long long)B * N;
  if (grad_cls) {
    const size_t smem = (size_t)kLossTileWarps * 32 * C1 * sizeof(float);
    if (smem > 48 * 1024) TAUV_CUDA(ensure_dynamic_smem((const void*)ycls_backward_kernel, smem));
    ycls_backward_kernel<<<loss_tile_grid(rows), kLossTileWarps * 32, smem, (cudaStream_t)stream>>>(
        cls, rows, C1, N, M, positive, selected, match_index, truth_cls, grad_cls_loss, n_pos_total, ratio, grad_cls);
    TAUV_LAUNCH_CHECK("ycls_backward_kernel");
  }
  if (grad_enc) {
    const int grid = (int)std::min<long long>((rows + 255) / 256, (long long)num_sms() * 8);
    ybox_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)enc, (const float4*)target, positive, rows,
                                                                 grad_box_loss, n_pos_total, (float4*)grad_enc);
    TAUV_LAUNCH_CHECK("ybox_backward_kernel");
  }
  return 0;
}

static int mask_loss_check(const MaskLossArgs& a, int B) {
  TAUV_REQUIRE(a.coeff && a.proto && a.pos_list && a.n_pos && a.match_index && a.truth_box && a.seg && a.img_valid && a.tsum &&
                   a.recs,
               TAUV_E_NULL, "pointers must not be NULL");
  TAUV_REQUIRE((uintptr_t)a.recs % 16 == 0, TAUV_E_ALIGN, "records must be 16-byte aligned");
  TAUV_REQUIRE(a.seg_bytes == 1 || a.seg_bytes == 4 || a.seg_bytes == 8, TAUV_E_UNSUPPORTED, "seg_bytes=%d (1, 4 or 8)", a.seg_bytes);
  TAUV_REQUIRE((uintptr_t)a.seg % a.seg_bytes == 0, TAUV_E_ALIGN, "seg is not aligned to its element size");
  TAUV_REQUIRE(B > 0 && a.N > 0 && a.K > 0 && a.M > 0 && a.PH > 0 && a.PW > 0 && a.SH > 0 && a.SW > 0, TAUV_E_SHAPE,
               "bad shape B=%d N=%d K=%d M=%d proto %dx%d seg %dx%d", B, a.N, a.K, a.M, a.PH, a.PW, a.SH, a.SW);
  TAUV_REQUIRE(a.K <= kMaskLossMaxK, TAUV_E_UNSUPPORTED, "K=%d exceeds the built-in limit %d", a.K, kMaskLossMaxK);
  TAUV_REQUIRE(B <= 65535 && a.M <= 65535, TAUV_E_UNSUPPORTED, "B=%d or M=%d exceeds the built-in limit 65535", B, a.M);
  return 0;
}

extern "C" int tauv_yolact_mask_loss_partials(void) { return kMaskLossWalkers; }
extern "C" size_t tauv_yolact_mask_loss_records_bytes(int B, int N) {
  return B > 0 && N > 0 ? (size_t)B * N * sizeof(MaskRec) : 0;
}

extern "C" int tauv_yolact_mask_loss(const float* coeff, const float* proto, const int32_t* pos_list, const int64_t* n_pos,
                                     const int64_t* match_index, const float* truth_box, const void* seg, int seg_bytes,
                                     const uint8_t* img_valid, int B, int N, int K, int M, int PH, int PW, int SH, int SW,
                                     double* tsum, void* records, double* partial, tauv_stream_t stream) {
  MaskLossArgs a{coeff, proto, pos_list, n_pos, match_index, truth_box, seg, seg_bytes, img_valid, N, K, M, PH, PW, SH, SW,
                 PH > 0 ? (float)SH / (float)PH : 0.f, PW > 0 ? (float)SW / (float)PW : 0.f, tsum, (MaskRec*)records, partial,
                 nullptr, nullptr, nullptr, nullptr};
  if (int rc = mask_loss_check(a, B)) return rc;
  TAUV_REQUIRE(partial, TAUV_E_NULL, "pointers must not be NULL");
  TAUV_REQUIRE((size_t)M * sizeof(unsigned long long) <= 48 * 1024, TAUV_E_UNSUPPORTED, "M=%d exceeds the built-in limit 6144", M);
  TAUV_CUDA(cudaMemsetAsync(tsum, 0, (size_t)B * M * sizeof(double), (cudaStream_t)stream));
  ymask_area_kernel<<<dim3((PH * PW + kMaskLossThreads - 1) / kMaskLossThreads, B), kMaskLossThreads,
                      (size_t)M * sizeof(unsigned long long), (cudaStream_t)stream>>>(a);
  TAUV_LAUNCH_CHECK("ymask_area_kernel");
  ymask_area_finish_kernel<<<(B * M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(tsum, B * M);
  TAUV_LAUNCH_CHECK("ymask_area_finish_kernel");
  ymask_records_kernel<<<dim3((N + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(a);
  TAUV_LAUNCH_CHECK("ymask_records_kernel");
  if (K == kMaskLossMaxK) ymask_positive_kernel<false, true><<<dim3(kMaskLossWalkers, B), kMaskLossThreads, 0, (cudaStream_t)stream>>>(a);
  else ymask_positive_kernel<false, false><<<dim3(kMaskLossWalkers, B), kMaskLossThreads, 0, (cudaStream_t)stream>>>(a);
  TAUV_LAUNCH_CHECK("ymask_positive_kernel<forward>");
  return 0;
}

extern "C" int tauv_yolact_mask_loss_backward(const float* coeff, const float* proto, const int32_t* pos_list,
                                              const int64_t* n_pos, const int64_t* match_index, const float* truth_box,
                                              const void* seg, int seg_bytes, const uint8_t* img_valid, int B, int N, int K, int M,
                                              int PH, int PW, int SH, int SW, const double* tsum,
                                              const void* records, const int64_t* n_pos_total, const float* grad_out, float* grad_coeff,
                                              float* grad_proto, tauv_stream_t stream) {
  MaskLossArgs a{coeff, proto, pos_list, n_pos, match_index, truth_box, seg, seg_bytes, img_valid, N, K, M, PH, PW, SH, SW,
                 PH > 0 ? (float)SH / (float)PH : 0.f, PW > 0 ? (float)SW / (float)PW : 0.f, const_cast<double*>(tsum),
                 (MaskRec*)const_cast<void*>(records), nullptr, grad_out, n_pos_total, grad_coeff, grad_proto};
  if (int rc = mask_loss_check(a, B)) return rc;
  TAUV_REQUIRE(n_pos_total && grad_out, TAUV_E_NULL, "pointers must not be NULL");
  if (grad_coeff) {
    TAUV_CUDA(cudaMemsetAsync(grad_coeff, 0, (size_t)B * N * K * sizeof(float), (cudaStream_t)stream));
    if (K == kMaskLossMaxK) ymask_positive_kernel<true, true><<<dim3(kMaskLossWalkers, B), kMaskLossThreads, 0, (cudaStream_t)stream>>>(a);
    else ymask_positive_kernel<true, false><<<dim3(kMaskLossWalkers, B), kMaskLossThreads, 0, (cudaStream_t)stream>>>(a);
    TAUV_LAUNCH_CHECK("ymask_positive_kernel<backward>");
  }
  if (grad_proto) {
    const dim3 pgrid(((PW + kMaskTileW - 1) / kMaskTileW) * ((PH + kMaskTileH - 1) / kMaskTileH), B);
    if (K == kMaskLossMaxK) ymask_backward_proto_kernel<true><<<pgrid, kMaskLossThreads, 0, (cudaStream_t)stream>>>(a);
    else ymask_backward_proto_kernel<false><<<pgrid, kMaskLossThreads, 0, (cudaStream_t)stream>>>(a);
    TAUV_LAUNCH_CHECK("ymask_backward_proto_kernel");
  }
  return 0;
}

extern "C" int tauv_yolact_loss_reduce(const double* sums, const int64_t* n_pos, int B, int ratio,
                                       const double* mask_partial, int n_partial, float* losses,
                                       int64_t* n_pos_total, tauv_stream_t stream) {
  TAUV_REQUIRE(n_pos && losses && (sums || mask_partial), TAUV_E_NULL, "pointers must not be NULL");
  TAUV_REQUIRE(B > 0 && ratio >= 0 && n_partial >= 0, TAUV_E_SHAPE, "bad shape B=%d ratio=%d n_partial=%d", B, ratio, n_partial);
  yloss_reduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(sums, n_pos, B, ratio, mask_partial, n_partial, losses, n_pos_total);
  TAUV_LAUNCH_CHECK("yloss_reduce_kernel");
  return 0;
}
