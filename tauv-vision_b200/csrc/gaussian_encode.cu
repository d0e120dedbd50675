// gaussian_encode.cu — CenterNet training-target rendering on sm_100a.
//
// Replaces (reference file:line under src/tauv_vision/centernet/model/loss.py):
//   :31-72    generate_heatmap            (B*n_objects Python iterations x 7 ATen launches)
//   :75-135   generate_keypoint_heatmap   (~25 launches per keypoint instance)
//   :138-142  out_index_for_position, :263-264 sub-pixel offset target
// plus gaussian_splat (call sites decode.py:328-332).
//
// Both renderers are pure output-write streams: one CTA per (frame, channel) plane band, the
// frame's objects filtered into shared memory once, then float4 stores.  exp() is monotone, and
// the reference uses one sigma for every object, so max_o exp(-d2_o/s) == exp(-(min_o d2_o)/s):
// one exp per pixel instead of one per (pixel, object), with bit-identical results.
#include "common.cuh"

namespace tauv {

constexpr int kEncThreads = 256;
constexpr int kEncBandElems = 16384;

// cy = floor(center*in / ratio), fp32 multiply then fp32 divide (loss.py:51-52), clamped so the
// integer distance arithmetic cannot overflow (far-away centres render as exact zeros anyway).
__device__ __forceinline__ int grid_floor(float c, float in_sz, float ratio) {
  float v = __fdiv_rn(__fmul_rn(c, in_sz), ratio);
  v = floorf(v);
  v = fminf(fmaxf(v, -1.0e9f), 1.0e9f);
  return (int)v;
}

__device__ __forceinline__ float gauss_from_d2(long long d2, float two_sigma2) {
  // -((x-cx)^2+(y-cy)^2) is an int64 tensor; "/ (2*sigma**2)" promotes to fp32 and divides in fp32
  return expf(__fdiv_rn((float)(-d2), two_sigma2));
}

// One CTA per (frame, group of kEncGroup classes, row band).  The frame's truth is read ONCE per CTA (one latency for
// up to kEncGroup * H * W * 4 bytes of stores instead of one per plane), bucketed by class in shared memory, and then
// every plane of the group is written exactly once: zero planes (four out of five at 16 objects over 80 classes) as
// plain 128-bit zero stores, planes with objects as one exp per pixel.
constexpr int kEncGroup = 8;

template <bool VEC>
__global__ void __launch_bounds__(kEncThreads) gaussian_encode_kernel(
    const uint8_t* __restrict__ valid, const int64_t* __restrict__ label, const float* __restrict__ center,
    int n_objects, int C, int H, int W, float in_h, float in_w, float ratio, float two_sigma2, int groups, int bands,
    int rows_per_band, float* __restrict__ out) {
  extern __shared__ int s_raw[];
  int* s_obj = s_raw;                   // [2*n_objects] (cy, cx), bucketed by class: class g holds [s_start[g], s_start[g+1])
  int* s_tmp = s_raw + 2 * n_objects;   // [3*n_objects] scratch: (class in group or -1, cy, cx) per object
  __shared__ int s_cnt[kEncGroup], s_start[kEncGroup + 1], s_fill[kEncGroup];
  __shared__ int s_big;
  const int tid = threadIdx.x;
  const int band = blockIdx.x % bands;
  const int grp = (blockIdx.x / bands) % groups;
  const long long b = blockIdx.x / ((long long)bands * groups);
  const int c0 = grp * kEncGroup;
  const int nc = min(kEncGroup, C - c0);
  if (tid < kEncGroup) { s_cnt[tid] = 0; s_fill[tid] = 0; }
  if (tid == 0) s_big = 0;
  __syncthreads();
  for (int o = tid; o < n_objects; o += kEncThreads) {
    const long long i = b * n_objects + o;
    int g = -1, cy = 0, cx = 0;
    if (valid[i]) {
      const long long l = label[i] - c0;
      if (l >= 0 && l < nc) {
        g = (int)l;
        cy = grid_floor(center[i * 2 + 0], in_h, ratio);
        cx = grid_floor(center[i * 2 + 1], in_w, ratio);
        atomicAdd(&s_cnt[g], 1);
        if (abs(cy) > 20000 || abs(cx) > 20000) s_big = 1;
      }
    }
    s_tmp[3 * o] = g;
    s_tmp[3 * o + 1] = cy;
    s_tmp[3 * o + 2] = cx;
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int g = 0; g < kEncGroup; ++g) { s_start[g] = acc; acc += s_cnt[g]; }
    s_start[kEncGroup] = acc;
  }
  __syncthreads();
  for (int o = tid; o < n_objects; o += kEncThreads) {
    const int g = s_tmp[3 * o];
    if (g >= 0) {
      const int slot = s_start[g] + atomicAdd(&s_fill[g], 1);  // (the order inside a class does not matter: max)
      s_obj[2 * slot] = s_tmp[3 * o + 1];
      s_obj[2 * slot + 1] = s_tmp[3 * o + 2];
    }
  }
  __syncthreads();
  const bool big = s_big != 0 || H > 10000 || W > 10000;
  const int y0 = band * rows_per_band;
  const int y1 = min(H, y0 + rows_per_band);

  for (int g = 0; g < nc; ++g) {
    const int j0 = s_start[g], j1 = s_start[g + 1];
    float* op = out + (((size_t)b * C + c0 + g) * H + y0) * W;
    auto value = [&](int y, int x) -> float {
      if (!big) {
        int best = 0x7fffffff;
        for (int j = j0; j < j1; ++j) {
          const int dy = y - s_obj[2 * j], dx = x - s_obj[2 * j + 1];
          best = min(best, dy * dy + dx * dx);
        }
        return gauss_from_d2((long long)best, two_sigma2);
      } else {
        long long best = 0x7fffffffffffffffLL;
        for (int j = j0; j < j1; ++j) {
          const long long dy = y - s_obj[2 * j], dx = x - s_obj[2 * j + 1];
          best = min(best, dy * dy + dx * dx);
        }
        return gauss_from_d2(best, two_sigma2);
      }
    };
    if (VEC) {
      const int S = W >> 2;
      const int total = (y1 - y0) * S;
      float4* o4 = reinterpret_cast<float4*>(op);
      if (j0 == j1) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int t = tid; t < total; t += kEncThreads) o4[t] = z;
      } else {
        for (int t = tid; t < total; t += kEncThreads) {
          const int y = y0 + t / S;
          const int x = (t % S) << 2;
          float4 v;
          v.x = value(y, x);
          v.y = value(y, x + 1);
          v.z = value(y, x + 2);
          v.w = value(y, x + 3);
          o4[t] = v;
        }
      }
    } else {
      const int total = (y1 - y0) * W;
      for (int t = tid; t < total; t += kEncThreads) op[t] = (j0 == j1) ? 0.f : value(y0 + t / W, t % W);
    }
  }
}

// Frames with at most 32 objects (the usual case): one WARP per 8 KB chunk of a plane, no shared memory and no block
// barrier.  Lane l reads object l of the frame (the same few lines for every warp of the frame: L1/L2 hits), a
// ballot says which objects render into this plane; if none do, the chunk is sixteen coalesced 128-bit zero stores per
// lane, otherwise the matching centres are broadcast with shuffles.  Warps are laid out in memory order, so the CTAs
// that are resident at any time write one contiguous window of the output, like a plain fill.
#ifndef TAUV_ENC_WARP_STRIPS
#define TAUV_ENC_WARP_STRIPS 256  // (measured, tools/variants_encode.sh: 128 -> 61.5 us, 256 -> 56.4, 512 -> 61.8, 1024 -> 68.6, 4096 -> 129 us per 64 frames)
#endif
constexpr int kEncWarpStrips = TAUV_ENC_WARP_STRIPS;  // 128-bit strips per warp (8 per lane)
constexpr int kKpWarpStrips = 512;  // the keypoint render (three outputs per strip) keeps 16 strips per lane: 244 vs 263 us

__global__ void __launch_bounds__(kEncThreads) gaussian_encode_warp_kernel(
    const uint8_t* __restrict__ valid, const int64_t* __restrict__ label, const float* __restrict__ center,
    int n_objects, int C, int H, int W, float in_h, float in_w, float ratio, float two_sigma2, int chunks_per_plane,
    long long n_chunks, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long chunk_id = (long long)blockIdx.x * (kEncThreads / 32) + (threadIdx.x >> 5);
  if (chunk_id >= n_chunks) return;
  const long long plane = chunk_id / chunks_per_plane;
  const int chunk = (int)(chunk_id - plane * chunks_per_plane);
  const int c = (int)(plane % C);
  const long long b = plane / C;
  bool mine = false;
  int cy = 0, cx = 0;
  if (lane < n_objects) {
    // (all three loads are issued at once: one latency before the stores can start, not three)
    const long long i = b * n_objects + lane;
    const uint8_t v = __ldg(valid + i);
    const long long l = __ldg(label + i);
    const float2 yx = __ldg(reinterpret_cast<const float2*>(center) + i);
    if (v && l == c) {
      mine = true;
      cy = grid_floor(yx.x, in_h, ratio);
      cx = grid_floor(yx.y, in_w, ratio);
    }
  }
  const unsigned mask = __ballot_sync(0xffffffffu, mine);
  const int S = W >> 2;
  const int plane_strips = H * S;
  const int s0 = chunk * kEncWarpStrips;
  const int s1 = min(plane_strips, s0 + kEncWarpStrips);
  float4* o4 = reinterpret_cast<float4*>(out + (size_t)plane * H * W);
  if (mask == 0u) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int t = s0 + lane; t < s1; t += 32) o4[t] = z;
    return;
  }
  const bool big = __any_sync(0xffffffffu, mine && (abs(cy) > 20000 || abs(cx) > 20000)) || H > 10000 || W > 10000;
  for (int t0 = s0; t0 < s1; t0 += 32) {  // (warp-uniform trip count: the shuffles below need every lane)
    const int t = t0 + lane;
    const int y = t / S;
    const int x = (t - y * S) << 2;
    float4 v;
    if (!big) {
      int b0 = 0x7fffffff, b1 = b0, b2 = b0, b3 = b0;
      for (unsigned m = mask; m; m &= m - 1) {
        const int j = __ffs(m) - 1;
        const int dy = y - __shfl_sync(0xffffffffu, cy, j), dx = x - __shfl_sync(0xffffffffu, cx, j);
        const int dy2 = dy * dy;
        b0 = min(b0, dy2 + dx * dx);
        b1 = min(b1, dy2 + (dx + 1) * (dx + 1));
        b2 = min(b2, dy2 + (dx + 2) * (dx + 2));
        b3 = min(b3, dy2 + (dx + 3) * (dx + 3));
      }
      v = make_float4(gauss_from_d2(b0, two_sigma2), gauss_from_d2(b1, two_sigma2), gauss_from_d2(b2, two_sigma2),
                      gauss_from_d2(b3, two_sigma2));
    } else {
      long long b0 = 0x7fffffffffffffffLL, b1 = b0, b2 = b0, b3 = b0;
      for (unsigned m = mask; m; m &= m - 1) {
        const int j = __ffs(m) - 1;
        const long long dy = y - __shfl_sync(0xffffffffu, cy, j), dx = x - __shfl_sync(0xffffffffu, cx, j);
        const long long dy2 = dy * dy;
        b0 = min(b0, dy2 + dx * dx);
        b1 = min(b1, dy2 + (dx + 1) * (dx + 1));
        b2 = min(b2, dy2 + (dx + 2) * (dx + 2));
        b3 = min(b3, dy2 + (dx + 3) * (dx + 3));
      }
      v = make_float4(gauss_from_d2(b0, two_sigma2), gauss_from_d2(b1, two_sigma2), gauss_from_d2(b2, two_sigma2),
                      gauss_from_d2(b3, two_sigma2));
    }
    if (t < s1) o4[t] = v;
  }
}

// nan_to_num(x, nan) semantics (loss.py:116-118, :131-133): NaN -> nan, +-inf -> +-FLT_MAX
__device__ __forceinline__ float nan_to_num(float v, float nan) {
  if (v != v) return nan;
  if (v == __int_as_float(0x7f800000)) return 3.4028234663852886e38f;
  if (v == __int_as_float(0xff800000)) return -3.4028234663852886e38f;
  return v;
}

template <bool VEC>
__global__ void __launch_bounds__(kEncThreads) keypoint_encode_kernel(
    const uint8_t* __restrict__ kp_valid, const int64_t* __restrict__ kp_label, const float* __restrict__ kp_center,
    const int64_t* __restrict__ kp_obj, const float* __restrict__ center, int m, int n_objects, int Kp, int H, int W,
    float in_h, float in_w, float ratio, float two_sh2, float two_sa2, int bands, int rows_per_band,
    float* __restrict__ heatmap, float* __restrict__ weight, float* __restrict__ affinity) {
  // per matching instance, in instance order: (cy, cx) ints + owning object's centre (fp32 y, x)
  extern __shared__ int s_raw[];
  int* s_c = s_raw;                                        // [2*m]
  float* s_o = reinterpret_cast<float*>(s_raw + 2 * m);    // [2*m]
  __shared__ int s_n, s_big;
  const int tid = threadIdx.x;
  const int band = blockIdx.x % bands;
  const long long plane = blockIdx.x / bands;
  const int k = (int)(plane % Kp);
  const long long b = plane / Kp;
  if (tid == 0) { s_n = 0; s_big = 0; }
  __syncthreads();
  if (tid < 32) {  // ordered compaction by one warp: instance order decides affinity ties (loss.py:122)
    int base = 0;
    for (int start = 0; start < m; start += 32) {
      const int i = start + tid;
      bool hit = false;
      long long gi = 0;
      if (i < m) {
        gi = b * m + i;
        hit = kp_valid[gi] && kp_label[gi] == k;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int slot = base + __popc(bal & ((1u << tid) - 1u));
        const int cy = grid_floor(kp_center[gi * 2 + 0], in_h, ratio);
        const int cx = grid_floor(kp_center[gi * 2 + 1], in_w, ratio);
        s_c[2 * slot] = cy;
        s_c[2 * slot + 1] = cx;
        long long oi = kp_obj[gi];
        if (oi < 0) oi += n_objects;  // torch negative indexing
        oi = max(0LL, min((long long)n_objects - 1, oi));
        s_o[2 * slot] = center[(b * n_objects + oi) * 2 + 0];
        s_o[2 * slot + 1] = center[(b * n_objects + oi) * 2 + 1];
        if (abs(cy) > 20000 || abs(cx) > 20000) s_big = 1;
      }
      base += __popc(bal);
    }
    if (tid == 0) s_n = base;
  }
  __syncthreads();
  const int n = s_n;
  const bool big = s_big != 0 || H > 10000 || W > 10000;
  const int y0 = band * rows_per_band;
  const int y1 = min(H, y0 + rows_per_band);
  const size_t hw = (size_t)H * W;
  float* hp = heatmap + (size_t)plane * hw;
  float* wp = weight + (size_t)plane * hw;
  float* ay = affinity + (size_t)plane * 2 * hw;
  float* ax = ay + hw;
  const float fh = (float)H, fw = (float)W;

  auto eval = [&](int y, int x, float& vh, float& vw, float& vay, float& vax) {
    long long best = 0x7fffffffffffffffLL;
    if (!big) {
      int bi = 0x7fffffff;
      for (int j = 0; j < n; ++j) {
        const int dy = y - s_c[2 * j], dx = x - s_c[2 * j + 1];
        bi = min(bi, dy * dy + dx * dx);
      }
      best = bi;
    } else {
      for (int j = 0; j < n; ++j) {
        const long long dy = y - s_c[2 * j], dx = x - s_c[2 * j + 1];
        best = min(best, dy * dy + dx * dx);
      }
    }
    vh = gauss_from_d2(best, two_sh2);
    vw = gauss_from_d2(best, two_sa2);
    // affinity: unit vector from the owning object's centre, nearest object wins, first on ties
    const float py = __fdiv_rn((float)y, fh), px = __fdiv_rn((float)x, fw);  // loss.py:114
    float cur = __int_as_float(0x7f800000);
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < n; ++j) {
      const float dy = nan_to_num(__fsub_rn(py, s_o[2 * j]), 0.f);
      const float dx = nan_to_num(__fsub_rn(px, s_o[2 * j + 1]), 0.f);
      const float dist = nan_to_num(__fsqrt_rn(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx))), 1.f);
      if (dist < cur) {
        a0 = __fdiv_rn(dy, dist);
        a1 = __fdiv_rn(dx, dist);
        cur = dist;
      }
    }
    vay = nan_to_num(a0, 0.f);
    vax = nan_to_num(a1, 0.f);
  };

  if (VEC) {
    const int S = W >> 2;
    const int total = (y1 - y0) * S;
    const size_t o0 = (size_t)y0 * W;
    float4* h4 = reinterpret_cast<float4*>(hp + o0);
    float4* w4 = reinterpret_cast<float4*>(wp + o0);
    float4* y4 = reinterpret_cast<float4*>(ay + o0);
    float4* x4 = reinterpret_cast<float4*>(ax + o0);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = tid; t < total; t += kEncThreads) {
      if (n == 0) {
        h4[t] = z; w4[t] = z; y4[t] = z; x4[t] = z;
        continue;
      }
      const int y = y0 + t / S;
      const int x = (t % S) << 2;
      float4 vh, vw, vy, vx;
      eval(y, x, vh.x, vw.x, vy.x, vx.x);
      eval(y, x + 1, vh.y, vw.y, vy.y, vx.y);
      eval(y, x + 2, vh.z, vw.z, vy.z, vx.z);
      eval(y, x + 3, vh.w, vw.w, vy.w, vx.w);
      h4[t] = vh; w4[t] = vw; y4[t] = vy; x4[t] = vx;
    }
  } else {
    const int total = (y1 - y0) * W;
    const size_t o0 = (size_t)y0 * W;
    for (int t = tid; t < total; t += kEncThreads) {
      float vh = 0.f, vw = 0.f, vy = 0.f, vx = 0.f;
      if (n) eval(y0 + t / W, t % W, vh, vw, vy, vx);
      hp[o0 + t] = vh; wp[o0 + t] = vw; ay[o0 + t] = vy; ax[o0 + t] = vx;
    }
  }
}

// Frames with at most 32*NCH keypoint instances: one WARP per 8 KB chunk of a keypoint plane, like
// gaussian_encode_warp_kernel.  Lane l holds instances l, l+32, ... (chunk-major ballots walked in ascending bit order
// ARE the instance order that decides affinity ties, loss.py:122); a plane without instances is four runs of coalesced zero stores (heatmap, weight,
// affinity y and x), the others broadcast the matching instances with shuffles.  No shared memory, no block barrier.
// MODE 0 renders the targets.  MODE 1 / 2 are the keypoint-affinity term of the loss fused with that render
// (loss.py:244-246: (affinity_weight.unsqueeze(2) * F.mse_loss(prediction.keypoint_affinity, keypoint_affinity,
// reduction="none")).sum(), before its lambda): MODE 1 reads the predicted field `pred` [B,Kp,2,H,W] and leaves one fp64
// partial sum per warp chunk in `partial` (planes without instances have weight 0 and are not read at all); MODE 2
// writes the gradient 2 g w (pred - target) into `affinity` (g = *grad_out).  No target is written in either.
template <int NCH, int MODE>
__global__ void __launch_bounds__(kEncThreads) keypoint_encode_warp_kernel(
    const uint8_t* __restrict__ kp_valid, const int64_t* __restrict__ kp_label, const float* __restrict__ kp_center,
    const int64_t* __restrict__ kp_obj, const float* __restrict__ center, int m, int n_objects, int Kp, int H, int W,
    float in_h, float in_w, float ratio, float two_sh2, float two_sa2, int chunks_per_plane, long long n_chunks,
    float* __restrict__ heatmap, float* __restrict__ weight, float* __restrict__ affinity,
    const float* __restrict__ pred, double* __restrict__ partial, const float* __restrict__ grad_out) {
  const int lane = threadIdx.x & 31;
  const long long chunk_id = (long long)blockIdx.x * (kEncThreads / 32) + (threadIdx.x >> 5);
  if (chunk_id >= n_chunks) return;
  const long long plane = chunk_id / chunks_per_plane;
  const int chunk = (int)(chunk_id - plane * chunks_per_plane);
  const int k = (int)(plane % Kp);
  const long long b = plane / Kp;
  int cy[NCH], cx[NCH];
  float oy[NCH], ox[NCH];
  unsigned mask[NCH];
  unsigned any_mask = 0u;
  bool far = false;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    bool mine = false;
    cy[c] = cx[c] = 0;
    oy[c] = ox[c] = 0.f;
    const int inst = lane + 32 * c;
    if (inst < m) {
      const long long gi = b * m + inst;
      const uint8_t v = __ldg(kp_valid + gi);
      const long long l = __ldg(kp_label + gi);
      const float c0 = __ldg(kp_center + gi * 2), c1 = __ldg(kp_center + gi * 2 + 1);
      long long oi = __ldg(kp_obj + gi);
      if (v && l == k) {
        mine = true;
        cy[c] = grid_floor(c0, in_h, ratio);
        cx[c] = grid_floor(c1, in_w, ratio);
        if (oi < 0) oi += n_objects;  // torch negative indexing
        oi = max(0LL, min((long long)n_objects - 1, oi));
        oy[c] = __ldg(center + (b * n_objects + oi) * 2 + 0);
        ox[c] = __ldg(center + (b * n_objects + oi) * 2 + 1);
        far |= abs(cy[c]) > 20000 || abs(cx[c]) > 20000;
      }
    }
    mask[c] = __ballot_sync(0xffffffffu, mine);
    any_mask |= mask[c];
  }
  const int S = W >> 2;
  const int plane_strips = H * S;
  const int s0 = chunk * kKpWarpStrips;
  const int s1 = min(plane_strips, s0 + kKpWarpStrips);
  const size_t hw = (size_t)H * W;
  float4* h4 = reinterpret_cast<float4*>(heatmap + (size_t)plane * hw);
  float4* w4 = reinterpret_cast<float4*>(weight + (size_t)plane * hw);
  float4* y4 = reinterpret_cast<float4*>(affinity + (size_t)plane * 2 * hw);
  float4* x4 = reinterpret_cast<float4*>(affinity + (size_t)plane * 2 * hw + hw);
  const float4* py4 = reinterpret_cast<const float4*>(pred + (MODE ? (size_t)plane * 2 * hw : 0));
  const float4* px4 = reinterpret_cast<const float4*>(pred + (MODE ? (size_t)plane * 2 * hw + hw : 0));
  if (any_mask == 0u) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == 1) {
      if (lane == 0) partial[chunk_id] = 0.0;
      return;
    }
#pragma unroll 4
    for (int t = s0 + lane; t < s1; t += 32) {
      if (MODE == 0) { h4[t] = z; w4[t] = z; }
      y4[t] = z; x4[t] = z;
    }
    return;
  }
  const bool big = __any_sync(0xffffffffu, far) || H > 10000 || W > 10000;
  const float fh = (float)H, fw = (float)W;
  double acc = 0.0;
  const float g2 = MODE == 2 ? __fmul_rn(2.0f, *grad_out) : 0.0f;
  for (int t0 = s0; t0 < s1; t0 += 32) {  // (warp-uniform trip count: the shuffles below need every lane)
    const int t = t0 + lane;
    const int y = t / S;
    const int xb = (t - y * S) << 2;
    const float py = __fdiv_rn((float)y, fh);  // loss.py:114
    long long best[4] = {0x7fffffffffffffffLL, 0x7fffffffffffffffLL, 0x7fffffffffffffffLL, 0x7fffffffffffffffLL};
    float cur[4], a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f}, px[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      cur[i] = __int_as_float(0x7f800000);
      px[i] = __fdiv_rn((float)(xb + i), fw);
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c)
    for (unsigned mm = mask[c]; mm; mm &= mm - 1) {  // chunk-major, ascending lane = instance order
      const int j = __ffs(mm) - 1;
      const int jcy = __shfl_sync(0xffffffffu, cy[c], j), jcx = __shfl_sync(0xffffffffu, cx[c], j);
      const float joy = __shfl_sync(0xffffffffu, oy[c], j), jox = __shfl_sync(0xffffffffu, ox[c], j);
      const float dyf = nan_to_num(__fsub_rn(py, joy), 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!big) {
          const int dy = y - jcy, dx = xb + i - jcx;
          best[i] = min(best[i], (long long)(dy * dy + dx * dx));
        } else {
          const long long dy = y - jcy, dx = xb + i - jcx;
          best[i] = min(best[i], dy * dy + dx * dx);
        }
        // affinity: unit vector from the owning object's centre, nearest object wins, first on ties
        const float dxf = nan_to_num(__fsub_rn(px[i], jox), 0.f);
        const float dist = nan_to_num(__fsqrt_rn(__fadd_rn(__fmul_rn(dyf, dyf), __fmul_rn(dxf, dxf))), 1.f);
        if (dist < cur[i]) {
          a0[i] = __fdiv_rn(dyf, dist);
          a1[i] = __fdiv_rn(dxf, dist);
          cur[i] = dist;
        }
      }
    }
    if (t < s1) {
      const float4 wv = make_float4(gauss_from_d2(best[0], two_sa2), gauss_from_d2(best[1], two_sa2),
                                    gauss_from_d2(best[2], two_sa2), gauss_from_d2(best[3], two_sa2));
      const float4 ty = make_float4(nan_to_num(a0[0], 0.f), nan_to_num(a0[1], 0.f), nan_to_num(a0[2], 0.f), nan_to_num(a0[3], 0.f));
      const float4 tx = make_float4(nan_to_num(a1[0], 0.f), nan_to_num(a1[1], 0.f), nan_to_num(a1[2], 0.f), nan_to_num(a1[3], 0.f));
      if (MODE == 0) {
        h4[t] = make_float4(gauss_from_d2(best[0], two_sh2), gauss_from_d2(best[1], two_sh2), gauss_from_d2(best[2], two_sh2),
                            gauss_from_d2(best[3], two_sh2));
        w4[t] = wv;
        y4[t] = ty;
        x4[t] = tx;
      } else {
        const float4 qy = py4[t], qx = px4[t];
        const float dy[4] = {qy.x - ty.x, qy.y - ty.y, qy.z - ty.z, qy.w - ty.w};
        const float dx[4] = {qx.x - tx.x, qx.y - tx.y, qx.z - tx.z, qx.w - tx.w};
        const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
        if (MODE == 1) {
#pragma unroll
          for (int i = 0; i < 4; ++i) acc += (double)(ww[i] * (dy[i] * dy[i])) + (double)(ww[i] * (dx[i] * dx[i]));
        } else {
          y4[t] = make_float4(g2 * ww[0] * dy[0], g2 * ww[1] * dy[1], g2 * ww[2] * dy[2], g2 * ww[3] * dy[3]);
          x4[t] = make_float4(g2 * ww[0] * dx[0], g2 * ww[1] * dx[1], g2 * ww[2] * dx[2], g2 * ww[3] * dx[3]);
        }
      }
    }
  }
  if (MODE == 1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) partial[chunk_id] = acc;
  }
}

__global__ void out_index_offset_kernel(const float* __restrict__ pos, long long n, float in_h, float in_w,
                                        float ratio, long long iratio, int out_h, int out_w,
                                        int64_t* __restrict__ index, float* __restrict__ offset) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= 2 * n) return;
  const int ax = (int)(t & 1);
  const float in_sz = ax ? in_w : in_h;
  const int out_sz = ax ? out_w : out_h;
  const float pix = __fmul_rn(pos[t], in_sz);
  // .to(torch.long) truncates toward zero; out-of-range / NaN follow the x86 cvttss2si convention
  const float q = __fdiv_rn(pix, ratio);
  long long cell;
  if (!(q > -9.2e18f && q < 9.2e18f)) cell = (long long)0x8000000000000000ULL;
  else cell = (long long)q;
  long long ci = cell < 0 ? 0 : (cell > out_sz - 1 ? out_sz - 1 : cell);
  index[t] = ci;
  if (offset) offset[t] = __fsub_rn(pix, (float)(iratio * cell));  // loss.py:263-264 (unclamped cell)
}

__global__ void gaussian_splat_kernel(int h, int w, int cy, int cx, float two_sigma2, float* __restrict__ out) {
  const long long n = (long long)h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long y = i / w, x = i % w;
    const long long dy = y - cy, dx = x - cx;
    out[i] = gauss_from_d2(dy * dy + dx * dx, two_sigma2);
  }
}

static void band_plan(int H, int W, long long planes, int* bands, int* rows) {
  int r = kEncBandElems / W;
  if (r < 1) r = 1;
  if (r > H) r = H;
  // enough CTAs to fill the machine a few times over even for tiny batches
  const int sms = num_sms();
  while (r > 1 && planes * ((H + r - 1) / r) < 4LL * sms) r = (r + 1) / 2;
  *rows = r;
  *bands = (H + r - 1) / r;
}


// ----------------------------------------------------------------------------------------------
// Heatmap focal loss fused with the target render (loss.py:182 + :233-236 + :302-317):
//   loss = focal_loss(sigmoid(logits), generate_heatmap(truth)).sum()
// The 335 MB target is never written: every warp renders its chunk of the target in registers (exactly as
// gaussian_encode_warp_kernel does), reads the same chunk of the logits once, and reduces.  The reference makes ~20
// elementwise passes over [B,C,H,W] tensors for this term (sigmoid, isclose, two pows, two clamps, two logs, masks,
// the division by N, the sum) plus the target write.  Forward: per-chunk partial sums (double) and positive counts, summed
// per frame in a fixed order by a second tiny launch (deterministic).  Backward: the gradient with respect to the logits
// in one more pass (the target is rendered again rather than stored).
// ----------------------------------------------------------------------------------------------
struct FocalPow {
  float e;
  int kind;  // 2: x*x, 3: x*x*x (ATen's special cases of pow(tensor, scalar)), 0: powf
};
__device__ __forceinline__ float focal_pow(float x, FocalPow p) {
  if (p.kind == 2) return __fmul_rn(x, x);
  if (p.kind == 3) return __fmul_rn(__fmul_rn(x, x), x);
  return powf(x, p.e);
}
// the same with the kind known at compile time (alpha is applied to every cell: the run-time tests were a tenth of the
// forward kernel's instructions)
template <int KIND>
__device__ __forceinline__ float focal_pow_k(float x, FocalPow p) {
  if (KIND == 2) return __fmul_rn(x, x);
  if (KIND == 3) return __fmul_rn(__fmul_rn(x, x), x);
  return powf(x, p.e);
}
template <int KIND>
__device__ __forceinline__ float focal_dpow_k(float x, FocalPow p) {
  if (KIND == 2) return __fmul_rn(2.0f, x);
  if (KIND == 3) return __fmul_rn(3.0f, __fmul_rn(x, x));
  return __fmul_rn(p.e, powf(x, p.e - 1.0f));
}
// d/dx x^e (autograd: e * x^(e-1))
__device__ __forceinline__ float focal_dpow(float x, FocalPow p) {
  if (p.kind == 2) return __fmul_rn(2.0f, x);
  if (p.kind == 3) return __fmul_rn(3.0f, __fmul_rn(x, x));
  return __fmul_rn(p.e, powf(x, p.e - 1.0f));
}
// torch.isclose(t, 1): |t - 1| <= atol + rtol * |1| with the defaults, evaluated in fp32 like ATen does
__device__ __forceinline__ bool focal_is_pos(float t) {
  const float allowed = __fadd_rn(1e-8f, __fmul_rn(1e-5f, 1.0f));
  const float err = fabsf(__fsub_rn(t, 1.0f));
  return err <= allowed;  // (false for NaN / inf, like isfinite(actual_error) & ...)
}

// the four target values of strip t of the plane (see gaussian_encode_warp_kernel); mask == 0: zeros
__device__ __forceinline__ float4 focal_target4(unsigned mask, int cy, int cx, int y, int x, float two_sigma2, bool big) {
  if (mask == 0u) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (!big) {
    int b0 = 0x7fffffff, b1 = b0, b2 = b0, b3 = b0;
    for (unsigned m = mask; m; m &= m - 1) {
      const int j = __ffs(m) - 1;
      const int dy = y - __shfl_sync(0xffffffffu, cy, j), dx = x - __shfl_sync(0xffffffffu, cx, j);
      const int dy2 = dy * dy;
      b0 = min(b0, dy2 + dx * dx);
      b1 = min(b1, dy2 + (dx + 1) * (dx + 1));
      b2 = min(b2, dy2 + (dx + 2) * (dx + 2));
      b3 = min(b3, dy2 + (dx + 3) * (dx + 3));
    }
    return make_float4(gauss_from_d2(b0, two_sigma2), gauss_from_d2(b1, two_sigma2), gauss_from_d2(b2, two_sigma2),
                       gauss_from_d2(b3, two_sigma2));
  }
  long long b0 = 0x7fffffffffffffffLL, b1 = b0, b2 = b0, b3 = b0;
  for (unsigned m = mask; m; m &= m - 1) {
    const int j = __ffs(m) - 1;
    const long long dy = y - __shfl_sync(0xffffffffu, cy, j), dx = x - __shfl_sync(0xffffffffu, cx, j);
    const long long dy2 = dy * dy;
    b0 = min(b0, dy2 + dx * dx);
    b1 = min(b1, dy2 + (dx + 1) * (dx + 1));
    b2 = min(b2, dy2 + (dx + 2) * (dx + 2));
    b3 = min(b3, dy2 + (dx + 3) * (dx + 3));
  }
  return make_float4(gauss_from_d2(b0, two_sigma2), gauss_from_d2(b1, two_sigma2), gauss_from_d2(b2, two_sigma2),
                     gauss_from_d2(b3, two_sigma2));
}

struct FocalArgs {
  const float* logits;
  const uint8_t* valid;
  const int64_t* label;
  const float* center;
  int n_objects, C, H, W;
  float in_h, in_w, ratio, two_sigma2;
  FocalPow a, b;
  int chunks_per_plane;
  long long n_chunks;
};

// this warp's chunk: which objects render into its plane (ballot), and their grid cells
__device__ __forceinline__ unsigned focal_chunk_objects(const FocalArgs& g, long long plane, int lane, int& cy, int& cx,
                                                        bool& big) {
  const int c = (int)(plane % g.C);
  const long long b = plane / g.C;
  bool mine = false;
  cy = 0;
  cx = 0;
  if (lane < g.n_objects) {
    const long long i = b * g.n_objects + lane;
    const uint8_t v = __ldg(g.valid + i);
    const long long l = __ldg(g.label + i);
    const float2 yx = __ldg(reinterpret_cast<const float2*>(g.center) + i);
    if (v && l == c) {
      mine = true;
      cy = grid_floor(yx.x, g.in_h, g.ratio);
      cx = grid_floor(yx.y, g.in_w, g.ratio);
    }
  }
  const unsigned mask = __ballot_sync(0xffffffffu, mine);
  big = __any_sync(0xffffffffu, mine && (abs(cy) > 20000 || abs(cx) > 20000)) || g.H > 10000 || g.W > 10000;
  return mask;
}

template <int KA>
__global__ void __launch_bounds__(kEncThreads) focal_forward_kernel(const __grid_constant__ FocalArgs g,
                                                                    double* __restrict__ part /*[n_chunks][2]*/,
                                                                    int* __restrict__ part_pos /*[n_chunks]*/) {
  const int lane = threadIdx.x & 31;
  const long long chunk_id = (long long)blockIdx.x * (kEncThreads / 32) + (threadIdx.x >> 5);
  if (chunk_id >= g.n_chunks) return;
  const long long plane = chunk_id / g.chunks_per_plane;
  const int chunk = (int)(chunk_id - plane * g.chunks_per_plane);
  int cy, cx;
  bool big;
  const unsigned mask = focal_chunk_objects(g, plane, lane, cy, cx, big);
  const int S = g.W >> 2;
  const int plane_strips = g.H * S;
  const int s0 = chunk * kEncWarpStrips;
  const int s1 = min(plane_strips, s0 + kEncWarpStrips);
  const float4* x4 = reinterpret_cast<const float4*>(g.logits + (size_t)plane * g.H * g.W);
  double acc_p = 0.0, acc_n = 0.0;
  int n_pos = 0;
  for (int t0 = s0; t0 < s1; t0 += 32) {  // (warp-uniform trip count: the shuffles of the render need every lane)
    const int t = t0 + lane;
    const int y = t / S;
    const int x = (t - y * S) << 2;
    const float4 tv = focal_target4(mask, cy, cx, y, x, g.two_sigma2, big);
    if (t < s1) {
      const float4 xv = ldg_stream4(reinterpret_cast<const float*>(x4 + t));
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ts[4] = {tv.x, tv.y, tv.z, tv.w};
      float sp = 0.0f, sn = 0.0f;  // the strip's four terms in fp32, one fp64 add per strip and sum
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p = sigmoid_ref(xs[i]);
        float tt = ts[i];
        if (tt != tt) tt = 0.0f;  // (generate_heatmap ends with nan_to_num)
        if (mask != 0u && focal_is_pos(tt)) {
          sp += __fmul_rn(focal_pow_k<KA>(__fsub_rn(1.0f, p), g.a), logf(fmaxf(p, 1e-4f)));
          ++n_pos;
        } else {
          float w = focal_pow_k<KA>(p, g.a);
          if (mask != 0u) w = __fmul_rn(focal_pow(__fsub_rn(1.0f, tt), g.b), w);
          sn += __fmul_rn(w, logf(fmaxf(__fsub_rn(1.0f, p), 1e-4f)));
        }
      }
      acc_p += (double)sp;
      acc_n += (double)sn;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_p += __shfl_xor_sync(0xffffffffu, acc_p, o);
    acc_n += __shfl_xor_sync(0xffffffffu, acc_n, o);
    n_pos += __shfl_xor_sync(0xffffffffu, n_pos, o);
  }
  if (lane == 0) {
    part[chunk_id * 2 + 0] = acc_p;
    part[chunk_id * 2 + 1] = acc_n;
    part_pos[chunk_id] = n_pos;
  }
}

// one warp per frame: the frame's chunk partials in a fixed order (lane-strided, then a butterfly: deterministic)
__global__ void __launch_bounds__(32) focal_reduce_kernel(const double* __restrict__ part, const int* __restrict__ part_pos,
                                                          long long chunks_per_frame, double* __restrict__ frame_sums,
                                                          int64_t* __restrict__ frame_pos) {
  const long long b = blockIdx.x;
  const int lane = threadIdx.x;
  double sp = 0.0, sn = 0.0;
  long long np = 0;
  for (long long i = lane; i < chunks_per_frame; i += 32) {
    sp += part[(b * chunks_per_frame + i) * 2 + 0];
    sn += part[(b * chunks_per_frame + i) * 2 + 1];
    np += part_pos[b * chunks_per_frame + i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sp += __shfl_xor_sync(0xffffffffu, sp, o);
    sn += __shfl_xor_sync(0xffffffffu, sn, o);
    np += __shfl_xor_sync(0xffffffffu, np, o);
  }
  if (lane == 0) {
    frame_sums[b * 2 + 0] = sp;
    frame_sums[b * 2 + 1] = sn;
    frame_pos[b] = np;
  }
}

// the batch's loss from the frames' sums (loss.py:313-317 summed): N = all positive cells; -(sp + sn) / N, or -sp when
// N == 0.  One warp, fixed order.
__global__ void __launch_bounds__(32) focal_finish_kernel(const double* __restrict__ frame_sums,
                                                          const int64_t* __restrict__ frame_pos, int B,
                                                          float* __restrict__ loss, int64_t* __restrict__ n_pos_total) {
  const int lane = threadIdx.x;
  double sp = 0.0, sn = 0.0;
  long long np = 0;
  for (int i = lane; i < B; i += 32) {
    sp += frame_sums[2 * i];
    sn += frame_sums[2 * i + 1];
    np += frame_pos[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sp += __shfl_xor_sync(0xffffffffu, sp, o);
    sn += __shfl_xor_sync(0xffffffffu, sn, o);
    np += __shfl_xor_sync(0xffffffffu, np, o);
  }
  if (lane == 0) {
    *loss = (float)(np > 0 ? -(sp + sn) / (double)np : -sp);
    *n_pos_total = np;
  }
}

// d(sum of the loss)/d(logits) * grad_out.  loss = -(loss_p + loss_n) / N for N > 0, -loss_p for N == 0.
template <int KA>
__global__ void __launch_bounds__(kEncThreads) focal_backward_kernel(const __grid_constant__ FocalArgs g,
                                                                     const int64_t* __restrict__ n_pos_total,
                                                                     const float* __restrict__ grad_out,
                                                                     float* __restrict__ grad) {
  const int lane = threadIdx.x & 31;
  const long long chunk_id = (long long)blockIdx.x * (kEncThreads / 32) + (threadIdx.x >> 5);
  if (chunk_id >= g.n_chunks) return;
  const long long plane = chunk_id / g.chunks_per_plane;
  const int chunk = (int)(chunk_id - plane * g.chunks_per_plane);
  int cy, cx;
  bool big;
  const unsigned mask = focal_chunk_objects(g, plane, lane, cy, cx, big);
  const long long N = __ldg(n_pos_total);
  // autograd: d(-(x)/N) = -1/N (an fp32 division of the incoming gradient by N), or -1 when the negatives are dropped
  const float go = __ldg(grad_out);
  const float scale = N > 0 ? -__fdiv_rn(go, (float)N) : -go;
  const bool with_neg = N > 0;
  const int S = g.W >> 2;
  const int plane_strips = g.H * S;
  const int s0 = chunk * kEncWarpStrips;
  const int s1 = min(plane_strips, s0 + kEncWarpStrips);
  const float4* x4 = reinterpret_cast<const float4*>(g.logits + (size_t)plane * g.H * g.W);
  float4* g4 = reinterpret_cast<float4*>(grad + (size_t)plane * g.H * g.W);
  for (int t0 = s0; t0 < s1; t0 += 32) {
    const int t = t0 + lane;
    const int y = t / S;
    const int x = (t - y * S) << 2;
    const float4 tv = focal_target4(mask, cy, cx, y, x, g.two_sigma2, big);
    if (t < s1) {
      const float4 xv = ldg_stream4(reinterpret_cast<const float*>(x4 + t));
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ts[4] = {tv.x, tv.y, tv.z, tv.w};
      float gs[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p = sigmoid_ref(xs[i]);
        const float q = __fsub_rn(1.0f, p);
        float tt = ts[i];
        if (tt != tt) tt = 0.0f;
        float dldp;  // d(loss_p + loss_n)/dp
        if (mask != 0u && focal_is_pos(tt)) {
          // (1-p)^a log(clamp(p)):  -a (1-p)^(a-1) log(c) + (1-p)^a [p >= 1e-4] / p
          const float lg = logf(fmaxf(p, 1e-4f));
          dldp = -focal_dpow_k<KA>(q, g.a) * lg + (p >= 1e-4f ? focal_pow_k<KA>(q, g.a) / p : 0.0f);
        } else if (with_neg) {
          // w p^a log(clamp(1-p)),  w = (1-t)^b:  w [a p^(a-1) log(c) - p^a [1-p >= 1e-4] / (1-p)]
          const float w = mask != 0u ? focal_pow(__fsub_rn(1.0f, tt), g.b) : 1.0f;
          const float lg = logf(fmaxf(q, 1e-4f));
          dldp = w * (focal_dpow_k<KA>(p, g.a) * lg - (q >= 1e-4f ? focal_pow_k<KA>(p, g.a) / q : 0.0f));
        } else {
          dldp = 0.0f;
        }
        gs[i] = scale * dldp * (p * q);  // sigmoid backward: p (1 - p)
      }
      g4[t] = make_float4(gs[0], gs[1], gs[2], gs[3]);
    }
  }
}

static float two_sigma_sq(double sigma) {
  if (sigma < 0.1) sigma = 0.1;  // loss.py:60-62 ("tiny sigma!")
  return (float)(2.0 * (sigma * sigma));
}

static FocalPow make_focal_pow(double e) {
  FocalPow p;
  p.e = (float)e;
  p.kind = e == 2.0 ? 2 : e == 3.0 ? 3 : 0;
  return p;
}

static int focal_args(FocalArgs* g, const float* logits, const uint8_t* valid, const int64_t* label, const float* center,
                      int B, int n_objects, int C, int H, int W, int in_h, int in_w, int downsample_ratio, double sigma,
                      double alpha, double beta) {
  TAUV_REQUIRE(logits, TAUV_E_NULL, "logits must not be NULL");
  TAUV_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && n_objects >= 0, TAUV_E_SHAPE, "bad shape");
  TAUV_REQUIRE(n_objects == 0 || (valid && label && center), TAUV_E_NULL, "valid/label/center must not be NULL");
  TAUV_REQUIRE(in_h > 0 && in_w > 0 && downsample_ratio > 0, TAUV_E_SHAPE, "bad model geometry");
  TAUV_REQUIRE(W % 4 == 0 && (uintptr_t)logits % 16 == 0 && n_objects <= 32 && (uintptr_t)center % 8 == 0,
               TAUV_E_UNSUPPORTED, "the fused focal loss needs W %% 4 == 0, 16-byte aligned logits and <= 32 objects per frame");
  g->logits = logits; g->valid = valid; g->label = label; g->center = center;
  g->n_objects = n_objects; g->C = C; g->H = H; g->W = W;
  g->in_h = (float)in_h; g->in_w = (float)in_w; g->ratio = (float)downsample_ratio; g->two_sigma2 = two_sigma_sq(sigma);
  g->a = make_focal_pow(alpha); g->b = make_focal_pow(beta);
  const long long plane_strips = (long long)H * (W / 4);
  const long long cpp = (plane_strips + kEncWarpStrips - 1) / kEncWarpStrips;
  g->chunks_per_plane = (int)cpp;
  g->n_chunks = (long long)B * C * cpp;
  TAUV_REQUIRE(cpp < (1LL << 31) && (g->n_chunks + kEncThreads / 32 - 1) / (kEncThreads / 32) < (1LL << 31),
               TAUV_E_UNSUPPORTED, "grid too large");
  return 0;
}

}  // namespace tauv

using namespace tauv;


extern "C" size_t tauv_centernet_focal_loss_workspace_bytes(int B, int C, int H, int W) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || W % 4 != 0) return 0;
  const long long plane_strips = (long long)H * (W / 4);
  const long long cpp = (plane_strips + kEncWarpStrips - 1) / kEncWarpStrips;
  const size_t n_chunks = (size_t)B * C * cpp;
  return align_up(n_chunks * 16, 256) + align_up(n_chunks * 4, 256);
}

extern "C" int tauv_centernet_focal_loss(const float* logits, const uint8_t* valid, const int64_t* label,
                                         const float* center, int B, int n_objects, int C, int H, int W, int in_h,
                                         int in_w, int downsample_ratio, double sigma, double alpha, double beta,
                                         double* frame_sums, int64_t* frame_pos, void* workspace,
                                         size_t workspace_bytes, tauv_stream_t stream) {
  TAUV_REQUIRE(frame_sums && frame_pos, TAUV_E_NULL, "frame_sums/frame_pos must not be NULL");
  FocalArgs g;
  if (int e = focal_args(&g, logits, valid, label, center, B, n_objects, C, H, W, in_h, in_w, downsample_ratio, sigma,
                         alpha, beta))
    return e;
  const size_t part_bytes = align_up((size_t)g.n_chunks * 16, 256);
  TAUV_REQUIRE(workspace != nullptr && (uintptr_t)workspace % 256 == 0, TAUV_E_WORKSPACE, "workspace must be 256-byte aligned");
  TAUV_REQUIRE(workspace_bytes >= part_bytes + align_up((size_t)g.n_chunks * 4, 256), TAUV_E_WORKSPACE,
               "workspace %zu too small", workspace_bytes);
  double* part = reinterpret_cast<double*>(workspace);
  int* part_pos = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(workspace) + part_bytes);
  const long long grid = (g.n_chunks + kEncThreads / 32 - 1) / (kEncThreads / 32);
  if (g.a.kind == 2) focal_forward_kernel<2><<<(unsigned)grid, kEncThreads, 0, (cudaStream_t)stream>>>(g, part, part_pos);
  else if (g.a.kind == 3) focal_forward_kernel<3><<<(unsigned)grid, kEncThreads, 0, (cudaStream_t)stream>>>(g, part, part_pos);
  else focal_forward_kernel<0><<<(unsigned)grid, kEncThreads, 0, (cudaStream_t)stream>>>(g, part, part_pos);
  TAUV_LAUNCH_CHECK("focal_forward_kernel");
  focal_reduce_kernel<<<(unsigned)B, 32, 0, (cudaStream_t)stream>>>(part, part_pos, (long long)C * g.chunks_per_plane,
                                                                    frame_sums, frame_pos);
  TAUV_LAUNCH_CHECK("focal_reduce_kernel");
  return 0;
}

extern "C" int tauv_centernet_focal_loss_reduce(const double* frame_sums, const int64_t* frame_pos, int B, float* loss,
                                               int64_t* n_pos_total, tauv_stream_t stream) {
  TAUV_REQUIRE(frame_sums && frame_pos && loss && n_pos_total, TAUV_E_NULL, "pointers must not be NULL");
  TAUV_REQUIRE(B > 0, TAUV_E_SHAPE, "bad shape B=%d", B);
  focal_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(frame_sums, frame_pos, B, loss, n_pos_total);
  TAUV_LAUNCH_CHECK("focal_finish_kernel");
  return 0;
}

extern "C" int tauv_centernet_focal_loss_backward(const float* logits, const uint8_t* valid, const int64_t* label,
                                                  const float* center, int B, int n_objects, int C, int H, int W,
                                                  int in_h, int in_w, int downsample_ratio, double sigma, double alpha,
                                                  double beta, const int64_t* n_pos_total, const float* grad_out,
                                                  float* grad_logits, tauv_stream_t stream) {
  TAUV_REQUIRE(n_pos_total && grad_out && grad_logits, TAUV_E_NULL, "n_pos_total/grad_out/grad_logits must not be NULL");
  TAUV_REQUIRE((uintptr_t)grad_logits % 16 == 0, TAUV_E_ALIGN, "grad_logits must be 16-byte aligned");
  FocalArgs g;
  if (int e = focal_args(&g, logits, valid, label, center, B, n_objects, C, H, W, in_h, in_w, downsample_ratio, sigma,
                         alpha, beta))
    return e;
  const long long grid = (g.n_chunks + kEncThreads / 32 - 1) / (kEncThreads / 32);
  if (g.a.kind == 2)
    focal_backward_kernel<2><<<(unsigned)grid, kEncThreads, 0, (cudaStream_t)stream>>>(g, n_pos_total, grad_out, grad_logits);
  else if (g.a.kind == 3)
    focal_backward_kernel<3><<<(unsigned)grid, kEncThreads, 0, (cudaStream_t)stream>>>(g, n_pos_total, grad_out, grad_logits);
  else
    focal_backward_kernel<0><<<(unsigned)grid, kEncThreads, 0, (cudaStream_t)stream>>>(g, n_pos_total, grad_out, grad_logits);
  TAUV_LAUNCH_CHECK("focal_backward_kernel");
  return 0;
}

extern "C" int tauv_gaussian_encode(const uint8_t* valid, const int64_t* label, const float* center, int B,
                                    int n_objects, int C, int H, int W, int in_h, int in_w, int downsample_ratio,
                                    double sigma, float* out, tauv_stream_t stream) {
  TAUV_REQUIRE(out, TAUV_E_NULL, "out must not be NULL");
  TAUV_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && n_objects >= 0, TAUV_E_SHAPE, "bad shape");
  TAUV_REQUIRE(n_objects == 0 || (valid && label && center), TAUV_E_NULL, "valid/label/center must not be NULL");
  TAUV_REQUIRE(in_h > 0 && in_w > 0 && downsample_ratio > 0, TAUV_E_SHAPE, "bad model geometry");
  TAUV_REQUIRE(n_objects <= 4096, TAUV_E_UNSUPPORTED, "n_objects=%d exceeds the built-in limit 4096", n_objects);
  const int groups = (C + kEncGroup - 1) / kEncGroup;
  int bands, rows;
  band_plan(H, W, (long long)B * groups, &bands, &rows);
  const long long grid = (long long)B * groups * bands;
  TAUV_REQUIRE(grid < (1LL << 31), TAUV_E_UNSUPPORTED, "grid too large");
  const size_t smem = (size_t)(n_objects > 0 ? n_objects : 1) * 20;
  const bool vec = (W % 4 == 0) && ((uintptr_t)out % 16 == 0);
  const float ts = two_sigma_sq(sigma);
  if (vec && n_objects <= 32 && (uintptr_t)center % 8 == 0) {
    const long long plane_strips = (long long)H * (W / 4);
    const long long cpp = (plane_strips + kEncWarpStrips - 1) / kEncWarpStrips;
    const long long n_chunks = (long long)B * C * cpp;
    const long long wgrid = (n_chunks + kEncThreads / 32 - 1) / (kEncThreads / 32);
    TAUV_REQUIRE(wgrid < (1LL << 31) && cpp < (1LL << 31), TAUV_E_UNSUPPORTED, "grid too large");
    gaussian_encode_warp_kernel<<<(unsigned)wgrid, kEncThreads, 0, (cudaStream_t)stream>>>(
        valid, label, center, n_objects, C, H, W, (float)in_h, (float)in_w, (float)downsample_ratio, ts, (int)cpp,
        n_chunks, out);
  } else if (vec)
    gaussian_encode_kernel<true><<<(unsigned)grid, kEncThreads, smem, (cudaStream_t)stream>>>(
        valid, label, center, n_objects, C, H, W, (float)in_h, (float)in_w, (float)downsample_ratio, ts, groups, bands,
        rows, out);
  else
    gaussian_encode_kernel<false><<<(unsigned)grid, kEncThreads, smem, (cudaStream_t)stream>>>(
        valid, label, center, n_objects, C, H, W, (float)in_h, (float)in_w, (float)downsample_ratio, ts, groups, bands,
        rows, out);
  TAUV_LAUNCH_CHECK("gaussian_encode_kernel");
  return 0;
}

extern "C" int tauv_keypoint_encode(const uint8_t* kp_valid, const int64_t* kp_label, const float* kp_center,
                                    const int64_t* kp_object_index, const float* center, int B, int m, int n_objects,
                                    int Kp, int H, int W, int in_h, int in_w, int downsample_ratio,
                                    double sigma_heatmap, double sigma_affinity, float* heatmap, float* weight,
                                    float* affinity, tauv_stream_t stream) {
  TAUV_REQUIRE(heatmap && weight && affinity, TAUV_E_NULL, "outputs must not be NULL");
  TAUV_REQUIRE(B > 0 && Kp > 0 && H > 0 && W > 0 && m >= 0 && n_objects >= 0, TAUV_E_SHAPE, "bad shape");
  TAUV_REQUIRE(m == 0 || (kp_valid && kp_label && kp_center && kp_object_index && center && n_objects > 0), TAUV_E_NULL,
               "keypoint inputs must not be NULL");
  TAUV_REQUIRE(in_h > 0 && in_w > 0 && downsample_ratio > 0, TAUV_E_SHAPE, "bad model geometry");
  TAUV_REQUIRE(m <= 4096, TAUV_E_UNSUPPORTED, "n_keypoint_instances=%d exceeds the built-in limit 4096", m);
  const long long planes = (long long)B * Kp;
  int bands, rows;
  band_plan(H, W, planes, &bands, &rows);
  const long long grid = planes * bands;
  TAUV_REQUIRE(grid < (1LL << 31), TAUV_E_UNSUPPORTED, "grid too large");
  const size_t smem = (size_t)(m > 0 ? m : 1) * 16;
  const bool vec = (W % 4 == 0) && ((uintptr_t)heatmap % 16 == 0) && ((uintptr_t)weight % 16 == 0) &&
                   ((uintptr_t)affinity % 16 == 0) && (((size_t)H * W) % 4 == 0);
  // generate_keypoint_heatmap does not floor sigma (loss.py:101,109): plain 2*sigma**2
  const float tsh = (float)(2.0 * (sigma_heatmap * sigma_heatmap));
  const float tsa = (float)(2.0 * (sigma_affinity * sigma_affinity));
  if (vec && m <= 128 && m > 0) {
    const long long plane_strips = (long long)H * (W / 4);
    const long long cpp = (plane_strips + kKpWarpStrips - 1) / kKpWarpStrips;
    const long long n_chunks = planes * cpp;
    const long long wgrid = (n_chunks + kEncThreads / 32 - 1) / (kEncThreads / 32);
    TAUV_REQUIRE(wgrid < (1LL << 31) && cpp < (1LL << 31), TAUV_E_UNSUPPORTED, "grid too large");
#define TAUV_KP_WARP(NCH)                                                                                             \
    keypoint_encode_warp_kernel<NCH, 0><<<(unsigned)wgrid, kEncThreads, 0, (cudaStream_t)stream>>>(                    \
        kp_valid, kp_label, kp_center, kp_object_index, center, m, n_objects, Kp, H, W, (float)in_h, (float)in_w,      \
        (float)downsample_ratio, tsh, tsa, (int)cpp, n_chunks, heatmap, weight, affinity, nullptr, nullptr, nullptr)
    if (m <= 32) TAUV_KP_WARP(1);
    else if (m <= 64) TAUV_KP_WARP(2);
    else TAUV_KP_WARP(4);
#undef TAUV_KP_WARP
  } else if (vec)
    keypoint_encode_kernel<true><<<(unsigned)grid, kEncThreads, smem, (cudaStream_t)stream>>>(
        kp_valid, kp_label, kp_center, kp_object_index, center, m, n_objects, Kp, H, W, (float)in_h, (float)in_w,
        (float)downsample_ratio, tsh, tsa, bands, rows, heatmap, weight, affinity);
  else
    keypoint_encode_kernel<false><<<(unsigned)grid, kEncThreads, smem, (cudaStream_t)stream>>>(
        kp_valid, kp_label, kp_center, kp_object_index, center, m, n_objects, Kp, H, W, (float)in_h, (float)in_w,
        (float)downsample_ratio, tsh, tsa, bands, rows, heatmap, weight, affinity);
  TAUV_LAUNCH_CHECK("keypoint_encode_kernel");
  return 0;
}

// Keypoint-affinity term fused with its target render (loss.py:244-246, before the lambda).
static int kp_affinity_plan(int B, int m, int n_objects, int Kp, int H, int W, int in_h, int in_w, int ratio,
                            const void* a, const void* b, long long* cpp, long long* n_chunks, long long* wgrid) {
  TAUV_REQUIRE(B > 0 && Kp > 0 && H > 0 && W > 0 && m > 0 && n_objects > 0, TAUV_E_SHAPE, "bad shape");
  TAUV_REQUIRE(in_h > 0 && in_w > 0 && ratio > 0, TAUV_E_SHAPE, "bad model geometry");
  TAUV_REQUIRE(W % 4 == 0 && m <= 128 && (uintptr_t)a % 16 == 0 && (uintptr_t)b % 16 == 0, TAUV_E_UNSUPPORTED,
               "the fused affinity loss needs W %% 4 == 0, 16-byte aligned tensors and <= 128 keypoint instances per frame");
  const long long planes = (long long)B * Kp, plane_strips = (long long)H * (W / 4);
  *cpp = (plane_strips + kKpWarpStrips - 1) / kKpWarpStrips;
  *n_chunks = planes * *cpp;
  *wgrid = (*n_chunks + kEncThreads / 32 - 1) / (kEncThreads / 32);
  TAUV_REQUIRE(*wgrid < (1LL << 31) && *cpp < (1LL << 31), TAUV_E_UNSUPPORTED, "grid too large");
  return 0;
}

extern "C" size_t tauv_keypoint_affinity_loss_partials(int B, int Kp, int H, int W) {
  if (B <= 0 || Kp <= 0 || H <= 0 || W <= 0 || W % 4) return 0;
  const long long plane_strips = (long long)H * (W / 4);
  return (size_t)((long long)B * Kp * ((plane_strips + kKpWarpStrips - 1) / kKpWarpStrips));
}

#define TAUV_KP_LOSS(NCH, MODE, OUT, PRED, PART, GRAD)                                                                \
  keypoint_encode_warp_kernel<NCH, MODE><<<(unsigned)wgrid, kEncThreads, 0, (cudaStream_t)stream>>>(                   \
      kp_valid, kp_label, kp_center, kp_object_index, center, m, n_objects, Kp, H, W, (float)in_h, (float)in_w,        \
      (float)downsample_ratio, 1.0f, tsa, (int)cpp, n_chunks, nullptr, nullptr, OUT, PRED, PART, GRAD)

extern "C" int tauv_keypoint_affinity_loss(const float* pred_affinity, const uint8_t* kp_valid, const int64_t* kp_label,
                                           const float* kp_center, const int64_t* kp_object_index, const float* center,
                                           int B, int m, int n_objects, int Kp, int H, int W, int in_h, int in_w,
                                           int downsample_ratio, double sigma_affinity, double* partial,
                                           tauv_stream_t stream) {
  TAUV_REQUIRE(pred_affinity && kp_valid && kp_label && kp_center && kp_object_index && center && partial, TAUV_E_NULL,
               "pointers must not be NULL");
  long long cpp, n_chunks, wgrid;
  if (int rc = kp_affinity_plan(B, m, n_objects, Kp, H, W, in_h, in_w, downsample_ratio, pred_affinity, pred_affinity, &cpp,
                                &n_chunks, &wgrid))
    return rc;
  const float tsa = (float)(2.0 * (sigma_affinity * sigma_affinity));
  if (m <= 32) TAUV_KP_LOSS(1, 1, nullptr, pred_affinity, partial, nullptr);
  else if (m <= 64) TAUV_KP_LOSS(2, 1, nullptr, pred_affinity, partial, nullptr);
  else TAUV_KP_LOSS(4, 1, nullptr, pred_affinity, partial, nullptr);
  TAUV_LAUNCH_CHECK("keypoint_encode_warp_kernel<loss>");
  return 0;
}

extern "C" int tauv_keypoint_affinity_loss_backward(const float* pred_affinity, const uint8_t* kp_valid,
                                                    const int64_t* kp_label, const float* kp_center,
                                                    const int64_t* kp_object_index, const float* center, int B, int m,
                                                    int n_objects, int Kp, int H, int W, int in_h, int in_w,
                                                    int downsample_ratio, double sigma_affinity, const float* grad_out,
                                                    float* grad_affinity, tauv_stream_t stream) {
  TAUV_REQUIRE(pred_affinity && kp_valid && kp_label && kp_center && kp_object_index && center && grad_out && grad_affinity,
               TAUV_E_NULL, "pointers must not be NULL");
  long long cpp, n_chunks, wgrid;
  if (int rc = kp_affinity_plan(B, m, n_objects, Kp, H, W, in_h, in_w, downsample_ratio, pred_affinity, grad_affinity, &cpp,
                                &n_chunks, &wgrid))
    return rc;
  const float tsa = (float)(2.0 * (sigma_affinity * sigma_affinity));
  if (m <= 32) TAUV_KP_LOSS(1, 2, grad_affinity, pred_affinity, nullptr, grad_out);
  else if (m <= 64) TAUV_KP_LOSS(2, 2, grad_affinity, pred_affinity, nullptr, grad_out);
  else TAUV_KP_LOSS(4, 2, grad_affinity, pred_affinity, nullptr, grad_out);
  TAUV_LAUNCH_CHECK("keypoint_encode_warp_kernel<loss backward>");
  return 0;
}
#undef TAUV_KP_LOSS

extern "C" int tauv_out_index_offset(const float* position, int64_t n, int in_h, int in_w, int downsample_ratio,
                                     int out_h, int out_w, int64_t* index, float* offset, tauv_stream_t stream) {
  TAUV_REQUIRE(n >= 0, TAUV_E_SHAPE, "n must be >= 0");
  if (n == 0) return 0;
  TAUV_REQUIRE(position && index, TAUV_E_NULL, "position/index must not be NULL");
  TAUV_REQUIRE(in_h > 0 && in_w > 0 && downsample_ratio > 0 && out_h > 0 && out_w > 0, TAUV_E_SHAPE, "bad model geometry");
  const long long tot = 2 * (long long)n;
  out_index_offset_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      position, n, (float)in_h, (float)in_w, (float)downsample_ratio, downsample_ratio, out_h, out_w, index, offset);
  TAUV_LAUNCH_CHECK("out_index_offset_kernel");
  return 0;
}

extern "C" int tauv_gaussian_splat(int h, int w, int cy, int cx, double sigma, float* out, tauv_stream_t stream) {
  TAUV_REQUIRE(out, TAUV_E_NULL, "out must not be NULL");
  TAUV_REQUIRE(h > 0 && w > 0, TAUV_E_SHAPE, "bad shape");
  const long long n = (long long)h * w;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gaussian_splat_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(h, w, cy, cx,
                                                                            (float)(2.0 * (sigma * sigma)), out);
  TAUV_LAUNCH_CHECK("gaussian_splat_kernel");
  return 0;
}
