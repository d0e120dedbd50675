// centernet_select.cuh — CenterNet decode as "block maxima, then select" (round 2; included by centernet_decode.cu).
//
// Replaces the same reference lines as centernet_decode.cu (centernet/model/decode.py:182 sigmoid, :239-252 heatmap_nms,
// :255-279 heatmap_detect, :204-234 per-detection gather + box arithmetic), for the common case: 128-bit aligned maps
// with W % 4 == 0, k <= 256, sigmoid + 3x3 peak mode.  Two launches, the second one programmatically dependent:
//
//   block_max_kernel : the ONLY pass over the logits.  No data-dependent work at all — a thread reads a block of
//       4 columns x 8 rows of one plane (eight independent 128-bit loads), keeps the maximum, writes it (one float per
//       32 cells = 3 % of the bytes read), and a warp reduction adds the maximum of each 32 consecutive blocks ("group",
//       1024 cells).  Nothing depends on a threshold, so there is no bootstrap, no cluster, no queue: the kernel runs at
//       the rate of a plain 128-bit read stream.
//   select_kernel    : one CTA (1024 threads) per frame works on the summaries, which are still in L2 (164 KB + 5 KB per
//       frame at 80 x 128 x 128).  A peak of value v lies in a block whose maximum is >= v, so for ANY threshold T the
//       peaks >= T are found by examining the blocks with maximum >= T only:
//         1. T := the K1-th largest of <= 1024 strided maxima of the group (or block) maxima, K1 = k + k/8 + 8 (radix
//            select) — on noise about 1.1 K1 blocks reach it, and 94 % of the block maxima are peaks;
//         2. hot groups -> hot blocks -> their cells >= T ("hot cells") -> 3x3 test with eight lanes per hot cell (the
//            neighbours come from HBM: two dependent round trips in all, every load of a phase in flight at once);
//         3. the peaks >= T get their sigmoid and final sort key (score desc, flat index asc); the rank of a key among
//            the (distinct) keys is its output slot — no sort;
//         4. the result is complete iff there are >= k of them AND every logit below T has a score strictly below the
//            k-th best score (reject_key_for_score, the same guard band as the round-1 kernel).  Otherwise K1 *= 4 and
//            again; when that runs out (plateaus, saturated scores, fewer than k peaks) the frame is done exhaustively
//            in segments with exact pruning (sel_slow) — slow, exact, and only for degenerate maps.
//       The same threads then gather size / offset / depth through the strided views and do the box arithmetic.
#pragma once

namespace tauv {

constexpr int kBmRows = 8;            // rows of a block (a block = 4 columns x kBmRows rows of one plane)
constexpr int kBmThreads = 256;
constexpr int kSelThreads = 1024;
constexpr int kSelHotCap = 1024;      // hot groups / hot blocks per attempt
constexpr int kSelCellCap = 2048;     // hot cells per attempt
constexpr int kSelCandCap = 2048;     // candidate list (final composites)
constexpr int kSelMaxK = 256;
constexpr int kSelSmallSeg = (kSelCandCap - 2 * kSelMaxK) / 32;  // blocks per exhaustive sub-step that cannot overflow

struct BmArgs {
  const float* hm;
  float* bm;    // [B][n_blk]  block maxima
  float* bm2;   // [B][G]      maxima of 32 consecutive blocks
  int C, H, W, W4, n_rg, n_blk, G;
};

struct SelArgs {
  const float* hm;
  const float* bm;
  const float* bm2;
  int C, H, W, k, W4, n_rg, n_blk, G;
  int64_t* out_index;
  int64_t* out_label;
  float* out_score;
  BoxArgs box;
};

__global__ void __launch_bounds__(kBmThreads) block_max_kernel(const __grid_constant__ BmArgs a) {
  // the dependent launch (select_kernel) may be set up as soon as every CTA of this grid has started; it waits for
  // this grid's completion and memory flush before it reads anything
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int b = blockIdx.y;
  const uint32_t i = blockIdx.x * (uint32_t)kBmThreads + threadIdx.x;  // block index inside the frame
  float m = TAUV_NEG_INF;
  const bool valid = i < (uint32_t)a.n_blk;
  if (valid) {
    const uint32_t q = i / (uint32_t)a.W4, c4 = i - q * (uint32_t)a.W4;  // q = c * n_rg + rg
    const uint32_t c = q / (uint32_t)a.n_rg, rg = q - c * (uint32_t)a.n_rg;
    const int r0 = (int)rg * kBmRows;
    const int nr = min(kBmRows, a.H - r0);
    const float* p = a.hm + (((size_t)b * a.C + c) * a.H + r0) * (size_t)a.W + 4 * c4;
    float4 x[kBmRows];
    if (nr == kBmRows) {  // all eight loads in flight before the first use
#pragma unroll
      for (int r = 0; r < kBmRows; ++r) x[r] = ldg_stream4(p + (size_t)r * a.W);
    } else {
#pragma unroll
      for (int r = 0; r < kBmRows; ++r)
        x[r] = r < nr ? ldg_stream4(p + (size_t)r * a.W) : make_float4(TAUV_NEG_INF, TAUV_NEG_INF, TAUV_NEG_INF, TAUV_NEG_INF);
    }
#pragma unroll
    for (int r = 0; r < kBmRows; ++r) m = fmaxf(m, fmaxf(fmaxf(x[r].x, x[r].y), fmaxf(x[r].z, x[r].w)));
    a.bm[(size_t)b * a.n_blk + i] = m;
  }
  float g = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, o));
  if ((threadIdx.x & 31) == 0 && valid) a.bm2[(size_t)b * a.G + (i >> 5)] = g;
}

struct __align__(16) SelShared {
  union {
    struct {
      unsigned long long keys[kSelThreads];  // T selection: composite keys of the strided maxima; later: the ranked keys
      uint32_t hist[kRadixBins];             // radix histogram; later: filler flags
    } s;
    unsigned long long cell[kSelCellCap];    // hot cells: logit bits << 32 | flat index (between collect and phase B)
  } u;
  unsigned long long cand[kSelCandCap];      // peaks: final composites (score key << 32 | ~flat index)
  uint32_t hot[kSelHotCap];
  uint32_t hgrp[kSelHotCap];
  uint32_t ctl[8];
  int n_hgrp, n_hot, n_cell, n_cand, n_out, first_below, flag;
  float s_k, xc_f;
  unsigned long long thr_c;
};

__device__ __forceinline__ bool sel_is_peak(float x, float m) {
  bool peak = x >= m;
  if (!peak && (x > 4.0f || m < -80.0f || (m - x) < 1e-3f)) peak = sigmoid_tie(x, m);
  return peak;
}

__device__ __forceinline__ unsigned long long sel_final_key(float x, uint32_t flat) {
  const float s = sigmoid_ref(x);
  return s > 0.0f ? (((unsigned long long)float_to_key(s) << 32) | (unsigned long long)(~flat)) : 0ull;
}

// warp-aggregated append: returns the slot of this lane's entry (or -1 when `want` is false); *counter only grows
__device__ __forceinline__ int sel_append(int* counter, bool want) {
  const unsigned bal = __ballot_sync(0xffffffffu, want);
  if (bal == 0u) return -1;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == __ffs(bal) - 1) base = atomicAdd(counter, __popc(bal));
  base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
  return want ? base + __popc(bal & ((1u << lane) - 1u)) : -1;
}

// Rank the n candidates (distinct non-zero composites) by counting, write the min(n, k) best in order, and leave the
// k-th best score in sh->s_k (only meaningful when n >= k).  Returns npos = min(n, k).  All threads call this.
__device__ __noinline__ int sel_rank_emit(const SelArgs& a, SelShared* sh, int b, int n) {
  const int tid = threadIdx.x, k = a.k;
  unsigned long long* ranked = sh->u.s.keys;
  const int npos = n < k ? n : k;
  for (int i0 = 0; i0 < n; i0 += kSelThreads / 8) {
    const int i = i0 + (tid >> 3), part = tid & 7;
    const unsigned long long my = i < n ? sh->cand[i] : ~0ull;
    int cnt = 0;
    for (int j = part; j < n; j += 8) cnt += sh->cand[j] > my ? 1 : 0;
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
    if (i < n && part == 0 && cnt < k) ranked[cnt] = my;
  }
  __syncthreads();
  const long long hw_elems = (long long)a.H * a.W;
  if (tid < npos) {
    const unsigned long long c = ranked[tid];
    const uint32_t flat = composite_idx(c);
    const float s = key_to_float(composite_key(c));
    const long long lab = flat / hw_elems;
    const long long rem = flat - lab * hw_elems;
    const int iy = (int)(rem / a.W), ix = (int)(rem - (long long)iy * a.W);
    const long long slot = (long long)b * k + tid;
    a.out_index[slot * 2 + 0] = iy;
    a.out_index[slot * 2 + 1] = ix;
    a.out_label[slot] = lab;
    a.out_score[slot] = s;
    if (tid == k - 1) sh->s_k = s;
    if (a.box.enabled) {
      box_one(a.box, b, slot, iy, ix);
      if (s < a.box.thr) atomicMin(&sh->first_below, tid);
    }
  }
  __syncthreads();
  return npos;
}

// The zero-score tail (fewer than k positive peaks in the whole frame) and the count of leading detections.
__device__ __noinline__ void sel_finish(const SelArgs& a, SelShared* sh, int b, int npos) {
  const int tid = threadIdx.x, k = a.k;
  if (npos < k) {
    uint32_t* flags = sh->u.s.hist;
    const unsigned long long* ranked = sh->u.s.keys;
    for (int i = tid; i < k; i += kSelThreads) flags[i] = 0u;
    __syncthreads();
    if (tid < npos) {
      const uint32_t flat = composite_idx(ranked[tid]);
      if (flat < (uint32_t)k) flags[flat] = 1u;
    }
    __syncthreads();
    topk_emit_fillers<kSelThreads>(flags, npos, b, k, a.H, a.W, a.out_index, a.out_label, a.out_score, a.box);
    if (a.box.enabled && 0.0f < a.box.thr && tid == 0) atomicMin(&sh->first_below, npos);
  }
  if (a.box.enabled) {
    __syncthreads();
    if (tid == 0) a.box.count[b] = sh->first_below;
  }
}

// One exhaustive step over the blocks [seg0, seg0 + len), len <= kSelThreads: blocks whose maximum reaches the logit
// filter are examined cell by cell (one warp per block, one lane per cell), peaks at or above the current k-th best
// key are appended.  Sets sh->flag when the list overflowed (the caller redoes the range in smaller steps).
__device__ __noinline__ void sel_slow_step(const SelArgs& a, SelShared* sh, const float* __restrict__ fhm,
                                           const float* __restrict__ bm, int seg0, int len) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) sh->n_hot = 0;
  __syncthreads();
  const float xc = sh->xc_f;
  const unsigned long long thr_c = sh->thr_c;
  {
    const bool hot = tid < len && !(bm[seg0 + tid] < xc);
    const int slot = sel_append(&sh->n_hot, hot);
    if (hot) sh->hot[slot] = (uint32_t)(seg0 + tid);
  }
  __syncthreads();
  const int nhot = sh->n_hot;
  const int HW = a.H * a.W;
  for (int j = warp; j < nhot; j += kSelThreads / 32) {
    const uint32_t blk = sh->hot[j];
    const uint32_t q = blk / (uint32_t)a.W4, c4 = blk - q * (uint32_t)a.W4;
    const uint32_t c = q / (uint32_t)a.n_rg, rg = q - c * (uint32_t)a.n_rg;
    const int r = (int)rg * kBmRows + (lane >> 2), col = (int)c4 * 4 + (lane & 3);
    unsigned long long fin = 0ull;
    if (r < a.H) {
      const float* pl = fhm + (size_t)c * HW;
      const float x = pl[r * a.W + col];
      if (!(x < xc)) {
        float m = TAUV_NEG_INF;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            if (dy == 0 && dx == 0) continue;
            const int rr = r + dy, cc = col + dx;
            if (rr >= 0 && rr < a.H && cc >= 0 && cc < a.W) m = fmaxf(m, pl[rr * a.W + cc]);
          }
        if (sel_is_peak(x, m)) fin = sel_final_key(x, c * (uint32_t)HW + (uint32_t)(r * a.W + col));
      }
    }
    const bool want = fin != 0ull && fin >= thr_c;
    const int slot = sel_append(&sh->n_cand, want);
    if (want) {
      if (slot < kSelCandCap) sh->cand[slot] = fin;
      else sh->flag = 1;
    }
  }
  __syncthreads();
}

// exact prune of the candidate list to its k best; raises the key threshold and the logit filter
__device__ __noinline__ void sel_slow_prune(const SelArgs& a, SelShared* sh) {
  const int tid = threadIdx.x, k = a.k;
  const int n = sh->n_cand;  // > k
  const unsigned long long T =
      block_kth_largest<kSelThreads>([&](int i) { return sh->cand[i]; }, n, k, sh->u.s.hist, sh->ctl);
  unsigned long long mine[kSelCandCap / kSelThreads];
#pragma unroll
  for (int u = 0; u < kSelCandCap / kSelThreads; ++u) {
    const int i = tid + u * kSelThreads;
    mine[u] = i < n ? sh->cand[i] : 0ull;
  }
  if (tid == 0) sh->n_cand = 0;
  __syncthreads();
#pragma unroll
  for (int u = 0; u < kSelCandCap / kSelThreads; ++u)
    if (mine[u] >= T && mine[u] != 0ull) sh->cand[atomicAdd(&sh->n_cand, 1)] = mine[u];
  if (tid == 0) {
    sh->thr_c = T;
    const uint32_t rk = reject_key_for_score(key_to_float(composite_key(T)));
    if (rk) {
      const float xc = key_to_float(rk);
      if (xc > sh->xc_f) sh->xc_f = xc;
    }
  }
  __syncthreads();
}

// The whole frame, exhaustively, in segments with exact pruning.  Returns the number of candidates left in sh->cand.
__device__ __noinline__ int sel_slow(const SelArgs& a, SelShared* sh, const float* __restrict__ fhm,
                                     const float* __restrict__ bm) {
  const int tid = threadIdx.x;
  const int soft = 2 * a.k;
  __syncthreads();
  if (tid == 0) {
    sh->n_cand = 0;
    sh->flag = 0;
    sh->thr_c = 0ull;
    sh->xc_f = TAUV_NEG_INF;
  }
  __syncthreads();
  int seg = 0;
  while (seg < a.n_blk) {
    int len = kSelThreads;  // optimistic: a segment whose peaks do not fit is redone in sub-steps that cannot overflow
    if (len > a.n_blk - seg) len = a.n_blk - seg;
    const int n0 = sh->n_cand;
    sel_slow_step(a, sh, fhm, bm, seg, len);
    if (sh->flag) {  // (uniform: read after the step's closing barrier)
      __syncthreads();
      if (tid == 0) {
        sh->n_cand = n0;
        sh->flag = 0;
      }
      __syncthreads();
      for (int s = seg; s < seg + len; s += kSelSmallSeg) {
        int l2 = kSelSmallSeg;
        if (l2 > seg + len - s) l2 = seg + len - s;
        sel_slow_step(a, sh, fhm, bm, s, l2);
        if (sh->n_cand > soft) sel_slow_prune(a, sh);
      }
    } else if (sh->n_cand > soft) {
      sel_slow_prune(a, sh);
    }
    seg += len;
  }
  __syncthreads();
  return sh->n_cand;
}

__global__ void __launch_bounds__(kSelThreads, 1) select_kernel(const __grid_constant__ SelArgs a) {
  __shared__ SelShared sh_;
  SelShared* sh = &sh_;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const int k = a.k;
  const int HW = a.H * a.W;
  const float* __restrict__ fhm = a.hm + (size_t)b * a.C * HW;
  const float* __restrict__ bm = a.bm + (size_t)b * a.n_blk;
  const float* __restrict__ bm2 = a.bm2 + (size_t)b * a.G;
  if (tid == 0) sh->first_below = k;
  asm volatile("griddepcontrol.wait;" ::: "memory");  // block_max_kernel has completed and flushed (no-op without the attribute)

  int K1 = k + (k >> 3) + 8;
  int n = 0;
  bool slow = false;
  for (;;) {
    // ---- 1. threshold: the K1-th largest of <= 1024 strided maxima of the group (or block) maxima ----
    const bool use_grp = a.G >= 4 * K1;
    const float* __restrict__ lv = use_grp ? bm2 : bm;
    const int nlv = use_grp ? a.G : a.n_blk;
    float T_f = TAUV_NEG_INF;
    __syncthreads();
    if (K1 < (nlv < kSelThreads ? nlv : kSelThreads)) {
      float tm = TAUV_NEG_INF;
      for (int i = tid; i < nlv; i += kSelThreads) tm = fmaxf(tm, lv[i]);
      sh->u.s.keys[tid] = make_composite(float_to_key(tm), (uint32_t)tid);
      __syncthreads();
      const unsigned long long Tc = block_kth_largest<kSelThreads>([&](int i) { return sh->u.s.keys[i]; }, kSelThreads,
                                                                   K1, sh->u.s.hist, sh->ctl);
      T_f = key_to_float(composite_key(Tc));
    }
    if (tid == 0) {
      sh->n_hgrp = 0;
      sh->n_hot = 0;
      sh->n_cell = 0;
      sh->n_cand = 0;
    }
    __syncthreads();
    // ---- 2. hot groups -> hot blocks ----
    if (use_grp) {
      for (int i0 = 0; i0 < a.G; i0 += kSelThreads) {
        const int i = i0 + tid;
        const bool hot = i < a.G && !(bm2[i] < T_f);
        const int slot = sel_append(&sh->n_hgrp, hot);
        if (hot && slot < kSelHotCap) sh->hgrp[slot] = (uint32_t)i;
      }
      __syncthreads();
      const int nh = sh->n_hgrp;
      if (nh > kSelHotCap) { slow = true; break; }
      for (int j = warp; j < nh; j += kSelThreads / 32) {
        const int idx = (int)sh->hgrp[j] * 32 + lane;
        const bool hot = idx < a.n_blk && !(bm[idx < a.n_blk ? idx : 0] < T_f);
        const int slot = sel_append(&sh->n_hot, hot);
        if (hot && slot < kSelHotCap) sh->hot[slot] = (uint32_t)idx;
      }
    } else {
      for (int i0 = 0; i0 < a.n_blk; i0 += kSelThreads) {
        const int i = i0 + tid;
        const bool hot = i < a.n_blk && !(bm[i < a.n_blk ? i : 0] < T_f);
        const int slot = sel_append(&sh->n_hot, hot);
        if (hot && slot < kSelHotCap) sh->hot[slot] = (uint32_t)i;
      }
    }
    __syncthreads();
    const int nhot = sh->n_hot;
    if (nhot > kSelHotCap) { slow = true; break; }
    // ---- 3a. the cells of the hot blocks (one warp per block, one lane per cell): which reach T? ----
    constexpr int UA = 8;
    for (int j0 = 0; j0 < nhot; j0 += UA * (kSelThreads / 32)) {
      float x[UA];
      uint32_t fl[UA];
#pragma unroll
      for (int u = 0; u < UA; ++u) {
        const int j = j0 + u * (kSelThreads / 32) + warp;
        x[u] = TAUV_NEG_INF;
        fl[u] = 0xffffffffu;
        if (j < nhot) {
          const uint32_t blk = sh->hot[j];
          const uint32_t q = blk / (uint32_t)a.W4, c4 = blk - q * (uint32_t)a.W4;
          const uint32_t c = q / (uint32_t)a.n_rg, rg = q - c * (uint32_t)a.n_rg;
          const int r = (int)rg * kBmRows + (lane >> 2), col = (int)c4 * 4 + (lane & 3);
          if (r < a.H) {
            fl[u] = c * (uint32_t)HW + (uint32_t)(r * a.W + col);
            x[u] = fhm[fl[u]];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UA; ++u) {
        const bool want = fl[u] != 0xffffffffu && !(x[u] < T_f);
        const int slot = sel_append(&sh->n_cell, want);
        if (want && slot < kSelCellCap) sh->u.cell[slot] = ((unsigned long long)__float_as_uint(x[u]) << 32) | fl[u];
      }
    }
    __syncthreads();
    const int ncell = sh->n_cell;
    if (ncell > kSelCellCap) { slow = true; break; }
    // ---- 3b. 3x3 test, eight lanes per hot cell (one neighbour each) ----
    for (int i0 = 0; i0 < ncell; i0 += kSelThreads / 8) {
      const int i = i0 + (tid >> 3), part = tid & 7;
      float x = 0.0f, v = TAUV_NEG_INF;
      uint32_t flat = 0;
      if (i < ncell) {
        const unsigned long long e = sh->u.cell[i];
        x = __uint_as_float((uint32_t)(e >> 32));
        flat = (uint32_t)e;
        const uint32_t c = flat / (uint32_t)HW, rem = flat - c * (uint32_t)HW;
        const int r = (int)(rem / (uint32_t)a.W), col = (int)(rem - (uint32_t)r * (uint32_t)a.W);
        const int p = part + (part >= 4 ? 1 : 0);  // 0..8 without the centre
        const int dy = p / 3 - 1, dx = p - (p / 3) * 3 - 1;
        const int rr = r + dy, cc = col + dx;
        if (rr >= 0 && rr < a.H && cc >= 0 && cc < a.W) v = fhm[(long long)flat + dy * a.W + dx];
      }
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
      unsigned long long fin = 0ull;
      if (i < ncell && part == 0 && sel_is_peak(x, v)) fin = sel_final_key(x, flat);
      const int slot = sel_append(&sh->n_cand, fin != 0ull);
      if (fin != 0ull) sh->cand[slot] = fin;  // (n_cand <= n_cell <= kSelCandCap)
    }
    __syncthreads();
    n = sh->n_cand;
    const bool have_T = T_f > TAUV_NEG_INF;
    if (n < k && have_T) {  // not enough peaks at or above T: lower it
      K1 *= 4;
      continue;
    }
    // ---- 4. rank, emit, and check that nothing below T could have made it ----
    const int npos = sel_rank_emit(a, sh, b, n);
    if (have_T) {
      if (tid == 0) {
        const uint32_t rk = reject_key_for_score(sh->s_k);
        sh->flag = (rk != 0u && rk >= float_to_key(T_f)) ? 0 : 1;
      }
      __syncthreads();
      if (sh->flag) {
        __syncthreads();
        if (tid == 0) {
          sh->flag = 0;
          sh->first_below = k;
        }
        K1 *= 4;
        continue;
      }
    }
    sel_finish(a, sh, b, npos);
    return;
  }
  if (slow) {
    if (tid == 0) sh->first_below = k;
    n = sel_slow(a, sh, fhm, bm);
    const int npos = sel_rank_emit(a, sh, b, n);
    sel_finish(a, sh, b, npos);
  }
}

// ---- host side ----
struct SelPlan {
  int W4, n_rg, n_blk, G;
  size_t bm_bytes, bm2_bytes;
};

static bool select_plan(int B, int C, int H, int W, int k, SelPlan* p) {
  if (W % 4 != 0 || k > kSelMaxK || B > 65535) return false;
  const long long n_rg = (H + kBmRows - 1) / kBmRows;
  const long long n_blk = (long long)C * n_rg * (W / 4);
  if (n_blk >= (1LL << 26)) return false;
  p->W4 = W / 4;
  p->n_rg = (int)n_rg;
  p->n_blk = (int)n_blk;
  p->G = (int)((n_blk + 31) / 32);
  p->bm_bytes = align_up((size_t)B * p->n_blk * 4, 256);
  p->bm2_bytes = align_up((size_t)B * p->G * 4, 256);
  return true;
}

static int run_select_decode(const float* hm, int B, int C, int H, int W, int k, int64_t* index, int64_t* label,
                             float* score, const BoxArgs& box, void* ws, size_t ws_bytes, const SelPlan& p,
                             cudaStream_t st) {
  TAUV_REQUIRE(ws != nullptr && (uintptr_t)ws % 256 == 0, TAUV_E_WORKSPACE, "workspace must be 256-byte aligned");
  TAUV_REQUIRE(ws_bytes >= p.bm_bytes + p.bm2_bytes, TAUV_E_WORKSPACE, "workspace %zu < required %zu", ws_bytes,
               p.bm_bytes + p.bm2_bytes);
  BmArgs ba;
  ba.hm = hm;
  ba.bm = reinterpret_cast<float*>(ws);
  ba.bm2 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ws) + p.bm_bytes);
  ba.C = C; ba.H = H; ba.W = W; ba.W4 = p.W4; ba.n_rg = p.n_rg; ba.n_blk = p.n_blk; ba.G = p.G;
  block_max_kernel<<<dim3((unsigned)((p.n_blk + kBmThreads - 1) / kBmThreads), (unsigned)B), kBmThreads, 0, st>>>(ba);
  TAUV_LAUNCH_CHECK("block_max_kernel");
  SelArgs sa;
  sa.hm = hm; sa.bm = ba.bm; sa.bm2 = ba.bm2;
  sa.C = C; sa.H = H; sa.W = W; sa.k = k; sa.W4 = p.W4; sa.n_rg = p.n_rg; sa.n_blk = p.n_blk; sa.G = p.G;
  sa.out_index = index; sa.out_label = label; sa.out_score = score;
  sa.box = box;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)B);
  cfg.blockDim = dim3(kSelThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TAUV_CUDA(cudaLaunchKernelEx(&cfg, select_kernel, sa));
  TAUV_LAUNCH_CHECK("select_kernel");
  return 0;
}

}  // namespace tauv
