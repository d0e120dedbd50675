// centernet_select.cuh — CenterNet decode as "block maxima, then select" (round 2; included by centernet_decode.cu).
//
// Replaces the same reference lines as centernet_decode.cu (centernet/model/decode.py:182 sigmoid, :239-252 heatmap_nms,
// :255-279 heatmap_detect, :204-234 per-detection gather + box arithmetic), for the common case: 128-bit aligned maps
// with W % 4 == 0, k <= 1024, sigmoid + 3x3 peak mode.  Two launches, the second one programmatically dependent:
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
//            k-th best score (reject_key_for_score, the same guard band as the round-1 kernel).  Otherwise K1 grows and
//            again; when that runs out (plateaus, saturated scores, fewer than k peaks) the frame is done exhaustively
//            in segments with exact pruning (sel_slow) — slow, exact, and only for degenerate maps.
//       The same threads then gather size / offset / depth through the strided views and do the box arithmetic.
#pragma once

namespace tauv {

constexpr int kBmRows = 8;            // rows of a block (a block = 4 columns x kBmRows rows of one plane)
#ifndef TAUV_BM_THREADS
#define TAUV_BM_THREADS 512
#endif
constexpr int kBmThreads = TAUV_BM_THREADS;
constexpr int kSelThreads = 1024;
// (capacities: SelTier below — the shared-memory footprint decides whether the CTAs can be set up next to the draining
// block_max_kernel CTAs: 144 KB can, 204 KB cannot and costs 4 us, so only k > 256 takes the large tier)
constexpr int kSelMaxK = 1024;

// Division by a launch constant as multiply-high + shift (x < 2^31; the scheme of CUTLASS' FastDivmod): the block and
// cell index arithmetic of the select pass sits on single-warp dependent chains, where a 32-bit division costs ~100 cycles.
struct FastDiv {
  uint32_t d, mul, shr;
};
static FastDiv make_fastdiv(uint32_t d) {
  FastDiv f{d, 0u, 0u};
  if (d > 1) {
    int lg = 0;
    while ((1ull << lg) < d) ++lg;
    const int p = 31 + lg;
    f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
    f.shr = (uint32_t)(p - 32);
  }
  return f;
}
__device__ __forceinline__ uint32_t fdiv_u32(uint32_t x, const FastDiv& f) {
  return f.d == 1u ? x : (__umulhi(x, f.mul) >> f.shr);
}

struct BmArgs {
  const float* hm;
  float* bm;    // [B][32 G]   block maxima (n_blk per frame, rows padded to whole groups)
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
  FastDiv dW4, dNrg, dHW, dW;  // divisions by W/4, rows groups per plane, H*W, W
  long long* trace;  // -DTAUV_DEBUG builds only (tools/select_trace.py): 32 int64 per frame, else NULL
};

#ifdef TAUV_DEBUG
static long long* g_sel_trace = nullptr;
#define SEL_STAMP(i)                                                  \
  do {                                                                \
    if (a.trace && threadIdx.x == 0) {                                \
      long long t_;                                                   \
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_));           \
      a.trace[(size_t)blockIdx.x * 32 + sel_rep * 16 + (i)] = t_;                   \
    }                                                                 \
  } while (0)
#define SEL_NOTE(i, v)                                                \
  do {                                                                \
    if (a.trace && threadIdx.x == 0) a.trace[(size_t)blockIdx.x * 32 + sel_rep * 16 + (i)] = (long long)(v); \
  } while (0)
#else
#define SEL_STAMP(i) do {} while (0)
#define SEL_NOTE(i, v) do {} while (0)
#endif

__device__ __forceinline__ void st_keep(float* p, float v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(policy) : "memory");
}

__global__ void __launch_bounds__(kBmThreads) block_max_kernel(const __grid_constant__ BmArgs a) {
  // the dependent launch (select_kernel) may be set up as soon as every CTA of this grid has started; it waits for
  // this grid's completion and memory flush before it reads anything
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // this grid is itself launched programmatically: its CTAs are set up while the kernel before it in the stream (the
  // producer of the logits) drains; wait for that kernel's completion and memory flush before the first load
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // the summaries are read again in a few tens of microseconds: keep them in L2 while the read-once logits stream through
  uint64_t keep;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
#ifdef TAUV_BM_EVICT1
  uint64_t first;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(first));
#endif
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
      for (int r = 0; r < kBmRows; ++r) {
#ifdef TAUV_BM_EVICT1
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=f"(x[r].x), "=f"(x[r].y), "=f"(x[r].z), "=f"(x[r].w) : "l"(p + (size_t)r * a.W), "l"(first));
#else
        x[r] = ldg_stream4(p + (size_t)r * a.W);
#endif
      }
    } else {
#pragma unroll
      for (int r = 0; r < kBmRows; ++r)
        x[r] = r < nr ? ldg_stream4(p + (size_t)r * a.W) : make_float4(TAUV_NEG_INF, TAUV_NEG_INF, TAUV_NEG_INF, TAUV_NEG_INF);
    }
#pragma unroll
    for (int r = 0; r < kBmRows; ++r) m = fmaxf(m, fmaxf(fmaxf(x[r].x, x[r].y), fmaxf(x[r].z, x[r].w)));
    st_keep(&a.bm[(size_t)b * a.G * 32 + i], m, keep);
  }
  float g = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, o));
  if ((threadIdx.x & 31) == 0 && valid) st_keep(&a.bm2[(size_t)b * a.G + (i >> 5)], g, keep);
}

constexpr int kSelCountRank = 512;    // up to this many candidates are ranked by counting, more by a bitonic sort

// Capacities of the select pass: the small tier serves k <= 256 (everything on the bench path), the large one k <= 1024.
struct SelTierSmall {
  static constexpr int kMaxK = 256, kHotCap = 1024, kCellCap = 2048, kBoxCap = 2048;
};
struct SelTierLarge {
  static constexpr int kMaxK = 1024, kHotCap = 2048, kCellCap = 4096, kBoxCap = 2048;
};

template <class TIER>
struct __align__(16) SelSharedT {
  static constexpr int kHotCap = TIER::kHotCap;    // hot groups / hot blocks per attempt
  static constexpr int kCellCap = TIER::kCellCap;  // hot cells per attempt
  static constexpr int kCandCap = TIER::kCellCap;  // candidate list (final composites): one slot per hot cell
  static constexpr int kBoxCap = TIER::kBoxCap;    // cells whose box arithmetic can be done ahead of the ranking
  static constexpr int kMaxK = TIER::kMaxK;
  static constexpr int kSmallSeg = (kCandCap - 2 * kMaxK) / 32;  // blocks per exhaustive sub-step that cannot overflow
  struct {
    unsigned long long keys[kSelThreads];  // later: the ranked keys
    uint32_t hist[kRadixBins];             // T selection: 1024 bins + per-warp columns; exhaustive path: radix histogram;
  } s;                                     // later: filler flags
  unsigned long long cand[kCandCap];       // peaks: final composites (score key << 32 | ~flat index)
  uint32_t hot[kHotCap];                   // fast path: flat index of the block's first cell; exhaustive path: block ids
  uint32_t hotpos[kHotCap];                // fast path: first row << 16 | first column
  uint32_t hgrp[kHotCap];
  uint32_t cell[kCellCap];                 // hot block ids while they are collected; then the hot cells: flat index,
  uint32_t cellpos[kCellCap];              //   row << 16 | column,
  float cellx[kCellCap];                   //   logit,
  float cellm[kCellCap];                   //   maximum of the eight neighbours
  BoxVals cellbox[kBoxCap];                //   size / offset / depth arithmetic of the cell (idle warps, while the cells are tested)
  uint32_t rankcell[kMaxK];                // the cell behind each ranked key
  uint32_t ctl[8];
  int n_hgrp, n_hot, n_xhot, n_xcell, n_cand, n_zero, first_below, flag;
  uint32_t T_key, T_key2;
  float s_T, xc_f;  // s_T: the sigmoid of the threshold T (computed by an idle warp while the cells are tested)
  unsigned long long thr_c;
};

__device__ __forceinline__ void sel_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

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

// Ordered, contention-free collection: list[...] <- the indices i < n with !(src[i] < T).  The ballots of up to eight
// rounds are kept in registers, so a warp needs ONE add on the shared counter per eight rounds (same-address
// shared-memory atomics serialise at ~20 cycles each: one per hit was most of this kernel's time).
__device__ __forceinline__ void sel_collect(const float* __restrict__ src, int n, float T_f, int* counter, uint32_t* list,
                                            int cap) {
  const int tid = threadIdx.x, lane = tid & 31;
  for (int c0 = 0; c0 < n; c0 += 8 * kSelThreads) {
    uint32_t m[8];
    int tot = 0;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      m[it] = 0u;
      if (c0 + it * kSelThreads < n) {
        const int i = c0 + it * kSelThreads + tid;
        m[it] = __ballot_sync(0xffffffffu, i < n && !(src[i < n ? i : 0] < T_f));
        tot += __popc(m[it]);
      }
    }
    int base = 0;
    if (lane == 0 && tot) base = atomicAdd(counter, tot);
    base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      if ((m[it] >> lane) & 1u) {
        const int slot = base + __popc(m[it] & ((1u << lane) - 1u));
        if (slot < cap) list[slot] = (uint32_t)(c0 + it * kSelThreads + tid);
      }
      base += __popc(m[it]);
    }
  }
}

// a cell that reached T, with the maximum of its neighbours: peak test, sigmoid, final key (0: not a peak / zero score).
// Dense: one thread per hot cell, a few warps per frame.
template <class SH>
__device__ __noinline__ unsigned long long sel_test_cell(const SH* sh, int i) {
  const float x = sh->cellx[i];
  if (!sel_is_peak(x, sh->cellm[i])) return 0ull;
  return sel_final_key(x, sh->cell[i]);
}

// Rank the candidates cand[0, n) (distinct composites; 0 = empty slot) by counting, write the min(valid, k) best in
// order.  With a threshold (T_key != 0) the thread that emits rank k - 1 also decides whether the result is complete:
// every logit x <= T has sigmoid_ref(x) <= sigmoid_ref(T) (1 + 1e-6) (three roundings on a monotone function), so nothing
// below T can reach the k-th best score s_k if sigmoid_ref(T) < s_k (1 - 2e-5) — the same guard band as
// reject_key_for_score, without its logarithm.  sh->flag = 1 if that fails or fewer than k peaks reached T.  Returns
// npos = min(valid, k).  All threads call this.
template <class SH>
__device__ __noinline__ int sel_rank_emit(const SelArgs& a, SH* sh, int b, int n, uint32_t T_key, bool have_box,
                                          const uint32_t* candcell) {
  const int tid = threadIdx.x, k = a.k;
  const int sel_rep = 0;
  (void)sel_rep;
  unsigned long long* ranked = sh->s.keys;
  if (n > kSelCountRank) {
    // long lists (k > ~450): counting would cost n^2 / 2048 shared-memory loads per thread — sort instead (zeros last)
    int p2 = 1;
    while (p2 < n) p2 <<= 1;
    for (int i = n + tid; i < p2; i += kSelThreads) sh->cand[i] = 0ull;
    __syncthreads();
    block_bitonic_sort_desc<kSelThreads>(sh->cand, p2);
    for (int i = tid; i < n; i += kSelThreads) {
      const unsigned long long c = sh->cand[i];
      if (c == 0ull) atomicAdd(&sh->n_zero, 1);
      else if (i < k) ranked[i] = c;
    }
    have_box = false;  // (the sort has moved the keys away from their cells)
  } else {
  const int tpc = n <= kSelThreads / 8 ? 8 : 4;  // threads per candidate
  const int n2 = (n + 1) >> 1;  // (the caller has zeroed cand[n] when n is odd)
  for (int i0 = 0; i0 < n; i0 += kSelThreads / tpc) {
    const int i = i0 + tid / tpc, part = tid & (tpc - 1);
    const unsigned long long my = i < n ? sh->cand[i] : 0ull;
    int cnt = 0;
    if (my != 0ull) {  // (whole warps beyond the list skip the loop)
      const ulonglong2* c2 = reinterpret_cast<const ulonglong2*>(sh->cand);  // (cand[n] = 0 when n is odd)
#pragma unroll 4
      for (int j = part; j < n2; j += tpc) {
        const ulonglong2 v = c2[j];
        cnt += (v.x > my ? 1 : 0) + (v.y > my ? 1 : 0);
      }
    }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
    if (tpc == 8) cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
    if (i < n && part == 0) {
      if (my == 0ull) atomicAdd(&sh->n_zero, 1);
      else if (cnt < k) {
        ranked[cnt] = my;
        sh->rankcell[cnt] = have_box ? candcell[i] : 0u;  // (the queued cell whose box arithmetic is in cellbox[])
      }
    }
  }
  }
  __syncthreads();
  SEL_STAMP(13);
  const int valid = n - sh->n_zero;
  const int npos = valid < k ? valid : k;
  const uint32_t hw_elems = (uint32_t)(a.H * a.W);
  if (tid < npos) {
    const unsigned long long c = ranked[tid];
    const uint32_t flat = composite_idx(c);
    const float s = key_to_float(composite_key(c));
    const uint32_t lab = fdiv_u32(flat, a.dHW);
    const uint32_t rem = flat - lab * hw_elems;
    const int iy = (int)fdiv_u32(rem, a.dW), ix = (int)(rem - (uint32_t)iy * (uint32_t)a.W);
    const long long slot = (long long)b * k + tid;
    a.out_index[slot * 2 + 0] = iy;
    a.out_index[slot * 2 + 1] = ix;
    a.out_label[slot] = lab;
    a.out_score[slot] = s;
    if (tid == k - 1 && T_key != 0u) {  // (denormal sigmoids carry no relative guard band: leave those to the exhaustive pass)
      const float s_T = sh->s_T;
      sh->flag = ((s_T == 0.0f || s_T >= 1e-30f) && s_T < __fmul_rn(s, 1.0f - 2e-5f)) ? 0 : 1;
    }
    if (a.box.enabled) {
      if (have_box) box_store(a.box, slot, sh->cellbox[sh->rankcell[tid]]);
      else box_one(a.box, b, slot, iy, ix);
      if (s < a.box.thr) atomicMin(&sh->first_below, tid);
    }
  } else if (tid == kSelThreads - 32 && T_key != 0u && valid < k) {
    sh->flag = 1;
  }
  __syncthreads();
  if (tid == 0 && a.box.enabled && npos == k) a.box.count[b] = sh->first_below;  // (npos < k: sel_finish, after the fillers)
  return npos;
}

// The zero-score tail (fewer than k positive peaks in the whole frame) and the count of leading detections.
template <class SH>
__device__ __noinline__ void sel_finish(const SelArgs& a, SH* sh, int b, int npos) {
  const int tid = threadIdx.x, k = a.k;
  if (npos < k) {
    uint32_t* flags = sh->s.hist;
    const unsigned long long* ranked = sh->s.keys;
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
template <class SH>
__device__ __noinline__ void sel_slow_step(const SelArgs& a, SH* sh, const float* __restrict__ fhm,
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
      if (slot < SH::kCandCap) sh->cand[slot] = fin;
      else sh->flag = 1;
    }
  }
  __syncthreads();
}

// exact prune of the candidate list to its k best; raises the key threshold and the logit filter
template <class SH>
__device__ __noinline__ void sel_slow_prune(const SelArgs& a, SH* sh) {
  const int tid = threadIdx.x, k = a.k;
  const int n = sh->n_cand;  // > k
  const unsigned long long T =
      block_kth_largest<kSelThreads>([&](int i) { return sh->cand[i]; }, n, k, sh->s.hist, sh->ctl);
  unsigned long long mine[SH::kCandCap / kSelThreads];
#pragma unroll
  for (int u = 0; u < SH::kCandCap / kSelThreads; ++u) {
    const int i = tid + u * kSelThreads;
    mine[u] = i < n ? sh->cand[i] : 0ull;
  }
  if (tid == 0) sh->n_cand = 0;
  __syncthreads();
#pragma unroll
  for (int u = 0; u < SH::kCandCap / kSelThreads; ++u)
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
template <class SH>
__device__ __noinline__ int sel_slow(const SelArgs& a, SH* sh, const float* __restrict__ fhm,
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
      for (int s = seg; s < seg + len; s += SH::kSmallSeg) {
        int l2 = SH::kSmallSeg;
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

// Large k (K1 > 1024): four strided maxima per thread (4096 in all); leaves every warp's largest key and its
// ceil(K1 / 32)-th largest in the bracket slots of the histogram array and returns this thread's four keys.
template <class SH>
__device__ __noinline__ uint4 sel_bracket_wide(SH* sh, const float* __restrict__ lv, int nlv, int K1) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* hist = sh->s.hist;
  float tm[4] = {TAUV_NEG_INF, TAUV_NEG_INF, TAUV_NEG_INF, TAUV_NEG_INF};
  for (int i = tid; i < nlv; i += 4 * kSelThreads) {
#pragma unroll
    for (int v = 0; v < 4; ++v)
      if (i + v * kSelThreads < nlv) tm[v] = fmaxf(tm[v], lv[i + v * kSelThreads]);
  }
  const uint4 key = make_uint4(float_to_key(tm[0]), float_to_key(tm[1]), float_to_key(tm[2]), float_to_key(tm[3]));
  const int j = (K1 + 31) >> 5;
  uint32_t c0 = key.x, c1 = key.y, c2 = key.z, c3 = key.w;
  uint32_t local = max(max(c0, c1), max(c2, c3));
  uint32_t mx = __reduce_max_sync(0xffffffffu, local);
  if (lane == 0) hist[1088 + warp] = mx;
  for (int jj = 1; jj < j; ++jj) {
    const unsigned bal = __ballot_sync(0xffffffffu, local == mx);
    if (lane == __ffs(bal) - 1) {  // remove ONE instance of the maximum
      if (c0 == mx) c0 = 0u;
      else if (c1 == mx) c1 = 0u;
      else if (c2 == mx) c2 = 0u;
      else c3 = 0u;
      local = max(max(c0, c1), max(c2, c3));
    }
    mx = __reduce_max_sync(0xffffffffu, local);
  }
  if (lane == 0) hist[1056 + warp] = mx;
  return key;
}

template <class TIER>
__global__ void __launch_bounds__(kSelThreads, 1) select_kernel(const __grid_constant__ SelArgs a) {
  using SH = SelSharedT<TIER>;
  extern __shared__ __align__(128) unsigned char sel_smem[];
  SH* sh = reinterpret_cast<SH*>(sel_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const int k = a.k;
  const int HW = a.H * a.W;
  const float* __restrict__ fhm = a.hm + (size_t)b * a.C * HW;
  const float* __restrict__ bm = a.bm + (size_t)b * a.G * 32;  // (rows padded to whole groups: 128-byte aligned)
  const float* __restrict__ bm2 = a.bm2 + (size_t)b * a.G;
  // Every instruction of a 1024-thread CTA costs >= 8 issue cycles (~15 measured): the phases are written for
  // instruction count, not for the number of round trips to memory (HBM latency is ~0.4 us; 250 instructions are 2 us).
  int sel_rep = 0;  // (-DTAUV_SEL_TWICE experiment: the whole pass a second time, with warm caches, traced separately)
  (void)sel_rep;
  if (tid == 0) sh->first_below = k;
  SEL_STAMP(0);
  // The CTA is resident about a microsecond before block_max_kernel completes: use it to take the TLB misses of every
  // region this pass will touch (the first access of an SM to a page costs ~1 us, and each phase below starts with
  // one): one load per 64 KB of the frame's logits, of its summaries and of the size / offset / depth planes.  L2-only
  // loads (.cg): nothing stale can stay in L1, and the values are discarded.
  {
    auto touch = [](const void* p) {
      uint32_t x;
      asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(x) : "l"(p) : "memory");
    };
    const size_t frame_floats = (size_t)a.C * HW;
    for (size_t off = (size_t)tid * 16384; off < frame_floats; off += (size_t)kSelThreads * 16384) touch(fhm + off);
    if (tid == kSelThreads - 1) touch(fhm + frame_floats - 1);
    if (tid >= 32 && tid < 64)
      for (int off = (tid - 32) * 16384; off < a.n_blk; off += 32 * 16384) touch(bm + off);
    if (tid == 64) touch(bm + a.n_blk - 1);
    if (tid == 65) touch(bm2);
    if (tid == 66) touch(bm2 + a.G - 1);
    if (a.box.enabled && tid >= 96 && tid < 96 + 16) {
      const int q = tid - 96, iy = (int)((long long)(a.H - 1) * (q >> 1) / 7), ix = (q & 1) ? a.W - 1 : 0;
      touch(a.box.size + b * a.box.ss[0] + iy * a.box.ss[1] + ix * a.box.ss[2]);
      touch(a.box.size + b * a.box.ss[0] + iy * a.box.ss[1] + ix * a.box.ss[2] + a.box.ss[3]);
      if (a.box.mode == TAUV_BOX_DECODE) {
        touch(a.box.offset + b * a.box.os[0] + iy * a.box.os[1] + ix * a.box.os[2]);
        touch(a.box.offset + b * a.box.os[0] + iy * a.box.os[1] + ix * a.box.os[2] + a.box.os[3]);
      }
      if (a.box.depth != nullptr) touch(a.box.depth + b * a.box.ds[0] + iy * a.box.ds[1] + ix * a.box.ds[2]);
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");  // block_max_kernel has completed and flushed (no-op without the attribute)
  SEL_STAMP(1);

  int K1 = k + (k >> 3) + 8;  // (a retry doubles the frame's time, and the launch waits for its slowest frame)
  int n = 0;
  bool slow = false;
  int attempts = 0;
  bool take_T2 = false;  // the next attempt takes sh->T_key2 as its threshold (step 2: smooth maps)
  for (;;) {
    ++attempts;
    // ---- 1. threshold T: (a lower bound of) the K1-th largest of 1024 strided maxima of the group (or block) maxima.
    //         L = the smallest of the (full) warps' j-th largest maxima (>= K1 keys reach it),
    //         U = the largest key; one 1024-bin histogram of the keys in [L, U] and a suffix count give the bin in which
    //         the count reaches K1; T = its lower edge. ----
    const bool use_grp = a.G >= 2 * K1;  // (at G = 2 K1 about 1.4 K1 blocks reach T: still tight, and 32x fewer values to look at)
    const float* __restrict__ lv = use_grp ? bm2 : bm;
    const int nlv = use_grp ? a.G : a.n_blk;
    uint32_t T_key = 0u;
    float T_f = TAUV_NEG_INF;
    __syncthreads();
    // (frames with fewer than 1024 values at this level fill only the first nwf warps: the bracket L comes from those —
    // a warp of -inf padding would drag L, and with it the histogram's resolution, down to nothing)
    const int nwf = (nlv < kSelThreads ? nlv : kSelThreads) >> 5;  // warps whose 32 keys are all real
    // large k (K1 > 1024): four strided maxima per thread, 4096 in all
    const bool wide = K1 > 32 * nwf && nlv >= 4 * kSelThreads && K1 <= 4 * kSelThreads;
    // The first attempt also takes the threshold for 2 K1 from the same histogram (T2 <= T): on a smooth map — hot
    // regions that span many blocks, seen in step 2 as several hot blocks per hot group — K1 groups hold fewer than
    // k peaks more often than not, and a whole second attempt doubles the frame's time.
    const int K2 = (attempts == 1 && use_grp && !wide && 2 * K1 <= 32 * nwf && a.G >= 4 * K1) ? 2 * K1 : K1;
    uint32_t T_key2 = 0u;
    uint32_t key = 0u;
    uint4 wkey = make_uint4(0u, 0u, 0u, 0u);
    if (take_T2) {
      T_key = sh->T_key2;
      T_f = key_to_float(T_key);
      take_T2 = false;
    } else if (nwf > 0 && (K1 <= 32 * nwf || wide)) {
      uint32_t* hist = sh->s.hist;  // [0,1024) bins, [1024,1056) warp totals, [1056,1088) j-th largest, [1088,1120) largest
      if (wide) {
        wkey = sel_bracket_wide(sh, lv, nlv, K1);  // (out of line: the common case below stays as lean as it was)
      } else {
        float tm = TAUV_NEG_INF;
        for (int i = tid; i < nlv; i += kSelThreads) tm = fmaxf(tm, lv[i]);
        key = float_to_key(tm);
        // the warp's largest key and its j-th largest (with multiplicity), j = ceil(K1 / nwf): j warp-wide max
        // reductions; every full warp holds at least j keys at or above its own j-th largest, so >= K1 keys reach L
        const int j = (K1 + nwf - 1) / nwf;
        uint32_t cur = key, mx = __reduce_max_sync(0xffffffffu, cur);
        if (lane == 0) hist[1088 + warp] = mx;
        if (warp < nwf) {
          for (int jj = 1; jj < j; ++jj) {
            const unsigned bal = __ballot_sync(0xffffffffu, cur == mx);
            if (lane == __ffs(bal) - 1) cur = 0u;
            mx = __reduce_max_sync(0xffffffffu, cur);
          }
        } else {
          mx = 0xffffffffu;  // (neutral for the minimum)
        }
        if (lane == 0) hist[1056 + warp] = mx;
      }
      hist[tid] = 0u;
      __syncthreads();
      uint32_t Lk = hist[1056 + lane], Uk = hist[1088 + lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Lk = min(Lk, __shfl_xor_sync(0xffffffffu, Lk, o));
        Uk = max(Uk, __shfl_xor_sync(0xffffffffu, Uk, o));
      }
      const uint32_t range = Uk - Lk;
      const int shft = range >= 1024u ? 22 - __clz(range) : 0;  // (range >> shft) < 1024
      if (!wide) {
        if (key >= Lk) atomicAdd(&hist[(key - Lk) >> shft], 1u);
      } else {
        if (wkey.x >= Lk) atomicAdd(&hist[(wkey.x - Lk) >> shft], 1u);
        if (wkey.y >= Lk) atomicAdd(&hist[(wkey.y - Lk) >> shft], 1u);
        if (wkey.z >= Lk) atomicAdd(&hist[(wkey.z - Lk) >> shft], 1u);
        if (wkey.w >= Lk) atomicAdd(&hist[(wkey.w - Lk) >> shft], 1u);
      }
      __syncthreads();
      const uint32_t c = hist[tid];
      uint32_t suf = c;  // inclusive suffix count inside the warp (lane 31 = highest bin)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_down_sync(0xffffffffu, suf, o);
        if (lane + o < 32) suf += v;
      }
      if (lane == 0) hist[1024 + warp] = suf;
      __syncthreads();
      uint32_t wsuf = hist[1024 + lane];  // suffix over the warps' totals
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_down_sync(0xffffffffu, wsuf, o);
        if (lane + o < 32) wsuf += v;
      }
      const uint32_t above = __shfl_sync(0xffffffffu, wsuf, (warp + 1) & 31);
      suf += warp < 31 ? above : 0u;
      if (suf >= (uint32_t)K1 && suf - c < (uint32_t)K1) sh->T_key = Lk + ((uint32_t)tid << shft);
      // (T2 from the same bracket: the bin in which the count reaches 2 K1, or L itself when fewer keys than that lie
      // in [L, U] — the bracket guarantees K1 of them and usually holds two or three times as many)
      if (K2 != K1 && ((suf >= (uint32_t)K2 && suf - c < (uint32_t)K2) || (tid == 0 && suf < (uint32_t)K2)))
        sh->T_key2 = Lk + ((uint32_t)tid << shft);
      __syncthreads();
      T_key = sh->T_key;
      T_f = key_to_float(T_key);
      if (!(T_f > TAUV_NEG_INF)) T_key = 0u;
      if (K2 != K1 && T_key != 0u) {
        T_key2 = sh->T_key2;
        if (!(T_key2 < T_key) || !(key_to_float(T_key2) > TAUV_NEG_INF)) T_key2 = 0u;
      }
    }
    if (tid == 0) {
      sh->n_hgrp = 0;
      sh->n_hot = 0;
      sh->n_xhot = 0;
      sh->n_xcell = 0;
      sh->n_cand = 0;
      sh->n_zero = 0;
      sh->flag = 0;
    }
    __syncthreads();
    SEL_STAMP(2);
    // ---- 2. hot groups -> hot blocks (ids in sh->cell), then their positions ----
    int nhot;
    if (use_grp) {
      sel_collect(bm2, a.G, T_f, &sh->n_hgrp, sh->hgrp, SH::kHotCap);
      __syncthreads();
      const int nh = sh->n_hgrp;
      if (nh > SH::kHotCap) { slow = true; break; }
      // eight lanes per hot group, one 128-bit load each.  The group's first hot block (there is one: the group
      // maximum is a block maximum) takes the group's own slot, further ones (rare) are appended behind the nh slots.
      const int ne = nh * 8;
      for (int e0 = 0; e0 < ne; e0 += 2 * kSelThreads) {
        float4 v[2];
        int idx[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int e = e0 + u * kSelThreads + tid;
          idx[u] = e < ne ? (int)sh->hgrp[e >> 3] * 32 + (lane & 7) * 4 : a.n_blk;
          v[u] = idx[u] < a.n_blk ? *reinterpret_cast<const float4*>(bm + idx[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          unsigned hm = 0u;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (idx[u] + c < a.n_blk && !(vv[c] < T_f)) hm |= 1u << c;
          const unsigned bal = __ballot_sync(0xffffffffu, hm != 0u);
          const unsigned seg = (bal >> (lane & 24)) & 0xffu;         // the eight lanes of this group
          const int gs = (e0 + u * kSelThreads + tid) >> 3;
          if (seg == 0u && (lane & 7) == 0 && idx[u] < a.n_blk) sh->cell[gs] = (uint32_t)idx[u];  // (cannot happen)
          if (hm) {
            bool first = (lane & 7) == __ffs(seg) - 1;                // the group's first lane with a hot block
            do {
              const int c = __ffs(hm) - 1;
              hm &= hm - 1;
              int slot = gs;
              if (!first) slot = nh + atomicAdd(&sh->n_xhot, 1);
              first = false;
              if (slot < SH::kCellCap) sh->cell[slot] = (uint32_t)(idx[u] + c);
            } while (hm);
          }
        }
      }
      __syncthreads();
      nhot = nh + sh->n_xhot;
      if (T_key2 != 0u && 2 * nhot >= 3 * nh) {  // smooth map (noise: 1.07 hot blocks per hot group): step 2 again, with T2
        K1 = K2;
        take_T2 = true;
        continue;
      }
    } else {
      sel_collect(bm, a.n_blk, T_f, &sh->n_hot, sh->cell, SH::kCellCap);
      __syncthreads();
      nhot = sh->n_hot;
    }
    if (nhot > SH::kHotCap) { slow = true; break; }
    for (int i = tid; i < nhot; i += kSelThreads) {  // (dense: a handful of warps)
      const uint32_t blk = sh->cell[i];
      const uint32_t q = fdiv_u32(blk, a.dW4), c4 = blk - q * (uint32_t)a.W4;
      const uint32_t c = fdiv_u32(q, a.dNrg), rg = q - c * (uint32_t)a.n_rg;
      const uint32_t r0 = rg * kBmRows, col0 = c4 * 4;
      sh->hot[i] = c * (uint32_t)HW + r0 * (uint32_t)a.W + col0;
      sh->hotpos[i] = (r0 << 16) | col0;
    }
    __syncthreads();
    SEL_STAMP(3);
    // ---- 3a. the rows of the hot blocks: eight lanes per block, one 128-bit load each; cells that reach T are
    //          queued — the block's first one (there is one: the block maximum) in the block's own slot, further
    //          ones (7 % of the blocks on noise) behind the nhot slots ----
    {
      const int ne = nhot * 8;
      const bool T_ge80 = T_f >= -80.0f;
      for (int e0 = 0; e0 < ne; e0 += 2 * kSelThreads) {
        float4 v[2];
        uint32_t org[2], pos[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int e = e0 + u * kSelThreads + tid;
          org[u] = 0xffffffffu;
          pos[u] = 0u;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e < ne) {
            const uint32_t p = sh->hotpos[e >> 3];
            const int r = (int)(p >> 16) + (lane & 7);
            if (r < a.H) {
              org[u] = sh->hot[e >> 3] + (uint32_t)((lane & 7) * a.W);
              pos[u] = ((uint32_t)r << 16) | (p & 0xffffu);
              v[u] = *reinterpret_cast<const float4*>(fhm + org[u]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          // A cell clearly below one of its neighbours INSIDE the block is no peak, whatever lies outside: leave it
          // out of the queue.  On noise this changes nothing (a block holds one or two cells >= T); on the smooth
          // maps of a trained head, whose hot regions span whole blocks, it cuts the queue from ~7 cells per hot
          // block to about one.  ("clearly": by the margin below which sel_is_peak still checks for a sigmoid tie.)
          const bool live = org[u] != 0xffffffffu;
          unsigned hm = 0u;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (live && !(vv[c] < T_f)) hm |= 1u << c;  // (NaN cells pass here and fail the peak test)
          // A block with ONE cell >= T (its maximum) needs no filter: on noise that is 93 % of the blocks, and three
          // warps in four skip the shuffles below (a warp covers four blocks).
          const int n_blocks_here = __popc(__ballot_sync(0xffffffffu, (lane & 7) == 0 && e0 + u * kSelThreads + tid < ne));
          if (__reduce_add_sync(0xffffffffu, (unsigned)__popc(hm)) > (unsigned)n_blocks_here) {
            const float w0 = live ? vv[0] : TAUV_NEG_INF, w1 = live ? vv[1] : TAUV_NEG_INF;
            const float w2 = live ? vv[2] : TAUV_NEG_INF, w3 = live ? vv[3] : TAUV_NEG_INF;
            // maxima over the columns c-1..c+1 of this row, then over the rows above and below (eight lanes = the eight
            // rows of a block; a shuffle of width 8 hands the first / last row its own value back, which is neutral)
            const float h0 = fmaxf(w0, w1), h3 = fmaxf(w2, w3), h1 = fmaxf(h0, w2), h2 = fmaxf(w1, h3);
            const float nm[4] = {
                fmaxf(h0, fmaxf(__shfl_up_sync(0xffffffffu, h0, 1, 8), __shfl_down_sync(0xffffffffu, h0, 1, 8))),
                fmaxf(h1, fmaxf(__shfl_up_sync(0xffffffffu, h1, 1, 8), __shfl_down_sync(0xffffffffu, h1, 1, 8))),
                fmaxf(h2, fmaxf(__shfl_up_sync(0xffffffffu, h2, 1, 8), __shfl_down_sync(0xffffffffu, h2, 1, 8))),
                fmaxf(h3, fmaxf(__shfl_up_sync(0xffffffffu, h3, 1, 8), __shfl_down_sync(0xffffffffu, h3, 1, 8)))};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              // (nm >= the cell itself >= T: with T >= -80 the third condition of sel_is_peak's margin holds by itself)
              const bool below = (nm[c] - vv[c]) >= 1e-3f && vv[c] <= 4.0f && (T_ge80 || nm[c] >= -80.0f);  // (false for NaN)
              if (below) hm &= ~(1u << c);
            }
          }
          const unsigned bal = __ballot_sync(0xffffffffu, hm != 0u);
          const unsigned seg = (bal >> (lane & 24)) & 0xffu;
          const int bs = (e0 + u * kSelThreads + tid) >> 3;
          if (seg == 0u && (lane & 7) == 0 && e0 + u * kSelThreads + tid < ne) {  // (cannot happen; keeps the slot defined)
            sh->cellpos[bs] = sh->hotpos[bs];
            sh->cellx[bs] = TAUV_NEG_INF;
            sh->cell[bs] = sh->hot[bs];
          }
          if (hm) {
            bool first = (lane & 7) == __ffs(seg) - 1;
            do {
              const int c = __ffs(hm) - 1;
              hm &= hm - 1;
              int slot = bs;
              if (!first) slot = nhot + atomicAdd(&sh->n_xcell, 1);
              first = false;
              if (slot < SH::kCellCap) {
                sh->cellpos[slot] = pos[u] + (uint32_t)c;
                sh->cellx[slot] = c == 0 ? v[u].x : c == 1 ? v[u].y : c == 2 ? v[u].z : v[u].w;
                sh->cell[slot] = org[u] + (uint32_t)c;  // (flat index; the block ids were consumed before the barrier)
              }
            } while (hm);
          }
        }
      }
    }
    __syncthreads();
    n = nhot + sh->n_xcell;
    if (n > SH::kCellCap) { slow = true; break; }
    SEL_STAMP(11);
    // ---- 3b. the eight neighbours of every queued cell: eight lanes per cell, one load each ----
    {
      const int ne = n * 8;
      for (int e0 = 0; e0 < ne; e0 += 2 * kSelThreads) {
        float v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int e = e0 + u * kSelThreads + tid;
          v[u] = TAUV_NEG_INF;
          if (e < ne) {
            const uint32_t p = sh->cellpos[e >> 3];
            const int part = lane & 7, pp = part + (part >= 4 ? 1 : 0);  // 0..8 without the centre
            const int dy = pp / 3 - 1, dx = pp - (pp / 3) * 3 - 1;
            const int rr = (int)(p >> 16) + dy, cc = (int)(p & 0xffffu) + dx;
            if (rr >= 0 && rr < a.H && cc >= 0 && cc < a.W) v[u] = fhm[(long long)sh->cell[e >> 3] + dy * a.W + dx];
          }
        }
        // (while the loads are in flight: the rank loop of step 4 reads the candidates in pairs, so an odd list must end
        // in an empty slot — clear the list the peaks of 3c will be appended to)
        if (e0 == 0)
          for (int i = tid; i < SH::kCandCap; i += kSelThreads) sh->cand[i] = 0ull;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float m = v[u];
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
          const int e = e0 + u * kSelThreads + tid;
          if (e < ne && (lane & 7) == 0) sh->cellm[e >> 3] = m;
        }
      }
    }
    __syncthreads();
    SEL_STAMP(12);
    // ---- 3c. one thread per queued cell: peak test, sigmoid, final key into the cell's candidate slot ----
    //          (warps 0-7), while the other warps do the size / offset / depth arithmetic of every queued cell — it
    //          would otherwise sit, with its gathers and two fp64 divisions, at the very end of the pass
    //          Only the peaks enter the candidate list (warp-aggregated append; candcell[] remembers the cell behind
    //          each): on smooth maps most queued cells are not peaks, and ranking costs the square of the list.
    static_assert(offsetof(SH, hotpos) == offsetof(SH, hot) + sizeof(uint32_t) * SH::kHotCap &&
                      2 * SH::kHotCap >= SH::kCellCap, "candcell[] spans hot[] and hotpos[]");
    uint32_t* candcell = sh->hot;  // (hot[] + hotpos[], both dead since 3a: kCellCap entries)
    if (tid < 256) {
      for (int i0 = 0; i0 < n; i0 += 256) {
        const int i = i0 + tid;
        const unsigned long long ck = i < n ? sel_test_cell(sh, i) : 0ull;
        const int slot = sel_append(&sh->n_cand, ck != 0ull);
        if (slot >= 0) {
          sh->cand[slot] = ck;
          candcell[slot] = (uint32_t)i;
        }
      }
      if (tid == 255) sh->s_T = sigmoid_ref(T_f);  // (for the completeness test of sel_rank_emit)
    } else if (a.box.enabled && n <= SH::kBoxCap) {
      for (int i = tid - 256; i < n; i += kSelThreads - 256) {
        const uint32_t pos = sh->cellpos[i];
        sh->cellbox[i] = box_values(a.box, b, (int)(pos >> 16), (int)(pos & 0xffffu));
      }
    }
    __syncthreads();
    SEL_STAMP(4);
    SEL_NOTE(8, attempts);
    SEL_NOTE(9, nhot);
    SEL_NOTE(10, n);
    const int ncand = sh->n_cand;
    SEL_NOTE(14, ncand);
    // ---- 4. rank, emit, and check that enough peaks reached T and nothing below T could have made it ----
    const int npos = sel_rank_emit(a, sh, b, ncand, T_key, n <= SH::kBoxCap, candcell);
    SEL_STAMP(5);
    if (T_key != 0u && sh->flag) {
      // Not enough peaks at or above T, or the k-th best too close to it: lower T.  The number of blocks that must be
      // examined per peak found is a property of the map (1.07 on noise, 5-10 on the smooth maps of a trained head,
      // whose hot regions span many blocks), so the next K1 aims at 1.5 k peaks at the ratio just measured — at most
      // what the hot-block list holds, once; if that is not enough either, the frame is done exhaustively.
      const int valid = ncand - sh->n_zero > 0 ? ncand - sh->n_zero : 1;
      __syncthreads();
      if (tid == 0) sh->first_below = k;
      long long next = (long long)K1 * k * 3 / (2 * valid) + 16;
      if (next < K1 + (K1 >> 1)) next = K1 + (K1 >> 1);
      if (next > (1 << 20)) next = 1 << 20;
      const int lim = SH::kHotCap - SH::kHotCap / 8;
      if (K1 < lim && next > lim) next = lim;
      K1 = (int)next;
      continue;
    }
    if (npos < k) sel_finish(a, sh, b, npos);
    SEL_STAMP(6);
#ifdef TAUV_SEL_TWICE
    if (sel_rep == 0) {
      sel_rep = 1;
      K1 = k + (k >> 3) + 8;
      attempts = 0;
      __syncthreads();
      if (tid == 0) sh->first_below = k;
      SEL_STAMP(1);
      continue;
    }
#endif
    return;
  }
  SEL_NOTE(8, -attempts);
  if (slow) {
    __syncthreads();
    if (tid == 0) {
      sh->first_below = k;
      sh->n_zero = 0;
    }
    n = sel_slow(a, sh, fhm, bm);
    if (tid == 0 && (n & 1) && n < SH::kCandCap) sh->cand[n] = 0ull;
    __syncthreads();
    const int npos = sel_rank_emit(a, sh, b, n, 0u, false, nullptr);
    if (npos < k) sel_finish(a, sh, b, npos);
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
  if (n_blk >= (1LL << 26) || (long long)C * H * W >= (1LL << 31)) return false;  // (FastDiv: dividends below 2^31)
  // a frame with fewer blocks than a few times K1 = k + k/8 + 8 has no selective threshold to find: the pass would go
  // straight to its exhaustive segments, which round 1's kernels do better (C = 1, 128 x 128, k = 1000: 81 vs 67 us)
  if (n_blk < 4LL * (k + k / 8 + 8) && (long long)k * 4 > 256) return false;
  p->W4 = W / 4;
  p->n_rg = (int)n_rg;
  p->n_blk = (int)n_blk;
  p->G = (int)((n_blk + 31) / 32);
  p->bm_bytes = align_up((size_t)B * p->G * 32 * 4, 256);
  p->bm2_bytes = align_up((size_t)B * p->G * 4, 256);
  return true;
}

// launch 1 of 2: the block maxima of every frame into the workspace (the only pass over the logits)
static int run_block_maxima(const float* hm, int B, int C, int H, int W, void* ws, size_t ws_bytes, const SelPlan& p,
                            cudaStream_t st) {
  TAUV_REQUIRE(ws != nullptr && (uintptr_t)ws % 256 == 0, TAUV_E_WORKSPACE, "workspace must be 256-byte aligned");
  TAUV_REQUIRE(ws_bytes >= p.bm_bytes + p.bm2_bytes, TAUV_E_WORKSPACE, "workspace %zu < required %zu", ws_bytes,
               p.bm_bytes + p.bm2_bytes);
  BmArgs ba;
  ba.hm = hm;
  ba.bm = reinterpret_cast<float*>(ws);
  ba.bm2 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ws) + p.bm_bytes);
  ba.C = C; ba.H = H; ba.W = W; ba.W4 = p.W4; ba.n_rg = p.n_rg; ba.n_blk = p.n_blk; ba.G = p.G;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((p.n_blk + kBmThreads - 1) / kBmThreads), (unsigned)B);
  cfg.blockDim = dim3(kBmThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TAUV_CUDA(cudaLaunchKernelEx(&cfg, block_max_kernel, ba));
  TAUV_LAUNCH_CHECK("block_max_kernel");
  return 0;
}

static int run_select_decode(const float* hm, int B, int C, int H, int W, int k, int64_t* index, int64_t* label,
                             float* score, const BoxArgs& box, void* ws, size_t ws_bytes, const SelPlan& p,
                             cudaStream_t st) {
  if (int e = run_block_maxima(hm, B, C, H, W, ws, ws_bytes, p, st)) return e;
  SelArgs sa;
  sa.hm = hm;
  sa.bm = reinterpret_cast<const float*>(ws);
  sa.bm2 = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(ws) + p.bm_bytes);
  sa.C = C; sa.H = H; sa.W = W; sa.k = k; sa.W4 = p.W4; sa.n_rg = p.n_rg; sa.n_blk = p.n_blk; sa.G = p.G;
  sa.out_index = index; sa.out_label = label; sa.out_score = score;
  sa.box = box;
  sa.dW4 = make_fastdiv((uint32_t)p.W4);
  sa.dNrg = make_fastdiv((uint32_t)p.n_rg);
  sa.dHW = make_fastdiv((uint32_t)(H * W));
  sa.dW = make_fastdiv((uint32_t)W);
#ifdef TAUV_DEBUG
  sa.trace = g_sel_trace;
#else
  sa.trace = nullptr;
#endif
  // (the small tier's 144 KB lets the CTAs be set up next to block_max_kernel's last wave; the large tier's 204 KB
  // does not — measured: 4 us on the whole call — so only k > 256 pays for it)
  const bool large = k > SelTierSmall::kMaxK;
  const void* kern = large ? (const void*)select_kernel<SelTierLarge> : (const void*)select_kernel<SelTierSmall>;
  const size_t sel_smem_bytes = large ? sizeof(SelSharedT<SelTierLarge>) : sizeof(SelSharedT<SelTierSmall>);
  TAUV_CUDA(ensure_dynamic_smem(kern, sel_smem_bytes));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)B);
  cfg.blockDim = dim3(kSelThreads);
  cfg.dynamicSmemBytes = sel_smem_bytes;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (large) TAUV_CUDA(cudaLaunchKernelEx(&cfg, select_kernel<SelTierLarge>, sa));
  else TAUV_CUDA(cudaLaunchKernelEx(&cfg, select_kernel<SelTierSmall>, sa));
  TAUV_LAUNCH_CHECK("select_kernel");
  return 0;
}

}  // namespace tauv
