// centernet_decode.cu — CenterNet head decode on sm_100a.
//
// Replaces (reference file:line under src/tauv_vision/centernet/model/):
//   decode.py:182      F.sigmoid(prediction.heatmap)
//   decode.py:239-252  heatmap_nms   (3x3 max-pool == test)
//   decode.py:255-279  heatmap_detect (joint top-k over C*H*W per frame)
//   decode.py:204-234  per-detection gather + box arithmetic
//
// Pipeline (2 launches for a whole batch):
//   K1 tile_topk_kernel : one CTA per "item" (a band of rows of one (frame, class) plane).  Every
//      thread streams 128-bit loads straight from HBM (measured on B200: 5.46 TB/s read-only for
//      plain vector loads vs 4.9 TB/s for a cp.async.bulk shared-memory ring, tools/ring_bench.cu)
//      and compares each strip with the frame's published rejection threshold; only strips that
//      pass get the 3x3 peak test (on the logits: sigmoid is monotone), the sigmoid is evaluated
//      only for survivors, and the item's top-k candidates (64-bit composite keys) go to a small
//      global candidate table and into a per-frame histogram from which the threshold is
//      republished.  The logits are read from HBM exactly once; nothing dense is written.
//   K2 merge_kernel     : one CTA per frame selects the frame's top-k from its items' candidates
//      (radix select + bitonic sort), fills zero-score slots like a dense stable top-k would,
//      and (optionally) gathers size/offset/depth and does the box arithmetic.
#include "common.cuh"

#include <cooperative_groups.h>

#include <atomic>
#include <mutex>

namespace cg = cooperative_groups;

namespace tauv {

constexpr int kTileThreads = 256;
constexpr int kItemElems = 16384;   // target elements per item (one CTA, 16 float4 loads per thread)
constexpr int kMergeThreads = 1024;
constexpr int kMaxK = 4096;
constexpr int kFrameBins = 4096;    // per-frame candidate histogram: top 12 bits of the order-preserving key
constexpr int kFrameStateWords = kFrameBins + 4;  // [0] published reject key, [1] highest occupied bin, [2] candidates so far

struct TopkPlan {
  int rows_per_item;
  int items_per_plane;
  int items_per_frame;  // C * items_per_plane
  int rows_per_frame;   // rows of the candidate table per frame: max(items_per_frame, cluster size)
  int cap;              // candidate-list capacity (entries)
  int soft;             // overflow-safe path: prune when the list grows beyond this
  int sub_elems;        // overflow-safe path: elements per sub-step (cap - soft)
  int vec;              // 1: 128-bit loads, 0: scalar loads (W % 4 != 0 or unaligned base)
  size_t smem_bytes;
  size_t cand_bytes, count_bytes, state_bytes;
};

static int make_plan(int B, int C, int H, int W, int k, const void* ptr, TopkPlan* p) {
  p->vec = (W % 4 == 0) && ((uintptr_t)ptr % 16 == 0);
  // item: a band of rows of one (frame, class) plane, about kItemElems elements, but enough items to fill the machine
  int rows_item = (kItemElems + W - 1) / W;
  if (rows_item < 1) rows_item = 1;
  const long long planes = (long long)B * C;
  const int sms = num_sms();
  // (two items per SM are plenty: the cluster kernel deals a unit's items to 8 CTAs and wants several rounds per item —
  // at batch 1 a finer split only made thousands of one-row items and a heavy merge)
  while (rows_item > 1 && planes * ((H + rows_item - 1) / rows_item) < 2LL * sms) rows_item = (rows_item + 1) / 2;
  if (rows_item > H) rows_item = H;
  p->rows_per_item = rows_item;
  p->items_per_plane = (H + rows_item - 1) / rows_item;
  p->items_per_frame = C * p->items_per_plane;
  // An item's peaks (about 1/9 of its cells on noise, far fewer once the frame has published a threshold) go to a
  // shared-memory list.  If they do not fit (plateaus, dense data without a threshold) the item is redone in
  // sub-steps of cap - soft elements with an exact prune to the top-k whenever the list passes soft = 2k.
  p->soft = 2 * k;
  p->cap = 4 * k > 2560 ? 4 * k : 2560;
  p->sub_elems = ((p->cap - p->soft) / 4) * 4;
  p->smem_bytes = (size_t)p->cap * 8 + kRadixBins * 4;
  // one row per item (scalar path) or per (unit, CTA) (cluster path: at most max(items, 8) per frame)
  p->rows_per_frame = p->items_per_frame > 8 ? p->items_per_frame : 8;
  p->cand_bytes = align_up((size_t)B * p->rows_per_frame * (size_t)k * 8, 256);
  p->count_bytes = align_up((size_t)B * p->rows_per_frame * 4, 256);
  p->state_bytes = align_up((size_t)B * kFrameStateWords * 4, 256);
  return 0;
}

// ----------------------------------------------------------------------------------------------
// K1
// ----------------------------------------------------------------------------------------------
// per-detection gather + box arithmetic of decode()/decode_keypoints() (optional tail of the ranked output)
struct BoxArgs {
  int enabled;
  const float* size; long long ss[4];
  const float* offset; long long os[4];
  const float* depth; long long ds[3];
  int mode, ratio, in_h, in_w, out_h, out_w;
  float thr;
  double* yx; float* hw; float* depth_out; int* count;
};

struct TileArgs {
  const float* hm;
  int B, C, H, W, k;
  int rows_per_item, items_per_plane, rows_per_frame;
  int cap, soft, sub_elems;
  unsigned long long* cand;  // [B*items_per_frame][k]
  int* cand_count;           // [B*items_per_frame]
  uint32_t* frame_state;     // [B][kFrameStateWords], zeroed before the launch
  // fused tail (cluster kernel, whole-frame units): the cluster selects the frame's top-k itself and writes the ranked
  // outputs, so no candidate table and no merge launch are needed
  int fuse;
  int64_t* out_index;
  int64_t* out_label;
  float* out_score;
  BoxArgs box;
  long long* trace;          // debug: per item {t_start, t_boot, t_scan, t_end, n_list, thr_key_at_start, 0, 0} or NULL
};

// x < m can still tie after the sigmoid (saturation, or a sub-ulp gap): the reference compares sigmoid values
// (decode.py:252), so those rare cases are decided on the sigmoids themselves.
__device__ __noinline__ bool sigmoid_tie(float x, float m) { return sigmoid_ref(x) == sigmoid_ref(m); }

// A logit x_c such that every x < x_c has sigmoid(x) strictly below the score s_k, with a relative guard band that
// covers the last-bit wobble of expf / logf.  Returns the order-preserving key of x_c, or 0 when no such logit is
// found cheaply (saturated scores) — then nothing is rejected up front and the exact selection alone decides.
__device__ __noinline__ uint32_t reject_key_for_score(float s_k) {
  if (!(s_k > 0.0f) || !(s_k < 1.0f)) return 0u;
  const float x0 = logf(__fdiv_rn(s_k, __fsub_rn(1.0f, s_k)));
  float margin = 1e-3f * fmaxf(1.0f, fabsf(x0));
  for (int t = 0; t < 4; ++t, margin *= 8.0f) {
    const float xc = x0 - margin;
    if (sigmoid_ref(xc) < s_k * (1.0f - 2e-5f)) return float_to_key(xc);
  }
  return 0u;
}

// Histogram bin of a final sort key.  RAW: top bits of the value key.  SIGMOID_PEAK: top bits of the key of the
// pseudo-logit log(s/(1-s)) — the score itself has no resolution left near 1.0, its logit does.
template <int MODE>
__device__ __forceinline__ uint32_t frame_bin(unsigned long long c) {
  if (MODE == TAUV_TOPK_SIGMOID_PEAK) {
    const float s = key_to_float(composite_key(c));
    return float_to_key(logf(__fdiv_rn(s, __fsub_rn(1.0f, s)))) >> 20;
  }
  return composite_key(c) >> 20;
}

// Per-CTA state in shared memory.
struct __align__(16) TileCtx {
  unsigned long long thr;  // push filter (a pre-composite, see push_entry)
  float thr_f;             // its value as a float (-inf: none)
  int count;               // list entries
  int n_conv;              // list[0, n_conv) already hold final sort keys (SIGMOID_PEAK)
  int n_boot;              // list[0, n_boot) are already counted in the frame histogram (-1: order lost, count nothing more)
  uint32_t emit, maxbin;
  int nhot;                // cluster kernel: queue tail (entries ever queued by the streaming warps)
  int qhead;               // cluster kernel: queue head (entries consumed by the service warp)
  int done;                // cluster kernel: streaming warps that have finished the unit
  int flags;               // bit 1 = the list overflowed
  int base, wsum[kTileThreads / 32];
  uint32_t sel[4];
  __align__(16) uint32_t fs[4];  // cp.async snapshot of frame_state[0..3]; possibly stale, never waited on mid-scan
};

// A list entry is a 64-bit "pre-composite": order-preserving key of the VALUE in the high word, ~flat index in the
// low word.  In RAW mode that already is the final sort key.  In SIGMOID_PEAK mode the value is the logit; the
// sigmoid (and the final key) is applied later, densely, by convert_entries().  Entries below ctx->thr are provably
// outside the frame's top-k and never enter the list.  Returns 2 when the entry had to be dropped (list full).
__device__ __forceinline__ int push_entry(const TileArgs& a, TileCtx* ctx, unsigned long long* list, float x,
                                          uint32_t flat) {
  const unsigned long long pre = make_composite(float_to_key(x), flat);
  if (pre < ctx->thr) return 0;
  // one shared-memory atomic per warp instruction instead of one per lane (same-address atomics serialise: in the
  // bootstrap round, where a ninth of all cells push, they were most of the round's time)
  const unsigned active = __activemask();
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(active) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(&ctx->count, __popc(active));
  base = __shfl_sync(active, base, leader);
  const int slot = base + __popc(active & ((1u << lane) - 1u));
  if (slot >= a.cap) return 2;
  list[slot] = pre;
  return 0;
}

// Full test of one strip of 4 (VEC) / one element that passed the threshold scan.  `plane` points at the (frame,
// class) plane in global memory, off is the float offset inside it; neighbours come from L1/L2.
template <bool SMEM>
__device__ __forceinline__ float ld1(const float* p) { return SMEM ? *p : __ldg(p); }
template <bool SMEM>
__device__ __forceinline__ float4 ld4(const float* p) {
  return SMEM ? *reinterpret_cast<const float4*>(p) : __ldg(reinterpret_cast<const float4*>(p));
}

// (SMEM: `plane` is a shared-memory copy that starts at plane offset `origin`; off stays a plane offset.)
template <int MODE, bool VEC, bool SMEM = false>
__device__ __noinline__ int examine(const TileArgs& a, TileCtx* ctx, unsigned long long* list,
                                    const float* __restrict__ plane, uint32_t plane_flat0, int off, int origin = 0) {
  const int W = a.W, H = a.H;
  const float thr_f = ctx->thr_f;
  const int r = off / W;
  const int col = off - r * W;
  const float* p1 = plane + (off - origin);
  const uint32_t flat = plane_flat0 + (uint32_t)off;
  int fl = 0;
  if (VEC) {
    const float4 x = ld4<SMEM>(p1);
    if (MODE == TAUV_TOPK_SIGMOID_PEAK) {
      // columns col-1 .. col+4 of rows r-1, r, r+1 (-inf outside the plane)
      float cm[6];  // column-wise max over the three rows
      {
        const bool hl = col > 0, hr = col + 4 < W;
        float4 u = make_float4(TAUV_NEG_INF, TAUV_NEG_INF, TAUV_NEG_INF, TAUV_NEG_INF), d = u;
        float ul = TAUV_NEG_INF, ur = TAUV_NEG_INF, dl = TAUV_NEG_INF, dright = TAUV_NEG_INF;
        if (r > 0) {
          u = ld4<SMEM>(p1 - W);
          if (hl) ul = ld1<SMEM>(p1 - W - 1);
          if (hr) ur = ld1<SMEM>(p1 - W + 4);
        }
        if (r + 1 < H) {
          d = ld4<SMEM>(p1 + W);
          if (hl) dl = ld1<SMEM>(p1 + W - 1);
          if (hr) dright = ld1<SMEM>(p1 + W + 4);
        }
        const float ml = hl ? ld1<SMEM>(p1 - 1) : TAUV_NEG_INF;
        const float mr = hr ? ld1<SMEM>(p1 + 4) : TAUV_NEG_INF;
        cm[0] = fmaxf(fmaxf(ul, ml), dl);
        cm[1] = fmaxf(fmaxf(u.x, x.x), d.x);
        cm[2] = fmaxf(fmaxf(u.y, x.y), d.y);
        cm[3] = fmaxf(fmaxf(u.z, x.z), d.z);
        cm[4] = fmaxf(fmaxf(u.w, x.w), d.w);
        cm[5] = fmaxf(fmaxf(ur, mr), dright);
      }
      const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const float xv = xs[cc];
        if (xv >= thr_f) {
          const float m = fmaxf(fmaxf(cm[cc], cm[cc + 1]), cm[cc + 2]);
          bool peak = (xv >= m);
          if (!peak && (xv > 4.0f || m < -80.0f || (m - xv) < 1e-3f)) peak = sigmoid_tie(xv, m);
          if (peak) fl |= push_entry(a, ctx, list, xv, flat + cc);
        }
      }
    } else {
      if (x.x >= thr_f) fl |= push_entry(a, ctx, list, x.x, flat);
      if (x.y >= thr_f) fl |= push_entry(a, ctx, list, x.y, flat + 1);
      if (x.z >= thr_f) fl |= push_entry(a, ctx, list, x.z, flat + 2);
      if (x.w >= thr_f) fl |= push_entry(a, ctx, list, x.w, flat + 3);
    }
  } else {
    const float xv = ld1<SMEM>(p1);
    if (MODE == TAUV_TOPK_SIGMOID_PEAK) {
      float m = TAUV_NEG_INF;
      for (int dy = -1; dy <= 1; ++dy) {
        if (r + dy < 0 || r + dy >= H) continue;
        for (int dx = -1; dx <= 1; ++dx)
          if (col + dx >= 0 && col + dx < W) m = fmaxf(m, ld1<SMEM>(p1 + dy * W + dx));
      }
      bool peak = (xv >= m);
      if (!peak && (xv > 4.0f || m < -80.0f || (m - xv) < 1e-3f)) peak = sigmoid_tie(xv, m);
      if (peak) fl |= push_entry(a, ctx, list, xv, flat);
    } else {
      fl |= push_entry(a, ctx, list, xv, flat);
    }
  }
  return fl;
}

__device__ __forceinline__ void set_thr(TileCtx* ctx, unsigned long long t) {  // one thread
  if (t > ctx->thr) {
    ctx->thr = t;
    ctx->thr_f = composite_key(t) ? key_to_float(composite_key(t)) : TAUV_NEG_INF;
  }
}

// Threshold scan of the elements [e0, e1) of a plane (both multiples of 4 on the VEC path): every thread streams
// four 128-bit loads at a time straight from HBM (read-once data: no L1 allocation) and compares the strip maximum
// with the threshold; only what passes is examined.  With a published frame threshold that is ~1 % of the strips.
template <int MODE, bool VEC>
__device__ __forceinline__ int scan_elems(const TileArgs& a, TileCtx* ctx, unsigned long long* list,
                                          const float* __restrict__ plane, uint32_t plane_flat0, int e0, int e1,
                                          const uint32_t* fstate = nullptr) {
  float thr_f = ctx->thr_f;
  const int tid = threadIdx.x;
  int fl = 0;
  if (VEC) {
    const int t1 = e1 >> 2;
#pragma unroll 1
    for (int t0 = (e0 >> 2) + tid; t0 < t1; t0 += 4 * kTileThreads) {
      if (fstate) {
        // Pick up what the frame's other items have published since this item started: thread 0 folds in the last
        // snapshot and requests a fresh one (asynchronously; not in the last round, so nothing is pending at exit).
        if (tid == 0) {
          set_thr(ctx, (unsigned long long)ctx->fs[0] << 32);
          if (t0 + 4 * kTileThreads < t1)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(ctx->fs)), "l"(fstate) : "memory");
        }
        thr_f = *reinterpret_cast<volatile float*>(&ctx->thr_f);
      }
      float4 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = t0 + u * kTileThreads;
        x[u] = (t < t1) ? ldg_stream4(plane + ((size_t)t << 2))
                        : make_float4(TAUV_NEG_INF, TAUV_NEG_INF, TAUV_NEG_INF, TAUV_NEG_INF);
      }
      uint32_t hot = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (fmaxf(fmaxf(x[u].x, x[u].y), fmaxf(x[u].z, x[u].w)) >= thr_f && t0 + u * kTileThreads < t1) hot |= 1u << u;
#pragma unroll 1
      while (hot) {
        const int u = __ffs(hot) - 1;
        hot &= hot - 1;
        fl |= examine<MODE, VEC>(a, ctx, list, plane, plane_flat0, (t0 + u * kTileThreads) << 2);
      }
    }
  } else {
#pragma unroll 1
    for (int t = e0 + tid; t < e1; t += kTileThreads)
      if (__ldg(plane + t) >= thr_f) fl |= examine<MODE, VEC>(a, ctx, list, plane, plane_flat0, t);
  }
  return fl;
}

// logit pre-composites [n_conv, n) -> final sort keys (0 for a sigmoid that underflowed to 0: zero-valued cells are
// supplied by the merge kernel's filler, like non-peaks)
template <int MODE>
__device__ void convert_entries(TileCtx* ctx, unsigned long long* list, int n) {
  if (MODE == TAUV_TOPK_SIGMOID_PEAK) {
    for (int i = ctx->n_conv + (int)threadIdx.x; i < n; i += kTileThreads) {
      const unsigned long long pre = list[i];
      const float s = sigmoid_ref(key_to_float(composite_key(pre)));
      list[i] = (s > 0.0f) ? (((unsigned long long)float_to_key(s) << 32) | (pre & 0xffffffffull)) : 0ull;
    }
    __syncthreads();
  }
}

// stable in-place compaction of list[0,n) by a predicate; the new length lands in ctx->base
template <class Keep>
__device__ void compact_list(TileCtx* ctx, unsigned long long* list, int n, Keep keep_fn) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) ctx->base = 0;
  __syncthreads();
  for (int start = 0; start < n; start += kTileThreads) {
    const int i = start + tid;
    unsigned long long c = 0ull;
    bool keep = false;
    if (i < n) {
      c = list[i];
      keep = keep_fn(c);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) ctx->wsum[warp] = __popc(bal);
    __syncthreads();
    int pos = ctx->base + __popc(bal & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) pos += ctx->wsum[w];
    if (keep) list[pos] = c;  // write index <= read index, and every read of this round is already done
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < kTileThreads / 32; ++w) tot += ctx->wsum[w];
      ctx->base += tot;
    }
    __syncthreads();
  }
}

// exact prune of the list to its top-k, then tighten the item's own push filter from the k-th key
template <int MODE>
__device__ __noinline__ void prune_list(const TileArgs& a, TileCtx* ctx, unsigned long long* list, uint32_t* hist) {
  const int n = ctx->count;
  convert_entries<MODE>(ctx, list, n);
  const unsigned long long T = block_kth_largest<kTileThreads>([&](int i) { return list[i]; }, n, a.k, hist, ctx->sel);
  compact_list(ctx, list, n, [&](unsigned long long c) { return c >= T && c != 0ull; });
  if (threadIdx.x == 0) {
    ctx->count = ctx->base;
    ctx->n_conv = ctx->base;
    if (ctx->n_boot > 0) ctx->n_boot = -1;  // compaction moved entries: stop adding this item to the frame's bins
    if (ctx->base >= a.k) {
      if (MODE == TAUV_TOPK_SIGMOID_PEAK)
        set_thr(ctx, (unsigned long long)reject_key_for_score(key_to_float(composite_key(T))) << 32);
      else
        set_thr(ctx, T);
    }
  }
  __syncthreads();
}

// The list overflowed during the one-shot scan: start the item over in sub-steps that cannot overflow, pruning to
// the exact top-k (which also raises the item's own threshold) whenever the list passes `soft`.
template <int MODE, bool VEC>
__device__ __noinline__ void rescan_item_safely(const TileArgs& a, TileCtx* ctx, unsigned long long* list,
                                                uint32_t* hist, const float* __restrict__ plane,
                                                uint32_t plane_flat0, int e0, int e1) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ctx->count = 0;
    ctx->n_conv = 0;
    if (ctx->n_boot > 0) ctx->n_boot = -1;
  }
  __syncthreads();
  for (int s0 = e0; s0 < e1; s0 += a.sub_elems) {
    scan_elems<MODE, VEC>(a, ctx, list, plane, plane_flat0, s0, min(s0 + a.sub_elems, e1));
    __syncthreads();
    if (ctx->count > a.soft) prune_list<MODE>(a, ctx, list, hist);  // uniform: nobody pushes before the next barrier
    __syncthreads();
  }
}

// Warp 0: account `n_added` new candidates of this CTA (already added to the bins; ctx->maxbin = their highest bin)
// and, when worthwhile, rescan the frame's bins and publish the rejection key.  Returns the key the frame has
// published after this call (0: none yet), broadcast to the warp.
template <int MODE>
__device__ __forceinline__ uint32_t frame_republish(const TileArgs& a, TileCtx* ctx, uint32_t* fstate,
                                                    uint32_t n_added, bool force) {
  const int lane = threadIdx.x & 31;
  uint32_t before = 0, maxbin = 0;
  if (lane == 0) {
    maxbin = max(atomicMax(fstate + 1, ctx->maxbin), ctx->maxbin);
    before = atomicAdd(fstate + 2, n_added);
  }
  before = __shfl_sync(0xffffffffu, before, 0);
  maxbin = __shfl_sync(0xffffffffu, maxbin, 0);
  // The bins are rescanned only when the frame's candidate count crosses k, 2k, 4k, ...: each doubling tightens
  // the threshold noticeably, more often does not pay for the scan's round trips to L2.
  const uint32_t k = (uint32_t)a.k, after = before + n_added;
  const bool rescan = after >= k && (force || before < k || (31 - __clz(after / k)) != (31 - __clz(before / k)));
  uint32_t key = 0;
  if (rescan) {
    __threadfence();
    uint32_t acc = 0;
    int found = -1;
    for (int it = 0; it < 8 && found < 0; ++it) {  // at most 256 bins below the top occupied one
      const int bin = (int)maxbin - it * 32 - lane;
      uint32_t v = bin >= 0 ? *reinterpret_cast<volatile uint32_t*>(fstate + 4 + bin) : 0u;
      uint32_t pre = v;  // inclusive prefix over lanes (lane 0 = highest bin)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += t;
      }
      const unsigned hit = __ballot_sync(0xffffffffu, acc + pre >= k);
      if (hit) found = (int)maxbin - it * 32 - (__ffs(hit) - 1);
      acc += __shfl_sync(0xffffffffu, pre, 31);
      if ((int)maxbin - (it + 1) * 32 < 0) break;
    }
    if (lane == 0 && found > 0) {
      const float edge = key_to_float((uint32_t)found << 20);  // lowest value of the bin
      if (MODE == TAUV_TOPK_SIGMOID_PEAK) key = reject_key_for_score(sigmoid_ref(edge));
      else key = (uint32_t)found << 20;
      if (key) key = max(atomicMax(fstate, key), key);
    }
  }
  return __shfl_sync(0xffffffffu, key, 0);
}

// end of an item: emit its top-k into the candidate table, add them to the frame's histogram, and republish the
// frame's rejection threshold (the lower edge of the highest bin b with at least k candidates of the whole frame in
// bins >= b; counts only grow, so a published threshold stays valid)
template <int MODE>
__device__ __noinline__ void finish_item(const TileArgs& a, TileCtx* ctx, unsigned long long* list, uint32_t* hist,
                                         int item, uint32_t* fstate) {
  const int tid = threadIdx.x;
  const int n = ctx->count;
  convert_entries<MODE>(ctx, list, n);
  const unsigned long long T = block_kth_largest<kTileThreads>([&](int i) { return list[i]; }, n, a.k, hist, ctx->sel);
  if (tid == 0) {
    ctx->emit = 0;
    ctx->maxbin = 0;
    ctx->base = 0;  // candidates newly added to the frame's bins
  }
  __syncthreads();
  unsigned long long* out = a.cand + (size_t)item * a.k;
  uint32_t my_maxbin = 0;
  const int n_boot = ctx->n_boot;  // list[0, n_boot) went into the frame's bins during the bootstrap round already
  for (int i = tid; i < n; i += kTileThreads) {
    const unsigned long long c = list[i];
    if (c >= T && c != 0ull) {
      out[atomicAdd(&ctx->emit, 1u)] = c;
      if (n_boot >= 0 && i >= n_boot) {
        const uint32_t bin = frame_bin<MODE>(c);
        atomicAdd(fstate + 4 + bin, 1u);
        atomicAdd(&ctx->base, 1);
        my_maxbin = max(my_maxbin, bin);
      }
    }
  }
  if (my_maxbin) atomicMax(&ctx->maxbin, my_maxbin);
  __syncthreads();
  const int n_emit = (int)ctx->emit;
  if (tid == 0) a.cand_count[item] = n_emit;
  if (tid < 32 && ctx->base > 0) frame_republish<MODE>(a, ctx, fstate, (uint32_t)ctx->base, false);
}

// One CTA per item.  Block index -> item interleaves the frames (consecutive blocks are different frames), so the
// first items of EVERY frame finish early and publish a threshold for the frame's other items.
template <int MODE, bool VEC>
__global__ void __launch_bounds__(kTileThreads) tile_topk_kernel(const __grid_constant__ TileArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned long long* list = reinterpret_cast<unsigned long long*>(smem_raw);
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + (size_t)a.cap * 8);
  __shared__ TileCtx s_ctx;
  TileCtx* ctx = &s_ctx;
  const int tid = threadIdx.x;

  const int item_in_frame = blockIdx.x / a.B;
  const int frame = blockIdx.x - item_in_frame * a.B;
  const int item = frame * a.rows_per_frame + item_in_frame;  // row of the candidate table
  const int c_in_frame = item_in_frame / a.items_per_plane;
  const int ip = item_in_frame - c_in_frame * a.items_per_plane;
  const int r0 = ip * a.rows_per_item;
  const int r1 = min(a.H, r0 + a.rows_per_item);
  const float* plane = a.hm + ((size_t)frame * a.C + c_in_frame) * a.H * a.W;
  const uint32_t plane_flat0 = (uint32_t)c_in_frame * (uint32_t)(a.H * a.W);
  uint32_t* fstate = a.frame_state + (size_t)frame * kFrameStateWords;

  long long tr[4] = {0, 0, 0, 0};
  auto now = []() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; };
  if (a.trace && tid == 0) tr[0] = now();
  if (tid == 0) {
    const uint32_t key = *reinterpret_cast<volatile uint32_t*>(fstate);  // what the frame's finished items published
    if (a.trace) a.trace[(size_t)blockIdx.x * 8 + 5] = key;
    ctx->count = 0;
    ctx->n_conv = 0;
    ctx->n_boot = 0;
    ctx->flags = 0;
    ctx->fs[0] = ctx->fs[1] = ctx->fs[2] = ctx->fs[3] = 0;
    ctx->thr = (unsigned long long)key << 32;
    ctx->thr_f = key ? key_to_float(key) : TAUV_NEG_INF;
  }
  __syncthreads();

  const int e0 = r0 * a.W, e1 = r1 * a.W;
  if (a.trace && tid == 0) tr[1] = now();
  const int fl = scan_elems<MODE, VEC>(a, ctx, list, plane, plane_flat0, e0, e1, fstate);
  if (fl) atomicOr(&ctx->flags, fl);
  if (tid == 0) asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  if (ctx->flags & 2) rescan_item_safely<MODE, VEC>(a, ctx, list, hist, plane, plane_flat0, e0, e1);

  if (a.trace && tid == 0) {
    tr[2] = now();
    a.trace[(size_t)blockIdx.x * 8 + 4] = ctx->count;
  }
  if (ctx->count == 0) {
    if (tid == 0) a.cand_count[item] = 0;
  } else {
    finish_item<MODE>(a, ctx, list, hist, item, fstate);
  }
  if (a.trace && tid == 0) {
    tr[3] = now();
    for (int i = 0; i < 4; ++i) a.trace[(size_t)blockIdx.x * 8 + i] = tr[i];
  }
}

// ----------------------------------------------------------------------------------------------
// Frame-level selection and ranked output (shared by the merge kernel and the cluster kernel's fused tail)
// ----------------------------------------------------------------------------------------------
struct BoxVals {
  float h, w, depth;
  double y, x;
};
// the per-detection arithmetic of decode() / decode_keypoints() for the cell (iy, ix) of frame b
__device__ __forceinline__ BoxVals box_values(const BoxArgs& g, int b, int iy, int ix) {
  BoxVals v;
  v.h = g.size[b * g.ss[0] + iy * g.ss[1] + ix * g.ss[2]];
  v.w = g.size[b * g.ss[0] + iy * g.ss[1] + ix * g.ss[2] + g.ss[3]];
  if (g.mode == TAUV_BOX_DECODE) {
    // decode.py:214-215: Python doubles
    const float oy = g.offset[b * g.os[0] + iy * g.os[1] + ix * g.os[2]];
    const float ox = g.offset[b * g.os[0] + iy * g.os[1] + ix * g.os[2] + g.os[3]];
    v.y = __ddiv_rn(__dadd_rn(__dmul_rn((double)g.ratio, (double)iy), (double)oy), (double)g.in_h);
    v.x = __ddiv_rn(__dadd_rn(__dmul_rn((double)g.ratio, (double)ix), (double)ox), (double)g.in_w);
  } else {
    // decode.py:87-88: int64 tensor / int -> fp32 true divide, then float()
    v.y = (double)__fdiv_rn((float)iy, (float)g.out_h);
    v.x = (double)__fdiv_rn((float)ix, (float)g.out_w);
  }
  v.depth = 0.0f;
  if (g.depth != nullptr && g.depth_out != nullptr) {
    const float d = g.depth[b * g.ds[0] + iy * g.ds[1] + ix * g.ds[2]];
    float inv = __fdiv_rn(1.0f, sigmoid_ref(d));
    if (g.mode == TAUV_BOX_DECODE) inv = __fsub_rn(inv, 1.0f);  // decode.py:324
    v.depth = inv;
  }
  return v;
}
__device__ __forceinline__ void box_store(const BoxArgs& g, long long slot, const BoxVals& v) {
  g.hw[slot * 2 + 0] = v.h;
  g.hw[slot * 2 + 1] = v.w;
  g.yx[slot * 2 + 0] = v.y;
  g.yx[slot * 2 + 1] = v.x;
  if (g.depth != nullptr && g.depth_out != nullptr) g.depth_out[slot] = v.depth;
}
__device__ __forceinline__ void box_one(const BoxArgs& g, int b, long long slot, int iy, int ix) {
  box_store(g, slot, box_values(g, b, iy, ix));
}

// Select the k best of pool[0, total) (distinct composite keys, in shared memory) into sel[0, npos), npos = min(k, total),
// unordered.  sel must hold p2 = next_pow2(k) entries and is zero-padded.  All NT threads of the CTA call this.
template <int NT>
__device__ int topk_select_pool(const unsigned long long* pool, int total, int k, int p2, unsigned long long* sel,
                                uint32_t* hist, uint32_t* ctl) {
  const int tid = threadIdx.x;
  for (int i = tid; i < p2; i += NT) sel[i] = 0ull;
  if (tid == 0) ctl[5] = 0;
  __syncthreads();
  unsigned long long T = 1ull;  // total <= k: everything valid (non-zero) is selected
  if (total > k) T = block_kth_largest<NT>([&](int i) { return pool[i]; }, total, k, hist, ctl);
  for (int i = tid; i < total; i += NT) {
    const unsigned long long c = pool[i];
    if (c >= T && c != 0ull) sel[atomicAdd(&ctl[5], 1u)] = c;
  }
  __syncthreads();
  return (int)ctl[5];
}

// The zero-score tail of a ranked list (SIGMOID_PEAK, fewer than k positive peaks): cells of value 0 in ascending flat
// index, skipping the cells flagged as selected peaks (flags[i] != 0 for flat index i < k).  Slots [npos, k).
template <int NT>
__device__ void topk_emit_fillers(const uint32_t* flags, int npos, int b, int k, int H, int W,
                                  int64_t* __restrict__ index, int64_t* __restrict__ label, float* __restrict__ score,
                                  const BoxArgs& g) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int s_fsum[NT / 32];
  const long long hw_elems = (long long)H * W;
  const int need = k - npos;
  int base = 0;
  for (int start = 0; start < k && base < need; start += NT) {
    const int i = start + tid;
    const bool freec = (i < k) && (flags[i] == 0u);
    const unsigned bal = __ballot_sync(0xffffffffu, freec);
    if (lane == 0) s_fsum[warp] = __popc(bal);
    __syncthreads();
    int pos = base + __popc(bal & ((1u << lane) - 1u));
    int tot = 0;
    for (int w = 0; w < NT / 32; ++w) {
      if (w < warp) pos += s_fsum[w];
      tot += s_fsum[w];
    }
    if (freec && pos < need) {
      const int r = npos + pos;
      const long long lab = i / hw_elems;
      const long long rem = i - lab * hw_elems;
      const int iy = (int)(rem / W), ix = (int)(rem - (long long)iy * W);
      const long long slot = (long long)b * k + r;
      index[slot * 2 + 0] = iy;
      index[slot * 2 + 1] = ix;
      label[slot] = lab;
      score[slot] = 0.0f;
      if (g.enabled) box_one(g, b, slot, iy, ix);
    }
    base += tot;
    __syncthreads();
  }
}

// Sort sel[0, p2) descending and write the frame's ranked outputs: index/label/score for the npos selected peaks,
// then (SIGMOID_PEAK) the zero-valued fillers a dense stable top-k would return, the box arithmetic, and the count of
// leading entries at or above the score threshold.  flags: k words of shared memory (may alias the radix histogram).
template <int MODE, int NT>
__device__ void topk_emit_ranked(unsigned long long* sel, int p2, int npos, uint32_t* flags, int b, int k, int H, int W,
                                 int64_t* __restrict__ index, int64_t* __restrict__ label, float* __restrict__ score,
                                 const BoxArgs& g) {
  const int tid = threadIdx.x;
  __shared__ int s_first_below;
  if (tid == 0) s_first_below = k;
  block_bitonic_sort_desc<NT>(sel, p2);
  const long long hw_elems = (long long)H * W;
  for (int r = tid; r < npos; r += NT) {
    const unsigned long long c = sel[r];
    const uint32_t flat = composite_idx(c);
    const float s = key_to_float(composite_key(c));
    const long long lab = flat / hw_elems;
    const long long rem = flat - lab * hw_elems;
    const int iy = (int)(rem / W), ix = (int)(rem - (long long)iy * W);
    const long long slot = (long long)b * k + r;
    index[slot * 2 + 0] = iy;
    index[slot * 2 + 1] = ix;
    label[slot] = lab;
    score[slot] = s;
    if (g.enabled) {
      box_one(g, b, slot, iy, ix);
      if (s < g.thr) atomicMin(&s_first_below, r);
    }
  }
  if (MODE == TAUV_TOPK_SIGMOID_PEAK && npos < k) {
    // Dense stable top-k semantics: the remaining slots are zero-valued cells in ascending flat index.  At most
    // npos of the first k cells are positive peaks, so [0,k) always suffices.
    __syncthreads();
    for (int i = tid; i < k; i += NT) flags[i] = 0u;
    __syncthreads();
    for (int r = tid; r < npos; r += NT) {
      const uint32_t flat = composite_idx(sel[r]);
      if (flat < (uint32_t)k) flags[flat] = 1u;
    }
    __syncthreads();
    topk_emit_fillers<NT>(flags, npos, b, k, H, W, index, label, score, g);
    if (g.enabled && 0.0f < g.thr) {
      __syncthreads();
      if (tid == 0) atomicMin(&s_first_below, npos);
    }
  }
  if (g.enabled) {
    __syncthreads();
    if (tid == 0) g.count[b] = s_first_below;
  }
}

// ----------------------------------------------------------------------------------------------
// K1 (vectorised path): one thread-block cluster per "unit" (a frame, or a contiguous share of a frame's items)
// ----------------------------------------------------------------------------------------------
// The rejection threshold of a unit and the candidate histogram it is derived from live in the DISTRIBUTED SHARED
// MEMORY of the cluster that owns the unit — no global state, no memset, no separate seed launch:
//   * bins: 8192 counters over the top 13 bits of the order-preserving key, 1024 per CTA (bin b lives in CTA b>>10);
//     every CTA adds its candidates with remote shared-memory atomics;
//   * thr_key: every CTA holds its own copy of the published rejection key; a publisher raises all eight copies with
//     remote atomicMax, so the streaming loop reads the threshold from local shared memory;
//   * next_item (rank 0): the unit's dynamic work queue.
// Timeline of a unit: all eight CTAs load round 0 of their first item (4096 cells each) plus its halo rows, run the
// full 3x3 peak test on it from a shared-memory tile (the only place every cell is tested), bin the peaks, and after
// one cluster barrier each CTA derives the first threshold from the 8 x 4096-cell sample: at least k genuine peaks
// of the frame lie at or above it, so nothing below it can be in the frame's top-k.  From then on the CTAs stream:
// the loads of the next round — of this item or of the CTA's next item — are issued before the current round is
// compared against the threshold, only strips that pass get the peak test (neighbours from L1/L2), and at the end of
// every item its candidates go to the global candidate table and into the bins, from which the threshold is
// republished whenever the unit's candidate count crosses k, 2k, 4k, ...
constexpr int kClSize = 8;                            // CTAs per cluster (portable maximum)
constexpr int kClBinShift = 19;                       // fine bin = key >> 19: sign + 8 exponent + 4 mantissa bits
constexpr int kClWin = 1024;                          // bins per window (64 binades at 16 bins each)
constexpr int kClBins = 2 * kClWin;                   // negative window + positive window
constexpr int kClNegBase = 1536;                      // fine bins [1536, 2560): -2^33 .. -2^-31
constexpr int kClPosBase = 5632;                      // fine bins [5632, 6656): +2^-31 .. +2^33
// (measured on B200: 2 beats 3 and 4 — wider rounds spill a loaded register, and a spill right after the load waits for it)
#ifndef TAUV_STREAM_PF
#define TAUV_STREAM_PF 3
#endif
#ifndef TAUV_ROUND_W
#define TAUV_ROUND_W 2
#endif
constexpr int kRoundW = TAUV_ROUND_W;                            // 128-bit strips per thread and round (2 x kRoundW live in registers)
constexpr int kStreamThreads = kTileThreads - 32;     // warps 1..7 stream, warp 0 serves
constexpr int kRoundF4 = kRoundW * kStreamThreads;    // 128-bit strips per streaming round per CTA
#ifndef TAUV_SCAN_DIV
#define TAUV_SCAN_DIV 8
#endif
constexpr int kScanDiv = TAUV_SCAN_DIV;
#ifndef TAUV_SERVE_W
#define TAUV_SERVE_W 1
#endif
constexpr int kServeW = TAUV_SERVE_W;                  // queue entries per lane and service step
#ifndef TAUV_BOOT_W
#define TAUV_BOOT_W 2
#endif
constexpr int kBootW = TAUV_BOOT_W;                    // 128-bit strips per thread in a bootstrap chunk
constexpr int kBootF4 = kBootW * kTileThreads;        // strips of a bootstrap chunk (all eight warps)
constexpr int kBootElems = 4 * kBootF4;               // cells of the bootstrap round
constexpr int kFuseMaxK = 256;                        // fused tail: flags / ranks for k output slots
constexpr int kHotCap = kBootW >= 2 ? 1024 : 512;       // ring of queued peak tests (entries of 8 bytes; it lives in the bootstrap tile)
constexpr int kClMaxW = 1016;                         // halo rows are held in two 128-bit registers per thread

struct __align__(16) ClusterCtx {
  uint32_t bins[kClBins];  // THIS CTA's candidates by window bin; a scan sums the eight CTAs' copies with remote loads
  uint32_t thr_key;        // this CTA's copy of the unit's published rejection key
  uint32_t maxbin;         // this CTA's copy of the highest occupied window bin of the unit
  int next_item;           // rank 0 only: next unclaimed item of the unit (index inside the frame)
};

// Window bin of an order-preserving key.  The map is monotone and only ever moves a value DOWN (values between or
// above the windows go to the top bin of the window below them), so "at least k candidates in bins >= b" still
// certifies that k candidates are >= the lower edge of b.  Values below the negative window are not counted (-1).
__device__ __forceinline__ int cl_window_bin(uint32_t key) {
  const int b = (int)(key >> kClBinShift);
  if (b >= kClPosBase) return kClWin + min(b - kClPosBase, kClWin - 1);
  if (b >= kClNegBase) return min(b - kClNegBase, kClWin - 1);
  return -1;
}
__device__ __forceinline__ float cl_window_edge(int wbin) {  // lowest value that maps to the bin
  const int b = wbin >= kClWin ? wbin - kClWin + kClPosBase : wbin + kClNegBase;
  return key_to_float((uint32_t)b << kClBinShift);
}

// Shared-memory layout of the cluster kernel (all dynamic, so that device functions reach it without pointer
// arguments): [ClusterCtx | TileCtx | list[cap] | radix histogram | bootstrap tile / queue of deferred peak tests]
constexpr int kClOffCtx = (int)((sizeof(ClusterCtx) + 15) / 16 * 16);
constexpr int kClOffList = kClOffCtx + (int)((sizeof(TileCtx) + 15) / 16 * 16);
__device__ __forceinline__ unsigned char* cl_smem() {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  return smem_raw;
}
__device__ __forceinline__ ClusterCtx* cl_cc() { return reinterpret_cast<ClusterCtx*>(cl_smem()); }
__device__ __forceinline__ TileCtx* cl_ctx() { return reinterpret_cast<TileCtx*>(cl_smem() + kClOffCtx); }
__device__ __forceinline__ unsigned long long* cl_list() { return reinterpret_cast<unsigned long long*>(cl_smem() + kClOffList); }
__device__ __forceinline__ uint32_t* cl_hist(const TileArgs& a) {
  return reinterpret_cast<uint32_t*>(cl_smem() + kClOffList + (size_t)a.cap * 8);
}
__device__ __forceinline__ float* cl_tile(const TileArgs& a) {
  return reinterpret_cast<float*>(cl_smem() + kClOffList + (size_t)a.cap * 8 + kRadixBins * 4);
}

struct ItemGeom {
  const float* plane;
  uint32_t plane_flat0;
  int e0, e1;  // element range inside the plane
  int item;    // global item number (row of the candidate table)
};

__device__ __forceinline__ ItemGeom item_geom(const TileArgs& a, int frame, int iif) {
  ItemGeom g;
  g.item = frame * (a.C * a.items_per_plane) + iif;
  const int c = iif / a.items_per_plane;
  const int ip = iif - c * a.items_per_plane;
  const int r0 = ip * a.rows_per_item;
  const int r1 = min(a.H, r0 + a.rows_per_item);
  g.plane = a.hm + ((size_t)frame * a.C + c) * a.H * a.W;
  g.plane_flat0 = (uint32_t)c * (uint32_t)(a.H * a.W);
  g.e0 = r0 * a.W;
  g.e1 = r1 * a.W;
  return g;
}

// kRoundW 128-bit strips per thread: strips s0 + u*NT + t of the item (s0 relative to the item's first strip)
template <int NT, int NW>
__device__ __forceinline__ void load_strips(const ItemGeom& g, int s0, int t, float4 (&x)[NW]) {
  const int t1 = g.e1 >> 2;
  const int tb = (g.e0 >> 2) + s0 + t;
#pragma unroll
  for (int u = 0; u < NW; ++u) {
    const int tt = tb + u * NT;
    x[u] = (tt < t1) ? ldg_stream4(g.plane + ((size_t)tt << 2))
                     : make_float4(TAUV_NEG_INF, TAUV_NEG_INF, TAUV_NEG_INF, TAUV_NEG_INF);
  }
}

template <int MODE>
__device__ __forceinline__ int cl_bin(unsigned long long c) {  // window bin of a FINAL sort key
  if (MODE == TAUV_TOPK_SIGMOID_PEAK) {
    const float s = key_to_float(composite_key(c));
    return cl_window_bin(float_to_key(logf(__fdiv_rn(s, __fsub_rn(1.0f, s)))));
  }
  return cl_window_bin(composite_key(c));
}

// Warp 0: scan the unit's histogram (the sum of the eight CTAs' local copies, read through distributed shared memory)
// from the highest occupied bin downwards, find the highest bin b with count(bins >= b) >= k, and return the
// rejection key of its lower edge (0: fewer than k candidates so far).  At most 512 bins are visited.
// (returns the bin, or -1 when fewer than k candidates have been binned; every lane gets the same value)
__device__ __forceinline__ int cl_scan_bin(cg::cluster_group& cluster, ClusterCtx* cc, int k) {
  const int lane = threadIdx.x & 31;
  const int maxbin = (int)*reinterpret_cast<volatile uint32_t*>(&cc->maxbin);
  uint32_t acc = 0;
  int found = -1;
  // (32 bins per step.  Four steps' worth of remote loads issued together were measured slower — 6.3 us instead of
  // 2.7 for the final prune, and the streaming end slipped by 2 us: the k-th best usually lies within the first 32-64
  // bins, and every extra remote load costs at the target CTA.)
  for (int it = 0; it < 16 && found < 0; ++it) {
    const int bin = maxbin - it * 32 - lane;
    uint32_t v = 0;
    if (bin >= 0) {
#pragma unroll
      for (int r = 0; r < kClSize; ++r) v += *reinterpret_cast<volatile uint32_t*>(&cluster.map_shared_rank(cc, r)->bins[bin]);
    }
    uint32_t pre = v;  // inclusive prefix over lanes (lane 0 = highest bin)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, acc + pre >= (uint32_t)k);
    if (hit) found = maxbin - it * 32 - (__ffs(hit) - 1);
    acc += __shfl_sync(0xffffffffu, pre, 31);
    if (maxbin - (it + 1) * 32 < 0) break;
  }
  return found;
}

template <int MODE>
__device__ __forceinline__ uint32_t cl_scan_threshold(cg::cluster_group& cluster, ClusterCtx* cc, int k) {
  const int lane = threadIdx.x & 31;
  const int found = cl_scan_bin(cluster, cc, k);
  uint32_t key = 0;
  if (lane == 0 && found >= 0) {
    const float edge = cl_window_edge(found);
    key = MODE == TAUV_TOPK_SIGMOID_PEAK ? reject_key_for_score(sigmoid_ref(edge)) : float_to_key(edge);
  }
  return __shfl_sync(0xffffffffu, key, 0);
}

// lanes 0..7 of one warp: raise every CTA's copy of a cluster-wide maximum
__device__ __forceinline__ void cl_raise_all(cg::cluster_group& cluster, uint32_t* local_word, uint32_t v) {
  const int lane = threadIdx.x & 31;
  if (v && lane < kClSize) atomicMax(cluster.map_shared_rank(local_word, lane), v);
}

// ---- streaming warps (1..7) -------------------------------------------------------------------------------------
// All items of this CTA in the unit (iif, iif + 8, ... < i_hi), the first one from strip s_begin on.  No barrier and no
// call in here, and nothing but the streaming state is live, so the next round's loads stay in registers: they are
// requested before the current round is compared against the threshold (of the same item, or the first round of the
// next item).  Strips that pass are only QUEUED for the service warp: their peak tests wait on neighbour loads, and
// done in line they would stall the streaming once per strip.
__device__ __noinline__ void cl_stream_all(const TileArgs& a, int frame, int iif, int i_hi, int s_begin) {
  const int st = (int)threadIdx.x - 32;  // thread index among the streaming threads
  TileCtx* const ctx = cl_ctx();
  int2* const hotq = reinterpret_cast<int2*>(cl_tile(a));
  ItemGeom g = item_geom(a, frame, iif);
  while (s_begin >= (g.e1 >> 2) - (g.e0 >> 2)) {  // the bootstrap round covered the whole first item
    iif += kClSize;
    s_begin = 0;
    if (iif >= i_hi) {
      iif = -1;
      break;
    }
    g = item_geom(a, frame, iif);
  }
  float4 xn[kRoundW];
  if (iif >= 0) load_strips<kStreamThreads, kRoundW>(g, s_begin, st, xn);
#pragma unroll 1
  while (iif >= 0) {
    const bool more = iif + kClSize < i_hi;
    const int t1 = g.e1 >> 2;
    const int n_strips = t1 - (g.e0 >> 2);
    if (a.trace && st == 0) {
      long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      a.trace[(size_t)g.item * 8 + 0] = t;
      a.trace[(size_t)g.item * 8 + 5] = composite_key(ctx->thr);
    }
#pragma unroll 1
    for (int s0 = s_begin; s0 < n_strips; s0 += kRoundF4) {
      float4 x[kRoundW];
#pragma unroll
      for (int u = 0; u < kRoundW; ++u) x[u] = xn[u];
      if (s0 + kRoundF4 < n_strips) load_strips<kStreamThreads, kRoundW>(g, s0 + kRoundF4, st, xn);
      else if (more) load_strips<kStreamThreads, kRoundW>(item_geom(a, frame, iif + kClSize), 0, st, xn);
#if TAUV_STREAM_PF > 0
      // L2 prefetch of the round TAUV_STREAM_PF ahead (inside the item): the registers hold two
      // rounds, ~28 KB per SM in flight, which covers ~0.8 us at 5 TB/s — HBM latency under this load is longer, an L2
      // hit is not.  One lane in eight covers its warp's 128-byte line per strip row.  Measured: 109.6 us without,
      // 104.7 / 104.5 / 106.0 us with 2 / 3 / 5 rounds ahead; carrying on into the CTA's next item: 105.5 (not kept).
      if ((st & 7) == 0) {
#pragma unroll
        for (int u = 0; u < kRoundW; ++u) {
          const int rel = s0 + TAUV_STREAM_PF * kRoundF4 + st + u * kStreamThreads;  // strip, relative to the item
          if (rel < n_strips) asm volatile("prefetch.global.L2 [%0];" ::"l"(g.plane + ((size_t)((g.e0 >> 2) + rel) << 2)));
        }
      }
#endif
      const float thr_f = *reinterpret_cast<volatile float*>(&ctx->thr_f);  // kept current by the service warp
      const int tb = (g.e0 >> 2) + s0 + st;
      uint32_t hot = 0;
#pragma unroll
      for (int u = 0; u < kRoundW; ++u)
        if (fmaxf(fmaxf(x[u].x, x[u].y), fmaxf(x[u].z, x[u].w)) >= thr_f && tb + u * kStreamThreads < t1) hot |= 1u << u;
#pragma unroll 1
      while (hot) {
        const int u = __ffs(hot) - 1;
        hot &= hot - 1;
        const int slot = atomicAdd(&ctx->nhot, 1);
        // ring full: wait for the service warp (it never waits for us, so this always ends)
        while (slot - *reinterpret_cast<volatile int*>(&ctx->qhead) >= kHotCap) __nanosleep(64);
        hotq[slot & (kHotCap - 1)] = make_int2(iif, (tb + u * kStreamThreads) << 2);
      }
    }
    if (a.trace && st == 0) {
      long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      a.trace[(size_t)g.item * 8 + 2] = t;
    }
    if (!more) break;
    iif += kClSize;
    s_begin = 0;
    g = item_geom(a, frame, iif);
  }
  __syncwarp();
  __threadfence_block();
  if ((threadIdx.x & 31) == 0) atomicAdd(&ctx->done, 1);
}

// ---- service warp (0) ----------------------------------------------------------------------------------------------
// Consumes the queue while the other warps stream: runs the peak tests (neighbours from L2/HBM; the latency is this
// warp's alone), appends candidates to the CTA's list, adds them to this CTA's bins, and whenever it has binned k/8
// new candidates rescans the unit's histogram (remote loads) and raises the rejection key in all eight CTAs.  It also
// folds keys published by the other CTAs into this CTA's threshold, which the streaming warps read every round.
template <int MODE>
__device__ __noinline__ void cl_service_warp(const TileArgs& a, int frame, cg::cluster_group& cluster) {
  const int lane = threadIdx.x & 31;
  TileCtx* const ctx = cl_ctx();
  ClusterCtx* const cc = cl_cc();
  unsigned long long* const list = cl_list();
  int2* const hotq = reinterpret_cast<int2*>(cl_tile(a));
  const int every = a.k >= kScanDiv ? a.k / kScanDiv : 1;  // new candidates between two rescans of the unit's histogram
  int head = 0, since_scan = 0;
  int n_binned = ctx->n_boot;  // list[0, n_binned) are in the bins already (bootstrap survivors); -1: stop binning
  while (true) {
    if (lane == 0) set_thr(ctx, (unsigned long long)(*reinterpret_cast<volatile uint32_t*>(&cc->thr_key)) << 32);
    __syncwarp();
    const int done = *reinterpret_cast<volatile int*>(&ctx->done);
    const int avail = *reinterpret_cast<volatile int*>(&ctx->nhot) - head;
    if (avail <= 0) {
      if (done == kStreamThreads / 32) break;
      __nanosleep(128);
      continue;
    }
    // up to kServeW entries per lane and step (measured: two per lane — half as many fixed costs per entry — is slower,
    // 104.9 -> 107.0 us: new candidates reach the bins, and with them the threshold, later)
    const int n = avail < 32 * kServeW ? avail : 32 * kServeW;
    int2 q[kServeW];
#pragma unroll
    for (int w = 0; w < kServeW; ++w) {
      q[w] = make_int2(-1, 0);
      if (lane + 32 * w < n) {
        volatile int2* slot = reinterpret_cast<volatile int2*>(&hotq[(head + lane + 32 * w) & (kHotCap - 1)]);
        while ((q[w].x = slot->x) < 0) {}  // (the producer is between its atomicAdd and its store)
        q[w].y = slot->y;
        slot->x = -1;                      // free the slot before the head moves past it
      }
    }
    __syncwarp();
    head += n;
    if (lane == 0) *reinterpret_cast<volatile int*>(&ctx->qhead) = head;
    int fl = 0;
#pragma unroll
    for (int w = 0; w < kServeW; ++w) {
      if (lane + 32 * w < n) {
        const ItemGeom g = item_geom(a, frame, q[w].x);
        fl |= examine<MODE, true>(a, ctx, list, g.plane, g.plane_flat0, q[w].y);
      }
    }
    fl = __reduce_or_sync(0xffffffffu, (unsigned)fl);
    if (fl & 2) {  // the list is full: everything is redone safely at the end of the unit; keep draining the queue
      if (lane == 0) ctx->flags |= 2;
      n_binned = -1;
      continue;
    }
    __syncwarp();
    const int cnt = *reinterpret_cast<volatile int*>(&ctx->count);
    if (n_binned >= 0 && cnt > n_binned) {
      // list[n_binned, cnt) are new and still carry logit keys (SIGMOID_PEAK), the space the bins live in
      uint32_t my_maxbin = 0;
      for (int i = n_binned + lane; i < cnt; i += 32) {
        const uint32_t key = composite_key(list[i]);
        if (MODE == TAUV_TOPK_SIGMOID_PEAK && !(key_to_float(key) > -80.0f)) continue;  // may underflow to score 0
        const int bin = cl_window_bin(key);
        if (bin < 0) continue;
        atomicAdd(&cc->bins[bin], 1u);
        my_maxbin = max(my_maxbin, (uint32_t)bin);
      }
      my_maxbin = __reduce_max_sync(0xffffffffu, my_maxbin);
      since_scan += cnt - n_binned;
      n_binned = cnt;
      if (my_maxbin > *reinterpret_cast<volatile uint32_t*>(&cc->maxbin)) cl_raise_all(cluster, &cc->maxbin, my_maxbin);
      if (since_scan >= every) {
        since_scan = 0;
        __threadfence_block();
        cl_raise_all(cluster, &cc->thr_key, cl_scan_threshold<MODE>(cluster, cc, a.k));
      }
    }
  }
}

// The list overflowed somewhere in the unit (plateaus, or no usable threshold): start this CTA's share of the unit
// over, in sub-steps that cannot overflow, pruning to the exact top-k whenever the list passes `soft`.
template <int MODE>
__device__ __noinline__ void cl_redo_all_safely(const TileArgs& a, TileCtx* ctx, unsigned long long* list, uint32_t* hist,
                                                int frame, int iif0, int i_hi) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ctx->count = 0;
    ctx->n_conv = 0;
    ctx->n_boot = -1;
    ctx->flags = 0;
  }
  __syncthreads();
  for (int iif = iif0; iif < i_hi; iif += kClSize) {
    const ItemGeom g = item_geom(a, frame, iif);
    for (int s0 = g.e0; s0 < g.e1; s0 += a.sub_elems) {
      scan_elems<MODE, true>(a, ctx, list, g.plane, g.plane_flat0, s0, min(s0 + a.sub_elems, g.e1));
      __syncthreads();
      if (ctx->count > a.soft) prune_list<MODE>(a, ctx, list, hist);  // uniform: nobody pushes before the next barrier
      __syncthreads();
    }
  }
}

// Bootstrap round of a unit: the full test on round 0 of this CTA's first item, from the shared-memory tile
// (tile[0] = plane cell `origin`; cells outside the plane are never read).  Peaks go to the list and into the bins.
template <int MODE>
__device__ __noinline__ void cl_bootstrap_round(const TileArgs& a, TileCtx* ctx, unsigned long long* list,
                                                uint32_t* hist, const float* tile, int origin, const ItemGeom& g,
                                                cg::cluster_group& cluster, ClusterCtx* cc) {
  const int tid = threadIdx.x;
  const int c_end = min(g.e1, g.e0 + kBootElems);
  int fl = 0;
  if (MODE == TAUV_TOPK_SIGMOID_PEAK) {
    // Every cell of the round gets the 3x3 test from the tile.  A ninth of them are peaks: they are kept in registers
    // and appended to the list at offsets from one block-wide scan — no atomics (pushed one by one through a
    // shared-memory counter they serialise and cost several microseconds here).
    const int W = a.W, H = a.H;
    const float NI = TAUV_NEG_INF;
    float pv[kBootW * 4];
    uint32_t pmask = 0;
#pragma unroll
    for (int u = 0; u < kBootW; ++u) {
      const int off = g.e0 + ((u * kTileThreads + tid) << 2);
      if (off < c_end) {
        const int r = off / W, col = off - r * W;
        const float* p1 = tile + (off - origin);
        const float4 x = *reinterpret_cast<const float4*>(p1);
        const bool hl = col > 0, hr = col + 4 < W;
        float4 up = make_float4(NI, NI, NI, NI), dn = up;
        float ul = NI, ur = NI, dl = NI, dr = NI;
        if (r > 0) {
          up = *reinterpret_cast<const float4*>(p1 - W);
          if (hl) ul = p1[-W - 1];
          if (hr) ur = p1[-W + 4];
        }
        if (r + 1 < H) {
          dn = *reinterpret_cast<const float4*>(p1 + W);
          if (hl) dl = p1[W - 1];
          if (hr) dr = p1[W + 4];
        }
        const float ml = hl ? p1[-1] : NI, mr = hr ? p1[4] : NI;
        float cm[6];
        cm[0] = fmaxf(fmaxf(ul, ml), dl);
        cm[1] = fmaxf(fmaxf(up.x, x.x), dn.x);
        cm[2] = fmaxf(fmaxf(up.y, x.y), dn.y);
        cm[3] = fmaxf(fmaxf(up.z, x.z), dn.z);
        cm[4] = fmaxf(fmaxf(up.w, x.w), dn.w);
        cm[5] = fmaxf(fmaxf(ur, mr), dr);
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const float xv = xs[cc];
          const float m = fmaxf(fmaxf(cm[cc], cm[cc + 1]), cm[cc + 2]);
          bool peak = (xv >= m);
          if (!peak && (xv > 4.0f || m < -80.0f || (m - xv) < 1e-3f)) peak = sigmoid_tie(xv, m);
          pv[u * 4 + cc] = xv;
          if (peak) pmask |= 1u << (u * 4 + cc);
        }
      }
    }
    // block-wide exclusive scan of the per-thread peak counts
    const int lane = tid & 31, warp = tid >> 5;
    const int mine = __popc(pmask);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) ctx->wsum[warp] = incl;
    __syncthreads();
    const int old_count = ctx->count;  // (entries of earlier bootstrap chunks)
    int base = old_count + incl - mine, total = old_count;
    for (int w = 0; w < kTileThreads / 32; ++w) {
      if (w < warp) base += ctx->wsum[w];
      total += ctx->wsum[w];
    }
    __syncthreads();  // everybody has read the old count
    if (total > a.cap) {
      fl = 2;  // (plateaus) the whole unit is redone safely
    } else {
#pragma unroll
      for (int i = 0; i < kBootW * 4; ++i) {
        if (pmask & (1u << i)) {
          const int off = g.e0 + (((i >> 2) * kTileThreads + tid) << 2) + (i & 3);
          list[base++] = make_composite(float_to_key(pv[i]), g.plane_flat0 + (uint32_t)off);
        }
      }
    }
    if (tid == 0) ctx->count = total > a.cap ? 0 : total;
  } else {
    // RAW: every cell is a candidate; keep this round's k best (exact selection straight from the tile)
    const int n = c_end - g.e0;
    auto load = [&](int i) { return make_composite(float_to_key(tile[g.e0 - origin + i]), g.plane_flat0 + (uint32_t)(g.e0 + i)); };
    const unsigned long long T = block_kth_largest<kTileThreads>(load, n, a.k, hist, ctx->sel);
    for (int i = tid; i < n; i += kTileThreads)
      if (load(i) >= T) fl |= push_entry(a, ctx, list, tile[g.e0 - origin + i], g.plane_flat0 + (uint32_t)(g.e0 + i));
  }
  if (fl) atomicOr(&ctx->flags, fl);
  if (tid == 0) ctx->maxbin = 0;
  if (a.trace && tid == 0) { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); a.trace[(size_t)g.item * 8 + 3] = t; }
  __syncthreads();
  if (a.trace && tid == 0) { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); a.trace[(size_t)g.item * 8 + 4] = t; }
  if (ctx->flags & 2) return;  // the list overflowed (plateaus): the item is redone safely later; nothing is binned now
  const int nb = ctx->count;
  uint32_t my_maxbin = 0;
  for (int i = max(ctx->n_boot, 0) + tid; i < nb; i += kTileThreads) {  // (what this chunk added)
    // (SIGMOID_PEAK: the entries still carry logit keys, which is the space the bins live in; a logit below -80 may
    // underflow to a zero score, which is no candidate)
    const uint32_t key = composite_key(list[i]);
    if (MODE == TAUV_TOPK_SIGMOID_PEAK && !(key_to_float(key) > -80.0f)) continue;
    const int bin = cl_window_bin(key);
    if (bin < 0) continue;
    atomicAdd(&cc->bins[bin], 1u);
    my_maxbin = max(my_maxbin, (uint32_t)bin);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_maxbin = max(my_maxbin, __shfl_xor_sync(0xffffffffu, my_maxbin, o));
  if ((tid & 31) == 0 && my_maxbin) atomicMax(&ctx->maxbin, my_maxbin);
  __syncthreads();
  if (tid < 32) cl_raise_all(cluster, &cc->maxbin, ctx->maxbin);
  if (tid == 0) ctx->n_boot = nb;
}

template <int MODE>
__global__ void __launch_bounds__(kTileThreads, 4) tile_cluster_kernel(const __grid_constant__ TileArgs a, int n_units,
                                                                       int parts) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  unsigned long long* list = cl_list();
  uint32_t* hist = cl_hist(a);
  float* tile = cl_tile(a);                    // [kBootElems + 2W + 8]
  int2* hotq = reinterpret_cast<int2*>(tile);  // the same memory after the bootstrap round: ring of queued peak tests
  TileCtx* ctx = cl_ctx();
  ClusterCtx* cc = cl_cc();
  const int tid = threadIdx.x;
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kClSize, ncl = gridDim.x / kClSize;
  const int ipf = a.C * a.items_per_plane;
  const int W = a.W;
  auto now = []() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; };
  const long long t_kernel = a.trace ? now() : 0;
  // one chunk (kBootElems cells per CTA) for k <= 128, then as many cells as 2048 per 128 of k, at most 16384
  const int n_boot_chunks = a.k <= 128 ? 1 : min(8, (a.k + 127) / 128) * (2048 / kBootElems);

#pragma unroll 1
  for (int unit = cid; unit < n_units; unit += ncl) {
    const int frame = unit / parts, part = unit - frame * parts;
    const int i_lo = (int)((long long)ipf * part / parts), i_hi = (int)((long long)ipf * (part + 1) / parts);
    // items of a unit are dealt round-robin to the cluster's CTAs: this CTA takes i_lo + rank, + 8, + 16, ...
    int iif = i_lo + rank;
    bool have = iif < i_hi;
    ItemGeom g = item_geom(a, frame, have ? iif : i_lo);
    float4 xn[kBootW];
    // halo of round 0: plane cells [e0 - W, e0) and [c_end, c_end + W + 4), two 128-bit strips per thread at most
    const int c_end = min(g.e1, g.e0 + kBootElems);
    const int origin = g.e0 - W;                 // plane cell held in tile[0] (may be negative: never read then)
    const int plane_cells = a.H * W;
    const int nh = W >> 2;                        // strips per halo row
    float4 halo[2];
    if (have) {
      load_strips<kTileThreads, kBootW>(g, 0, tid, xn);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int h = tid + j * kTileThreads;     // [0, nh): row above; [nh, 2nh+1): row below (+1 strip)
        int cell = -1;
        if (h < nh) cell = g.e0 - W + (h << 2);
        else if (h < 2 * nh + 1) cell = c_end + ((h - nh) << 2);
        halo[j] = (cell >= 0 && cell < plane_cells) ? ldg_stream4(g.plane + cell) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // reset the unit's state while the loads are in flight
    for (int i = tid; i < kClBins / 4; i += kTileThreads) reinterpret_cast<uint4*>(cc->bins)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
      cc->thr_key = 0u;
      cc->maxbin = 0u;
      ctx->count = 0;
      ctx->n_conv = 0;
      ctx->n_boot = 0;
      ctx->nhot = 0;
      ctx->qhead = 0;
      ctx->done = 0;
      ctx->flags = 0;
      ctx->thr = 0ull;
      ctx->thr_f = TAUV_NEG_INF;
    }
    // (1) every CTA's bins and words are ready (and nobody is still reading the previous unit's).  Split barrier: the
    // wait sits after the tile has been filled, so its latency hides behind the sample loads.
    cluster.barrier_arrive();

    long long tr0 = 0;
    if (a.trace && tid == 0) tr0 = now();
    if (have) {
      // ---- bootstrap: round 0 of the first item, every cell tested, from shared memory
#pragma unroll
      for (int u = 0; u < kBootW; ++u) {
        const int off = g.e0 + ((u * kTileThreads + tid) << 2);
        if (off < c_end) *reinterpret_cast<float4*>(tile + (off - origin)) = xn[u];
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int h = tid + j * kTileThreads;
        int cell = -1;
        if (h < nh) cell = g.e0 - W + (h << 2);
        else if (h < 2 * nh + 1) cell = c_end + ((h - nh) << 2);
        if (cell >= 0 && cell < plane_cells) *reinterpret_cast<float4*>(tile + (cell - origin)) = halo[j];
      }
      __syncthreads();
      if (a.trace && tid == 0) a.trace[(size_t)g.item * 8 + 1] = now();
    }
    cluster.barrier_wait();  // (1)
    if (have) {
      cl_bootstrap_round<MODE>(a, ctx, list, hist, tile, origin, g, cluster, cc);
      // A large k needs a larger sample: the first threshold lets through about k / (sample fraction) cells, and with
      // 2048 cells per CTA a k of 1000 would make every strip pass.  One more 2048-cell chunk per 128 of k.
      for (int c = 1; c < n_boot_chunks; ++c) {
        ItemGeom gb = g;
        gb.e0 = g.e0 + c * kBootElems;
        if (gb.e0 >= g.e1) break;
        const int ce = min(gb.e1, gb.e0 + kBootElems), org = gb.e0 - W;
        __syncthreads();  // the tile is re-used
        float4 t4[kBootW];
        load_strips<kTileThreads, kBootW>(gb, 0, tid, t4);
#pragma unroll
        for (int u = 0; u < kBootW; ++u) {
          const int off = gb.e0 + ((u * kTileThreads + tid) << 2);
          if (off < ce) *reinterpret_cast<float4*>(tile + (off - org)) = t4[u];
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int h = tid + j * kTileThreads;
          int cell = -1;
          if (h < nh) cell = gb.e0 - W + (h << 2);
          else if (h < 2 * nh + 1) cell = ce + ((h - nh) << 2);
          if (cell >= 0 && cell < plane_cells) *reinterpret_cast<float4*>(tile + (cell - org)) = ldg_stream4(g.plane + cell);
        }
        __syncthreads();
        cl_bootstrap_round<MODE>(a, ctx, list, hist, tile, org, gb, cluster, cc);
      }
      if (a.trace && tid == 0) a.trace[(size_t)g.item * 8 + 6] = now();
    }
    cluster.sync();  // (2) the sample of all eight CTAs is in their bins, the highest occupied bin is known everywhere
    // first threshold: every CTA derives it for itself (warp 0, remote loads only; nothing to publish)
    if (tid < 32) {
      const uint32_t key = cl_scan_threshold<MODE>(cluster, cc, a.k);
      if (tid == 0 && key) {
        atomicMax(&cc->thr_key, key);
        set_thr(ctx, (unsigned long long)(*reinterpret_cast<volatile uint32_t*>(&cc->thr_key)) << 32);
      }
    }
    __syncthreads();
    // (the bootstrap collected every peak of its sample; those below the first threshold stay in the list — they are
    // binned already and the tail's final prune drops them — rather than paying three barriers for a compaction here)
    if (a.trace && tid == 0 && have) a.trace[(size_t)g.item * 8 + 7] = now();

    // ---- stream: warps 1..7 stream every item of this CTA, warp 0 serves the queue; no barrier until both are done
    if (have) {
      for (int i = tid; i < kHotCap; i += kTileThreads) hotq[i] = make_int2(-1, 0);  // (the tile is free now)
      __syncthreads();
      if (tid < 32) cl_service_warp<MODE>(a, frame, cluster);
      else cl_stream_all(a, frame, iif, i_hi, n_boot_chunks * kBootF4);
      __syncthreads();
      if (ctx->flags & 2) cl_redo_all_safely<MODE>(a, ctx, list, hist, frame, iif, i_hi);
      if (a.trace && tid == 0) {
        const long long t = now();
        int last = iif;
        for (int j = iif; j < i_hi; j += kClSize) {
          if (j != iif) a.trace[(size_t)item_geom(a, frame, j).item * 8 + 3] = t;
          if (j != iif) a.trace[(size_t)item_geom(a, frame, j).item * 8 + 4] = ctx->count;
          last = j;
        }
        if (last != iif) {
          unsigned smid;
          asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
          a.trace[(size_t)item_geom(a, frame, last).item * 8 + 7] = smid;
          a.trace[(size_t)item_geom(a, frame, last).item * 8 + 6] = t_kernel;
          a.trace[(size_t)item_geom(a, frame, last).item * 8 + 3] = t;
        }
      }
    }

    if (a.fuse) {
      // ---- fused tail (parts == 1).  The published threshold lags behind the final k-th best, so the eight lists
      // together hold several times k candidates.  One last scan of the (now complete) histogram gives the bin that
      // holds the k-th best; everything below its lower edge is dropped, which leaves k plus a handful.  Then the
      // ranking is shared out: every CTA copies the eight pruned lists into its own shared memory and ranks ITS OWN
      // entries against them with 16 threads per entry — the rank of a key among distinct keys is its output slot, so
      // there is no selection pass and no sort, and only O(1) barriers.  Rank 0 adds the zero-score fillers and the
      // threshold count.
      __shared__ unsigned long long s_keyT;
      __shared__ int s_total, s_nge;
      __shared__ uint32_t s_flags[kFuseMaxK];  // rank 0: flat indices < k taken by a selected peak (for the fillers)
      __shared__ int s_rank[kFuseMaxK];
      const int n = ctx->count;
      // (debug trace, tools/tail_trace.py: tail stamps go to unused columns of the CTA's second and third item rows)
      auto tail_stamp = [&](int which) {
        if (a.trace && tid == 0 && iif + 2 * kClSize < i_hi)
          a.trace[(size_t)item_geom(a, frame, iif + (which < 3 ? 1 : 2) * kClSize).item * 8 + (which % 3 == 0 ? 1 : 5 + which % 3)] = now();
      };
      convert_entries<MODE>(ctx, list, n);
      for (int i = tid; i < a.k; i += kTileThreads) s_flags[i] = 0u;
      if (tid == 0) s_nge = 0;
      cluster.sync();  // (3a) every CTA has finished streaming and binning; rank 0's flags are clear
      tail_stamp(0);
      if (tid < 32) {
        const int bin = cl_scan_bin(cluster, cc, a.k);
        if (tid == 0) {
          unsigned long long kt = 1ull;  // fewer than k candidates binned: keep everything that is non-zero
          if (bin >= 0) {
            const float edge = cl_window_edge(bin);
            // at least k candidates have values >= edge; in SIGMOID_PEAK mode the list holds their sigmoids, whose
            // rounding (<= 2 ulp) the relative guard band covers
            const float lowest = MODE == TAUV_TOPK_SIGMOID_PEAK ? sigmoid_ref(edge) * (1.0f - 4e-5f) : edge;
            kt = (unsigned long long)float_to_key(lowest) << 32;
            if (kt == 0ull) kt = 1ull;
          }
          s_keyT = kt;
        }
      }
      __syncthreads();
      {
        const unsigned long long kt = s_keyT;
        compact_list(ctx, list, n, [&](unsigned long long c) { return c >= kt && c != 0ull; });
      }
      if (ctx->base > a.k) {  // (ties / plateaus: more than k survive in this CTA alone; the pool holds 8k keys)
        const int n2 = ctx->base;
        const unsigned long long T = block_kth_largest<kTileThreads>([&](int i) { return list[i]; }, n2, a.k, hist, ctx->sel);
        compact_list(ctx, list, n2, [&](unsigned long long c) { return c >= T; });
      }
      if (tid == 0) ctx->emit = (uint32_t)ctx->base;
      tail_stamp(1);
      cluster.sync();  // (3b) all eight pruned lists are final
      tail_stamp(2);
      unsigned long long* pool = reinterpret_cast<unsigned long long*>(hist);  // hist + tile: contiguous, >= 8k keys
      {
        // the eight counts first (independent remote loads: one DSMEM round trip, not eight in a chain), then all lists
        int cnt[kClSize], total_n = 0;
#pragma unroll
        for (int r = 0; r < kClSize; ++r) cnt[r] = (int)*cluster.map_shared_rank(&ctx->emit, r);
#pragma unroll
        for (int r = 0; r < kClSize; ++r) {
          const unsigned long long* rl = cluster.map_shared_rank(list, r);
          for (int i = tid; i < cnt[r]; i += kTileThreads) pool[total_n + i] = rl[i];
          total_n += cnt[r];
        }
        if (tid == 0) s_total = total_n;
      }
      const int n_own = (int)ctx->emit;
      for (int i = tid; i < n_own; i += kTileThreads) s_rank[i] = 0;
      __syncthreads();
      {
        // 16 threads per own entry, each over a sixteenth of the pool; partial ranks meet in shared memory
        const int total = s_total;
        const int part = tid & 15;
        for (int e = tid >> 4; e < n_own; e += kTileThreads / 16) {
          const unsigned long long c = list[e];
          int r = 0;
          for (int j = part; j < total; j += 16) r += (pool[j] > c);
          if (r) atomicAdd(&s_rank[e], r);
        }
      }
      __syncthreads();
      tail_stamp(3);
      {
        const BoxArgs& g = a.box;
        const uint32_t hw_elems = (uint32_t)(a.H * a.W);  // (the cluster path requires W <= 1016 and C*H*W < 2^32)
        int my_ge = 0;
        for (int i = tid; i < n_own; i += kTileThreads) {
          const int r = s_rank[i];
          if (r < a.k) {
            const unsigned long long c = list[i];
            const uint32_t flat = composite_idx(c);
            const float sc = key_to_float(composite_key(c));
            const uint32_t lab = flat / hw_elems;  // 32-bit: a 64-bit divide is ~100 dependent instructions
            const uint32_t rem = flat - lab * hw_elems;
            const int iy = (int)(rem / (uint32_t)a.W), ix = (int)(rem - (uint32_t)iy * (uint32_t)a.W);
            const long long slot = (long long)frame * a.k + r;
            a.out_index[slot * 2 + 0] = iy;
            a.out_index[slot * 2 + 1] = ix;
            a.out_label[slot] = lab;
            a.out_score[slot] = sc;
            if (g.enabled) {
              box_one(g, frame, slot, iy, ix);
              if (!(sc < g.thr)) ++my_ge;
            }
            if (flat < (uint32_t)a.k) *cluster.map_shared_rank(&s_flags[flat], 0) = 1u;
          }
        }
        // one remote add per warp, not per thread (remote shared-memory atomics cost ~8 ns each at the target)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_ge += __shfl_xor_sync(0xffffffffu, my_ge, o);
        if ((tid & 31) == 0 && my_ge) atomicAdd(cluster.map_shared_rank(&s_nge, 0), my_ge);
      }
      tail_stamp(4);
      cluster.sync();  // (3) nobody touches this unit's distributed state any more; rank 0 sees flags and count
      tail_stamp(5);
      if (rank == 0) {
        const int npos = min(a.k, s_total);
        if (MODE == TAUV_TOPK_SIGMOID_PEAK && npos < a.k)
          topk_emit_fillers<kTileThreads>(s_flags, npos, frame, a.k, a.H, a.W, a.out_index, a.out_label, a.out_score, a.box);
        if (a.box.enabled && tid == 0) {  // entries before the first score < threshold (ranked scores descend; fillers score 0)
          int cnt = s_nge;
          if (MODE == TAUV_TOPK_SIGMOID_PEAK && npos < a.k && !(0.0f < a.box.thr)) cnt += a.k - npos;
          a.box.count[frame] = cnt;
        }
        __syncthreads();
      }
      continue;
    }
    // ---- the CTA's candidates of the whole unit: exact top-k, one row of the candidate table
    {
      const int row = frame * a.rows_per_frame + part * kClSize + rank;
      if (part == 0 && rank == 0)  // rows of the frame that no CTA owns hold no candidates
        for (int r = parts * kClSize + tid; r < a.rows_per_frame; r += kTileThreads) a.cand_count[frame * a.rows_per_frame + r] = 0;
      const int n = ctx->count;
      if (n == 0) {
        if (tid == 0) a.cand_count[row] = 0;
      } else {
        convert_entries<MODE>(ctx, list, n);
        const unsigned long long T = block_kth_largest<kTileThreads>([&](int i) { return list[i]; }, n, a.k, hist, ctx->sel);
        if (tid == 0) ctx->emit = 0;
        __syncthreads();
        unsigned long long* out = a.cand + (size_t)row * a.k;
        for (int i = tid; i < n; i += kTileThreads) {
          const unsigned long long c = list[i];
          if (c >= T && c != 0ull) out[atomicAdd(&ctx->emit, 1u)] = c;
        }
        __syncthreads();
        if (tid == 0) a.cand_count[row] = (int)ctx->emit;
      }
    }
    cluster.sync();  // (3) nobody touches this unit's distributed state any more
  }
}

// ----------------------------------------------------------------------------------------------
// K2: per-frame merge (+ optional box decode) of the candidate table
// ----------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kMergeThreads) merge_kernel(
    const unsigned long long* __restrict__ cand, const int* __restrict__ cand_count,
    int items_per_frame, int k, int H, int W, int pool_cap, int64_t* __restrict__ index,
    int64_t* __restrict__ label, float* __restrict__ score, BoxArgs g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  int p2 = 1;
  while (p2 < k) p2 <<= 1;
  unsigned long long* sel = reinterpret_cast<unsigned long long*>(smem_raw);                 // [p2]
  unsigned long long* pool = sel + p2;                                                        // [pool_cap]
  uint32_t* hist = reinterpret_cast<uint32_t*>(pool + pool_cap);                              // [max(2048,k)]
  __shared__ uint32_t ctl[8];
  __shared__ int s_total, s_pool_n;

  const unsigned long long* fc = cand + (size_t)b * items_per_frame * k;
  const int* cnt = cand_count + (size_t)b * items_per_frame;
  const int nslots = items_per_frame * k;

  if (tid == 0) {
    s_total = 0;
    s_pool_n = 0;
  }
  __syncthreads();
  int part = 0;
  for (int i = tid; i < items_per_frame; i += kMergeThreads) part += cnt[i];
  if (part) atomicAdd(&s_total, part);
  __syncthreads();
  const int total = s_total;

  int npos;
  if (total <= pool_cap) {
    // usual case (items reject most of their candidates against the frame threshold): pull the frame's
    // candidates into shared memory once, one warp per row, then select there
    for (int it = warp; it < items_per_frame; it += kMergeThreads / 32) {
      const int c = cnt[it];
      if (c == 0) continue;
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_pool_n, c);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (int i = lane; i < c; i += 32) pool[base + i] = fc[(size_t)it * k + i];
    }
    __syncthreads();
    npos = topk_select_pool<kMergeThreads>(pool, total, k, p2, sel, hist, ctl);
  } else {
    // candidates live in a padded [rows][k] table: slot i is valid iff (i % k) < cnt[i / k]
    auto load = [&](int i) -> unsigned long long {
      const int it = i / k;
      return (i - it * k) < cnt[it] ? fc[i] : 0ull;  // 0 never beats a real composite
    };
    for (int i = tid; i < p2; i += kMergeThreads) sel[i] = 0ull;
    if (tid == 0) ctl[5] = 0;
    __syncthreads();
    unsigned long long T = 1ull;
    if (total > k) T = block_kth_largest<kMergeThreads>(load, nslots, k, hist, ctl);
    for (int i = tid; i < nslots; i += kMergeThreads) {
      const unsigned long long c = load(i);
      if (c >= T && c != 0ull) sel[atomicAdd(&ctl[5], 1u)] = c;
    }
    __syncthreads();
    npos = (int)ctl[5];
  }
  topk_emit_ranked<MODE, kMergeThreads>(sel, p2, npos, hist, b, k, H, W, index, label, score, g);
}

// Stand-alone box stage for callers that already hold index/score (tauv_centernet_boxes).
__global__ void boxes_kernel(const int64_t* __restrict__ index, const float* __restrict__ score,
                             int k, BoxArgs g) {
  const int b = blockIdx.x;
  __shared__ int s_first_below;
  if (threadIdx.x == 0) s_first_below = k;
  __syncthreads();
  for (int r = threadIdx.x; r < k; r += blockDim.x) {
    const long long slot = (long long)b * k + r;
    box_one(g, b, slot, (int)index[slot * 2], (int)index[slot * 2 + 1]);
    if (score[slot] < g.thr) atomicMin(&s_first_below, r);
  }
  __syncthreads();
  if (threadIdx.x == 0) g.count[b] = s_first_below;
}

// ----------------------------------------------------------------------------------------------
// heatmap_nms (dense, for the drop-in signature; the fused path never materialises this)
// ----------------------------------------------------------------------------------------------
template <bool SIG>
__global__ void __launch_bounds__(256) heatmap_nms_kernel(const float* __restrict__ in,
                                                          float* __restrict__ out, long long planes,
                                                          int H, int W, int rad) {
  // one thread per element; neighbours come from L1/L2 (each line is re-used 9x within a CTA)
  const long long n = planes * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const float* pl = in + (i - (long long)y * W - x);
    float c = pl[(long long)y * W + x];
    if (SIG) c = sigmoid_ref(c);
    bool is_max = true;  // also reproduces NaN: NaN centre -> (max == c) false -> 0*NaN = NaN
    for (int dy = -rad; dy <= rad && is_max; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = -rad; dx <= rad; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        float v = pl[(long long)yy * W + xx];
        if (SIG) v = sigmoid_ref(v);
        if (v > c || v != v) {
          is_max = false;
          break;
        }
      }
    }
    // (max == h).float() * h   (decode.py:252)
    out[i] = (c != c) ? c : (is_max ? c : __fmul_rn(0.0f, c));
  }
}

__global__ void gather_at_kernel(const float* __restrict__ src, long long sb, long long ssel,
                                 long long sc, long long sy, long long sx, int nch,
                                 const int64_t* __restrict__ index,
                                 const int64_t* __restrict__ label, long long n,
                                 int k, float* __restrict__ out) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n * nch) return;
  const long long j = t / nch;
  const int c = (int)(t - j * nch);
  const long long b = j / k;
  const long long sel = label ? label[j] : 0;
  out[t] = src[b * sb + sel * ssel + c * sc + index[j * 2] * sy + index[j * 2 + 1] * sx];
}

// backward of gather_at (label == NULL): dst[b, c, iy, ix] += grad[b, j, c].  No atomics: the FIRST object of a frame
// that sits on a cell adds up, in object order, the gradients of every object on that cell (two objects on one cell
// happen: loss.py:196-227 gathers at out_index_for_position of every object), the later ones do nothing.
__global__ void scatter_add_at_kernel(const float* __restrict__ grad, const int64_t* __restrict__ index, long long n, int k,
                                      int nch, float* __restrict__ dst, long long sb, long long sc, long long sy,
                                      long long sx) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n * nch) return;
  const long long j = t / nch;
  const int c = (int)(t - j * nch);
  const long long b = j / k, j0 = b * k;
  const long long iy = index[j * 2], ix = index[j * 2 + 1];
  for (long long q = j0; q < j; ++q)
    if (index[q * 2] == iy && index[q * 2 + 1] == ix) return;
  float acc = grad[t];
  for (long long q = j + 1; q < j0 + k; ++q)
    if (index[q * 2] == iy && index[q * 2 + 1] == ix) acc += grad[q * nch + c];
  dst[b * sb + c * sc + iy * sy + ix * sx] = acc;
}

}  // namespace tauv

#include "centernet_select.cuh"  // round 2: block maxima + select (the default decode path)

namespace tauv {

// ----------------------------------------------------------------------------------------------
// Host entry points
// ----------------------------------------------------------------------------------------------
static int check_topk_shape(int B, int C, int H, int W, int k) {
  TAUV_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, TAUV_E_SHAPE, "heatmap shape [%d,%d,%d,%d] must be positive", B, C, H, W);
  TAUV_REQUIRE(k > 0, TAUV_E_SHAPE, "k=%d must be positive", k);
  const long long chw = (long long)C * H * W;
  TAUV_REQUIRE((long long)k <= chw, TAUV_E_K_RANGE, "selected index k out of range (k=%d > C*H*W=%lld)", k, chw);
  TAUV_REQUIRE(chw < (1LL << 32), TAUV_E_UNSUPPORTED, "C*H*W=%lld does not fit the 32-bit flat index", chw);
  TAUV_REQUIRE(k <= kMaxK, TAUV_E_UNSUPPORTED, "k=%d exceeds the built-in limit %d", k, kMaxK);
  TAUV_REQUIRE(W <= 4096, TAUV_E_UNSUPPORTED, "W=%d exceeds the built-in limit 4096", W);
  return 0;
}

#ifdef TAUV_DEBUG
static long long* g_debug_trace = nullptr;  // experiment hook (tools/tile_trace.py); only in -DTAUV_DEBUG builds
#endif

static int plan_and_check(const float* hm, int B, int C, int H, int W, int k, void* ws, size_t ws_bytes, TopkPlan* p,
                          TileArgs* a) {
  make_plan(B, C, H, W, k, hm, p);
  TAUV_REQUIRE(ws != nullptr && (uintptr_t)ws % 256 == 0, TAUV_E_WORKSPACE, "workspace must be 256-byte aligned");
  TAUV_REQUIRE(ws_bytes >= p->cand_bytes + p->count_bytes + p->state_bytes, TAUV_E_WORKSPACE,
               "workspace %zu < required %zu", ws_bytes, p->cand_bytes + p->count_bytes + p->state_bytes);
  TAUV_REQUIRE(p->smem_bytes <= 227 * 1024, TAUV_E_UNSUPPORTED, "tile needs %zu B shared memory", p->smem_bytes);
  a->hm = hm; a->B = B; a->C = C; a->H = H; a->W = W; a->k = k;
  a->rows_per_item = p->rows_per_item; a->items_per_plane = p->items_per_plane; a->rows_per_frame = p->rows_per_frame;
  a->cap = p->cap; a->soft = p->soft; a->sub_elems = p->sub_elems;
  a->cand = reinterpret_cast<unsigned long long*>(ws);
  a->cand_count = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(ws) + p->cand_bytes);
  a->frame_state = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(ws) + p->cand_bytes + p->count_bytes);
#ifdef TAUV_DEBUG
  a->trace = g_debug_trace;
#else
  a->trace = nullptr;
#endif
  a->fuse = 0; a->out_index = nullptr; a->out_label = nullptr; a->out_score = nullptr;
  a->box = BoxArgs{};
  const long long items = (long long)B * p->items_per_frame;
  TAUV_REQUIRE(items < (1LL << 31), TAUV_E_UNSUPPORTED, "too many items (%lld)", items);
  return 0;
}

// What the fused tail of the cluster kernel writes (run_topk hands this in; the two-stage API never fuses).
struct FuseOut {
  int64_t* index;
  int64_t* label;
  float* score;
  const BoxArgs* box;
};

// stage 1: per-item candidates into the workspace — or, when `fo` is given and the launch qualifies (whole-frame
// units, 8k candidates fit the list), the complete ranked output (*fused = true: stage 2 must be skipped)
static int run_stage1(const float* hm, int B, int C, int H, int W, int k, int mode, void* ws, size_t ws_bytes,
                      cudaStream_t st, const FuseOut* fo = nullptr, bool* fused = nullptr) {
  if (fused) *fused = false;
  TopkPlan p;
  TileArgs a;
  if (int e = plan_and_check(hm, B, C, H, W, k, ws, ws_bytes, &p, &a)) return e;
  const long long items = (long long)B * p.items_per_frame;
  if (!(p.vec && W <= kClMaxW)) {
    // scalar path (W % 4 != 0, unaligned base) and very wide maps: one CTA per item against per-frame state in the
    // workspace, zeroed here (key 0 = "no threshold published yet")
    void (*kern)(const TileArgs) = nullptr;
    if (mode == TAUV_TOPK_SIGMOID_PEAK) kern = p.vec ? tile_topk_kernel<1, true> : tile_topk_kernel<1, false>;
    else kern = p.vec ? tile_topk_kernel<0, true> : tile_topk_kernel<0, false>;
    TAUV_CUDA(ensure_dynamic_smem((const void*)kern, p.smem_bytes));
    TAUV_CUDA(cudaMemsetAsync(a.cand_count, 0, p.count_bytes + p.state_bytes, st));  // (adjacent in the workspace)
    kern<<<(unsigned)items, kTileThreads, p.smem_bytes, st>>>(a);
    TAUV_LAUNCH_CHECK("tile_topk_kernel");
    return 0;
  }
  // vectorised path: persistent clusters of 8 CTAs, one unit (a frame, or a share of a frame's items) at a time
  void (*ck)(const TileArgs, int, int) = mode == TAUV_TOPK_SIGMOID_PEAK ? tile_cluster_kernel<1> : tile_cluster_kernel<0>;
  const size_t csmem = (size_t)kClOffList + p.smem_bytes + (size_t)(kBootElems + 2 * W + 8) * 4;
  TAUV_REQUIRE(csmem <= 227 * 1024, TAUV_E_UNSUPPORTED, "tile needs %zu B shared memory", csmem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kClSize);
  cfg.blockDim = dim3(kTileThreads);
  cfg.dynamicSmemBytes = csmem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kClSize;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int ncl = 0;  // clusters that are resident at once (the query costs tens of microseconds of host time: cached)
  {
    // per (mode, device): the largest dynamic shared-memory size the kernel has been opted in to (only ever raised,
    // under a lock, so concurrent callers cannot lower it under each other), and the last occupancy answer
    static std::mutex mu;
    static std::atomic<unsigned long long> max_smem[2][64], occ[2][64];
    int dev = 0;
    TAUV_CUDA(cudaGetDevice(&dev));
    const int m = mode == TAUV_TOPK_SIGMOID_PEAK, d = (dev >= 0 && dev < 64) ? dev : 63;
    if (max_smem[m][d].load(std::memory_order_acquire) < csmem || dev != d) {
      std::lock_guard<std::mutex> lock(mu);
      if (max_smem[m][d].load(std::memory_order_relaxed) < csmem || dev != d) {
        TAUV_CUDA(cudaFuncSetAttribute(ck, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
        if (dev == d) max_smem[m][d].store(csmem, std::memory_order_release);
      }
    }
    const unsigned long long seen = occ[m][d].load(std::memory_order_relaxed);  // (csmem + 1) << 16 | ncl; idempotent
    if (dev == d && (seen >> 16) == (unsigned long long)csmem + 1) {
      ncl = (int)(seen & 0xffff);
    } else {
      TAUV_CUDA(cudaOccupancyMaxActiveClusters(&ncl, ck, &cfg));
      occ[m][d].store((((unsigned long long)csmem + 1) << 16) | (unsigned long long)(ncl & 0xffff), std::memory_order_relaxed);
    }
  }
  TAUV_REQUIRE(ncl >= 1, TAUV_E_UNSUPPORTED, "no cluster of %d CTAs fits the device with %zu B shared memory", kClSize, csmem);
  // units: whole frames when there are at least as many frames as clusters, otherwise every frame is split into
  // `parts` contiguous shares of its items (each share keeps its own threshold; the merge kernel joins them)
  int parts = 1;
  if (B < ncl) {
    parts = ncl / B;
    const int by_items = p.items_per_frame / (2 * kClSize);  // at least two items per CTA
    if (parts > by_items) parts = by_items;
    if (parts > 8) parts = 8;  // (every part pays its own bootstrap and adds 8 rows to the merge; 64 CTAs per frame suffice)
    if (parts < 1) parts = 1;
  }
  const long long n_units = (long long)B * parts;
  const long long ncl_used = n_units < ncl ? n_units : ncl;
  int p2 = 1;
  while (p2 < k) p2 <<= 1;
  if (fo && fused && parts == 1 && k <= kFuseMaxK &&
      (size_t)kClSize * k * 8 <= (size_t)kRadixBins * 4 + (size_t)(kBootElems + 2 * W + 8) * 4 && !debug_env("TAUV_NO_FUSE")) {
    a.fuse = 1;
    a.out_index = fo->index;
    a.out_label = fo->label;
    a.out_score = fo->score;
    a.box = *fo->box;
    *fused = true;
  }
  cfg.gridDim = dim3((unsigned)(ncl_used * kClSize));
  TAUV_CUDA(cudaLaunchKernelEx(&cfg, ck, a, (int)n_units, parts));
  return 0;
}

// stage 2: per-frame merge of the workspace candidates (+ boxes)
static int run_stage2(int B, int C, int H, int W, int k, int mode, int64_t* index, int64_t* label, float* score,
                      const BoxArgs& box, void* ws, size_t ws_bytes, cudaStream_t st) {
  TopkPlan p;
  TileArgs a;
  if (int e = plan_and_check(nullptr, B, C, H, W, k, ws, ws_bytes, &p, &a)) return e;
  int p2 = 1;
  while (p2 < k) p2 <<= 1;
  long long pool_cap = (long long)p.rows_per_frame * k;
  if (pool_cap > 12288) pool_cap = 12288;
  const size_t msmem = (size_t)p2 * 8 + (size_t)pool_cap * 8 + (size_t)(kRadixBins > k ? kRadixBins : k) * 4;
  if (mode == TAUV_TOPK_SIGMOID_PEAK) {
    TAUV_CUDA(ensure_dynamic_smem((const void*)merge_kernel<1>, msmem));
    merge_kernel<1><<<B, kMergeThreads, msmem, st>>>(a.cand, a.cand_count, p.rows_per_frame, k, H, W, (int)pool_cap,
                                                     index, label, score, box);
  } else {
    TAUV_CUDA(ensure_dynamic_smem((const void*)merge_kernel<0>, msmem));
    merge_kernel<0><<<B, kMergeThreads, msmem, st>>>(a.cand, a.cand_count, p.rows_per_frame, k, H, W, (int)pool_cap,
                                                     index, label, score, box);
  }
  TAUV_LAUNCH_CHECK("merge_kernel");
  return 0;
}

static int run_topk(const float* hm, int B, int C, int H, int W, int k, int mode, int64_t* index, int64_t* label,
                    float* score, const BoxArgs& box, void* ws, size_t ws_bytes, cudaStream_t st) {
  SelPlan sp;
  if (mode == TAUV_TOPK_SIGMOID_PEAK && (uintptr_t)hm % 16 == 0 && select_plan(B, C, H, W, k, &sp) &&
      !debug_env("TAUV_OLD_DECODE"))
    return run_select_decode(hm, B, C, H, W, k, index, label, score, box, ws, ws_bytes, sp, st);
  const FuseOut fo{index, label, score, &box};
  bool fused = false;
  if (int e = run_stage1(hm, B, C, H, W, k, mode, ws, ws_bytes, st, &fo, &fused)) return e;
  if (fused) return 0;
  return run_stage2(B, C, H, W, k, mode, index, label, score, box, ws, ws_bytes, st);
}

static int fill_box_args(BoxArgs* g, const float* size, const int64_t* ss, const float* offset, const int64_t* os,
                         const float* depth, const int64_t* ds, int mode, int ratio, int in_h, int in_w, int out_h,
                         int out_w, float thr, double* yx, float* hw, float* depth_out, int32_t* count) {
  TAUV_REQUIRE(size && ss && yx && hw && count, TAUV_E_NULL, "size/size_strides/yx/hw/count must not be NULL");
  TAUV_REQUIRE(mode == TAUV_BOX_DECODE || mode == TAUV_BOX_KEYPOINTS, TAUV_E_SHAPE, "bad box mode %d", mode);
  if (mode == TAUV_BOX_DECODE) TAUV_REQUIRE(offset && os, TAUV_E_NULL, "offset/offset_strides must not be NULL in decode mode");
  if (depth) TAUV_REQUIRE(ds && depth_out, TAUV_E_NULL, "depth given without strides / output");
  TAUV_REQUIRE(ratio > 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, TAUV_E_SHAPE, "bad model geometry");
  g->enabled = 1;
  g->size = size;
  for (int i = 0; i < 4; ++i) g->ss[i] = ss[i];
  g->offset = offset;
  for (int i = 0; i < 4; ++i) g->os[i] = os ? os[i] : 0;
  g->depth = depth;
  for (int i = 0; i < 3; ++i) g->ds[i] = ds ? ds[i] : 0;
  g->mode = mode; g->ratio = ratio; g->in_h = in_h; g->in_w = in_w; g->out_h = out_h; g->out_w = out_w;
  g->thr = thr; g->yx = yx; g->hw = hw; g->depth_out = depth_out; g->count = count;
  return 0;
}

}  // namespace tauv

using namespace tauv;

extern "C" size_t tauv_heatmap_topk_workspace_bytes(int B, int C, int H, int W, int k) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || k <= 0) return 0;
  TopkPlan p;
  // alignment only affects the load path, never the sizes
  make_plan(B, C, H, W, k, nullptr, &p);
  size_t need = p.cand_bytes + p.count_bytes + p.state_bytes;
  SelPlan sp;
  if (select_plan(B, C, H, W, k, &sp) && sp.bm_bytes + sp.bm2_bytes > need) need = sp.bm_bytes + sp.bm2_bytes;
  return need;
}

extern "C" int tauv_heatmap_topk(const float* heatmap, int B, int C, int H, int W, int k, int mode, int64_t* index,
                                 int64_t* label, float* score, void* workspace, size_t workspace_bytes,
                                 tauv_stream_t stream) {
  TAUV_REQUIRE(heatmap && index && label && score, TAUV_E_NULL, "heatmap/index/label/score must not be NULL");
  TAUV_REQUIRE(mode == TAUV_TOPK_RAW || mode == TAUV_TOPK_SIGMOID_PEAK, TAUV_E_SHAPE, "bad top-k mode %d", mode);
  if (int e = check_topk_shape(B, C, H, W, k)) return e;
  BoxArgs none{};
  none.enabled = 0;
  return run_topk(heatmap, B, C, H, W, k, mode, index, label, score, none, workspace, workspace_bytes,
                  (cudaStream_t)stream);
}

extern "C" int tauv_centernet_boxes(const int64_t* index, const float* score, int B, int k, int H, int W,
                                    const float* size, const int64_t size_strides[4], const float* offset,
                                    const int64_t offset_strides[4], const float* depth,
                                    const int64_t depth_strides[3], int mode, int downsample_ratio, int in_h,
                                    int in_w, int out_h, int out_w, float score_threshold, double* yx, float* hw,
                                    float* depth_out, int32_t* count, tauv_stream_t stream) {
  TAUV_REQUIRE(index && score, TAUV_E_NULL, "index/score must not be NULL");
  TAUV_REQUIRE(B > 0 && k > 0 && H > 0 && W > 0, TAUV_E_SHAPE, "bad shape");
  BoxArgs g{};
  if (int e = fill_box_args(&g, size, size_strides, offset, offset_strides, depth, depth_strides, mode,
                            downsample_ratio, in_h, in_w, out_h, out_w, score_threshold, yx, hw, depth_out, count))
    return e;
  boxes_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(index, score, k, g);
  TAUV_LAUNCH_CHECK("boxes_kernel");
  return 0;
}

extern "C" int tauv_centernet_decode(const float* heatmap_logits, int B, int C, int H, int W, int k, const float* size,
                                     const int64_t size_strides[4], const float* offset,
                                     const int64_t offset_strides[4], const float* depth,
                                     const int64_t depth_strides[3], int mode, int downsample_ratio, int in_h,
                                     int in_w, float score_threshold, int64_t* index, int64_t* label, float* score,
                                     double* yx, float* hw, float* depth_out, int32_t* count, void* workspace,
                                     size_t workspace_bytes, tauv_stream_t stream) {
  TAUV_REQUIRE(heatmap_logits && index && label && score, TAUV_E_NULL, "heatmap/index/label/score must not be NULL");
  if (int e = check_topk_shape(B, C, H, W, k)) return e;
  BoxArgs g{};
  if (int e = fill_box_args(&g, size, size_strides, offset, offset_strides, depth, depth_strides, mode,
                            downsample_ratio, in_h, in_w, H, W, score_threshold, yx, hw, depth_out, count))
    return e;
  return run_topk(heatmap_logits, B, C, H, W, k, TAUV_TOPK_SIGMOID_PEAK, index, label, score, g, workspace,
                  workspace_bytes, (cudaStream_t)stream);
}

extern "C" int tauv_centernet_block_maxima(const float* heatmap_logits, int B, int C, int H, int W, int k, void* workspace,
                                           size_t workspace_bytes, tauv_stream_t stream) {
  TAUV_REQUIRE(heatmap_logits, TAUV_E_NULL, "heatmap must not be NULL");
  if (int e = check_topk_shape(B, C, H, W, k)) return e;
  SelPlan sp;
  TAUV_REQUIRE((uintptr_t)heatmap_logits % 16 == 0 && select_plan(B, C, H, W, k, &sp), TAUV_E_UNSUPPORTED,
               "the block-maxima path needs a 16-byte aligned map with W %% 4 == 0 and k <= %d", kSelMaxK);
  return run_block_maxima(heatmap_logits, B, C, H, W, workspace, workspace_bytes, sp, (cudaStream_t)stream);
}

extern "C" int tauv_heatmap_nms(const float* in, float* out, int B, int C, int H, int W, int kernel_size,
                                int apply_sigmoid, tauv_stream_t stream) {
  TAUV_REQUIRE(in && out, TAUV_E_NULL, "in/out must not be NULL");
  TAUV_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, TAUV_E_SHAPE, "heatmap shape must be positive");
  TAUV_REQUIRE(kernel_size >= 1 && kernel_size % 2 == 1, TAUV_E_KERNEL, "kernel_size=%d must be odd and >= 1", kernel_size);
  const long long planes = (long long)B * C;
  const long long n = planes * H * W;
  long long blocks = (n + 255) / 256;
  const long long maxb = (long long)num_sms() * 32;
  if (blocks > maxb) blocks = maxb;
  if (apply_sigmoid)
    heatmap_nms_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, planes, H, W, kernel_size / 2);
  else
    heatmap_nms_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, planes, H, W, kernel_size / 2);
  TAUV_LAUNCH_CHECK("heatmap_nms_kernel");
  return 0;
}

extern "C" int tauv_gather_at(const float* src, int64_t sb, int64_t ssel, int64_t sc, int64_t sy, int64_t sx, int nch,
                              const int64_t* index, const int64_t* label, int B, int k, float* out,
                              tauv_stream_t stream) {
  TAUV_REQUIRE(src && index && out, TAUV_E_NULL, "src/index/out must not be NULL");
  TAUV_REQUIRE(B > 0 && k > 0 && nch > 0, TAUV_E_SHAPE, "bad shape");
  const long long n = (long long)B * k;
  const long long tot = n * nch;
  gather_at_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, sb, ssel, sc, sy, sx, nch,
                                                                                    index, label, n, k, out);
  TAUV_LAUNCH_CHECK("gather_at_kernel");
  return 0;
}

extern "C" int tauv_scatter_add_at(const float* grad, const int64_t* index, int B, int k, int nch, float* dst, int64_t sb,
                                   int64_t sc, int64_t sy, int64_t sx, tauv_stream_t stream) {
  TAUV_REQUIRE(grad && index && dst, TAUV_E_NULL, "grad/index/dst must not be NULL");
  TAUV_REQUIRE(B > 0 && k > 0 && nch > 0, TAUV_E_SHAPE, "bad shape");
  const long long n = (long long)B * k;
  const long long tot = n * nch;
  scatter_add_at_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(grad, index, n, k, nch, dst, sb, sc,
                                                                                         sy, sx);
  TAUV_LAUNCH_CHECK("scatter_add_at_kernel");
  return 0;
}

extern "C" int tauv_heatmap_topk_stage1(const float* heatmap, int B, int C, int H, int W, int k, int mode,
                                        void* workspace, size_t workspace_bytes, tauv_stream_t stream) {
  TAUV_REQUIRE(heatmap, TAUV_E_NULL, "heatmap must not be NULL");
  TAUV_REQUIRE(mode == TAUV_TOPK_RAW || mode == TAUV_TOPK_SIGMOID_PEAK, TAUV_E_SHAPE, "bad top-k mode %d", mode);
  if (int e = check_topk_shape(B, C, H, W, k)) return e;
  return run_stage1(heatmap, B, C, H, W, k, mode, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int tauv_centernet_decode_stage2(int B, int C, int H, int W, int k, const float* size,
                                            const int64_t size_strides[4], const float* offset,
                                            const int64_t offset_strides[4], const float* depth,
                                            const int64_t depth_strides[3], int mode, int downsample_ratio, int in_h,
                                            int in_w, float score_threshold, int64_t* index, int64_t* label,
                                            float* score, double* yx, float* hw, float* depth_out, int32_t* count,
                                            void* workspace, size_t workspace_bytes, tauv_stream_t stream) {
  TAUV_REQUIRE(index && label && score, TAUV_E_NULL, "index/label/score must not be NULL");
  if (int e = check_topk_shape(B, C, H, W, k)) return e;
  BoxArgs g{};
  if (int e = fill_box_args(&g, size, size_strides, offset, offset_strides, depth, depth_strides, mode,
                            downsample_ratio, in_h, in_w, H, W, score_threshold, yx, hw, depth_out, count))
    return e;
  return run_stage2(B, C, H, W, k, TAUV_TOPK_SIGMOID_PEAK, index, label, score, g, workspace, workspace_bytes,
                    (cudaStream_t)stream);
}

#ifdef TAUV_DEBUG
// Debug hook for tools/tile_trace.py (only in -DTAUV_DEBUG builds; the default library exports nothing the header does
// not declare): per-item timestamps of the next tile_topk launches land in `buf` (8 int64 per item; NULL = off).
extern "C" void tauv_debug_tile_trace(long long* buf) { tauv::g_debug_trace = buf; }
// the same for select_kernel (tools/select_trace.py): 16 int64 per frame
extern "C" void tauv_debug_select_trace(long long* buf) { tauv::g_sel_trace = buf; }
#endif
