// centernet_stream.cuh — the CenterNet decode as ONE persistent streaming kernel (round 2).  Included by
// centernet_decode.cu after the shared helpers (keys, sigmoid_tie, reject_key_for_score, window bins, box_one).
//
// Replaces, in one launch (reference file:line under src/tauv_vision/centernet/model/):
//   decode.py:182 sigmoid · :239-252 heatmap_nms (3x3, plateaus survive) · :255-279 heatmap_detect (joint top-k) ·
//   :204-234 per-detection gather + box arithmetic (:87-88 / :65 for decode_keypoints).
//
// Why this shape (measured, profiles/r2_stream_bench_v*.txt): a B200 streams a read-once 335 MB tensor at
// 6.2-6.4 TB/s through a cp.async.bulk (1-D TMA) shared-memory ring with one CTA per SM and 96-128 KB per SM in
// flight — as fast as register loads, but the bytes in flight do not depend on what the consuming warps are doing,
// the 3x3 neighbourhood of every cell is already in shared memory (no re-reads: DRAM traffic = algorithmic bytes),
// and an L2 evict-first policy on the copies keeps the previous kernel's dirty lines from stalling the stream.
//
// Work split: the B*C*H rows of the batch are cut into G equal contiguous ranges, one per CTA (G = SMs x CTAs/SM):
// every SM moves the same number of bytes whatever B is.  A range crosses at most a few frame boundaries; the part
// of a range inside one frame is a RUN.  Per run the CTA keeps, in shared memory, a candidate list (64-bit composite
// keys: score key << 32 | ~flat index, so plain descending order = score desc, index asc — the order the reference's
// own KAT asserts, decode.py:327-339), a 2048-bin histogram of the candidates' logits, and a rejection threshold
// derived from it (the k-th best candidate so far bounds the frame's k-th best from below).  Cells below the
// threshold cost one max + compare per 128-bit strip; only strips that pass get the 3x3 test.  At the end of a run
// the survivors (k + a handful) go to a small global table; the CTA that completes a frame's last run (an epoch-stamped
// ticket per frame: no memset, no second launch) merges the frame's rows, ranks them, gathers size/offset/depth
// through the strided views and writes the packed outputs.
//
// Exactness under ties: everything that decides order is done on the final keys (the sigmoid VALUES, like the
// reference).  The cheap filter works on logits with a guard band (reject_key_for_score) so that it never rejects a
// logit whose sigmoid could tie with the k-th score.  If the list overflows (plateaus, no usable threshold) the run
// switches to a safe mode: sub-steps that cannot overflow, exact pruning to the top-k by radix select, and an exact
// 64-bit composite threshold, so that an all-equal map costs time but never correctness.
#pragma once

namespace tauv {

constexpr int kSdCW = 8;                  // consumer warps
constexpr int kSdNC = kSdCW * 32;         // consumer threads (threadIdx.x < kSdNC)
constexpr int kSdThreads = kSdNC + 32;    // + one producer warp (one lane issues the bulk copies)
constexpr int kSdCap = 4096;              // candidate-list capacity (entries)
constexpr int kSdSoft = 2048;             // prune when the list grows beyond this
constexpr int kSdMaxK = 1024;
constexpr int kSdMaxW = 1024;
constexpr int kSdSlice = kSdCap / kSdNC;  // list entries per thread in a compaction
constexpr int kSdMaxStages = 8;

using SdSync = SyncNamed<1, kSdNC>;       // named barrier over the consumer warps; the producer never joins

struct SdArgs {
  const float* hm;
  int B, C, H, W, k;
  int G;                  // CTAs (every one owns rows [i*R/G, (i+1)*R/G))
  long long rows_total;   // R = B*C*H
  int rows_frame;         // C*H
  int chunk_rows, stages; // rows per bulk copy (power of two), ring slots (power of two)
  int tbl_rows, row_cap;  // candidate table: rows per frame, entries per row (2k)
  unsigned long long* cand;    // [B][tbl_rows][row_cap]
  int* cand_count;             // [B][tbl_rows]
  unsigned long long* ticket;  // [B]  epoch << 32 | runs finished
  uint32_t epoch;              // unique per launch (never 0)
  int64_t* out_index;
  int64_t* out_label;
  float* out_score;
  BoxArgs box;
};

struct __align__(16) SdCtx {
  unsigned long long T;     // exact composite threshold (0: none): entries <= T cannot be in the frame's top-k
  unsigned long long keyT;  // scratch: broadcast of a prune threshold
  long long load_row0;      // global row held at stream position 0
  long long frame_row0;     // global row of the current frame's first row
  float thr_f;              // cheap filter in logit (SIGMOID_PEAK) / value (RAW) space; -inf: none
  int count;                // list entries
  int flags;                // bit 1: a push found the list full
  uint32_t maxbin;          // highest occupied window bin
  int safe;                 // 1: overflow happened in this run; the bins are no longer trusted
  int last_scan;            // list length at the last histogram scan
  int found_bin;            // result of the last scan (-1: fewer than k binned)
  int base;                 // scratch for compactions
  int is_last;
  int total;
  int nge;
  int wsum[kSdCW];
  uint32_t sel[8];
};

__device__ __forceinline__ void sd_sync() { SdSync::sync(); }

__device__ __forceinline__ uint64_t sd_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void sd_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void sd_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Shared-memory carve-up (dynamic): ring | list | bins | radix histogram | flags | mbarriers | ctx
struct SdSmem {
  float* ring;
  unsigned long long* list;
  uint32_t* bins;
  uint32_t* hist;
  uint32_t* flags;  // [kSdMaxK]
  uint64_t* full;
  uint64_t* empty;
  SdCtx* ctx;
};
__host__ __device__ inline size_t sd_smem_bytes(int chunk_rows, int stages, int W) {
  return (size_t)stages * chunk_rows * W * 4 + (size_t)kSdCap * 8 + (size_t)kClBins * 4 + (size_t)kRadixBins * 4 +
         (size_t)kSdMaxK * 4 + 2 * kSdMaxStages * 8 + sizeof(SdCtx) + 128;
}
__device__ __forceinline__ SdSmem sd_carve(const SdArgs& a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SdSmem s;
  unsigned char* p = smem_raw;
  s.ring = reinterpret_cast<float*>(p);
  p += (size_t)a.stages * a.chunk_rows * a.W * 4;
  s.list = reinterpret_cast<unsigned long long*>(p);
  p += (size_t)kSdCap * 8;
  s.bins = reinterpret_cast<uint32_t*>(p);
  p += (size_t)kClBins * 4;
  s.hist = reinterpret_cast<uint32_t*>(p);
  p += (size_t)kRadixBins * 4;
  s.flags = reinterpret_cast<uint32_t*>(p);
  p += (size_t)kSdMaxK * 4;
  s.full = reinterpret_cast<uint64_t*>(p);
  s.empty = s.full + kSdMaxStages;
  p += 2 * kSdMaxStages * 8;
  s.ctx = reinterpret_cast<SdCtx*>(p);
  return s;
}

// ---- candidates ----------------------------------------------------------------------------------------------------
// x: logit (SIGMOID_PEAK, already known to be a 3x3 peak) or value (RAW); flat: index inside the frame.
template <int MODE>
__device__ __forceinline__ void sd_push(const SdSmem& sm, float x, uint32_t flat) {
  SdCtx* const ctx = sm.ctx;
  uint32_t key;
  if (MODE == TAUV_TOPK_SIGMOID_PEAK) {
    const float s = sigmoid_ref(x);
    if (!(s > 0.0f)) return;  // underflowed to 0: zero-valued cells are supplied by the filler, like non-peaks
    key = float_to_key(s);
  } else {
    key = float_to_key(x);
  }
  const unsigned long long c = make_composite(key, flat);
  if (c <= ctx->T) return;
  // one shared-memory atomic per warp instruction (same-address atomics serialise)
  const unsigned active = __activemask();
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(active) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(&ctx->count, __popc(active));
  base = __shfl_sync(active, base, leader);
  const int slot = base + __popc(active & ((1u << lane) - 1u));
  if (slot >= kSdCap) {
    ctx->flags = 2;
    return;
  }
  sm.list[slot] = c;
  if (!ctx->safe) {
    if (MODE == TAUV_TOPK_SIGMOID_PEAK && !(x > -80.0f)) return;
    const int bin = cl_window_bin(float_to_key(x));
    if (bin < 0) return;
    atomicAdd(&sm.bins[bin], 1u);
    if ((uint32_t)bin > *reinterpret_cast<volatile uint32_t*>(&ctx->maxbin)) atomicMax(&ctx->maxbin, (uint32_t)bin);
  }
}

// Full test of one 128-bit strip that passed the threshold scan: stream position p (row), column col, values x.
// The rows above and below come from the ring (the halo row of a range is loaded with it); rows outside the plane
// do not exist (-inf padding, decode.py:245-250).
template <int MODE>
__device__ __noinline__ void sd_examine(const SdArgs& a, const SdSmem& sm, int ring_mask, int p, int col, float4 x,
                                        float thr_f) {
  const SdCtx* const ctx = sm.ctx;
  const int W = a.W;
  const int fr = (int)(ctx->load_row0 + p - ctx->frame_row0);  // row inside the frame
  const uint32_t flat = (uint32_t)fr * (uint32_t)W + (uint32_t)col;
  const float xs[4] = {x.x, x.y, x.z, x.w};
  if (MODE != TAUV_TOPK_SIGMOID_PEAK) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
      if (xs[cc] >= thr_f) sd_push<MODE>(sm, xs[cc], flat + cc);
    return;
  }
  const int y = fr % a.H;
  const float NI = TAUV_NEG_INF;
  const float* mid = sm.ring + (size_t)(p & ring_mask) * W + col;
  const bool hl = col > 0, hr = col + 4 < W;
  float4 u = make_float4(NI, NI, NI, NI), d = u;
  float ul = NI, ur = NI, dl = NI, dr = NI;
  if (y > 0) {
    const float* up = sm.ring + (size_t)((p - 1) & ring_mask) * W + col;
    u = *reinterpret_cast<const float4*>(up);
    if (hl) ul = up[-1];
    if (hr) ur = up[4];
  }
  if (y + 1 < a.H) {
    const float* dn = sm.ring + (size_t)((p + 1) & ring_mask) * W + col;
    d = *reinterpret_cast<const float4*>(dn);
    if (hl) dl = dn[-1];
    if (hr) dr = dn[4];
  }
  const float ml = hl ? mid[-1] : NI, mr = hr ? mid[4] : NI;
  float cm[6];  // column-wise max over the three rows, columns col-1 .. col+4
  cm[0] = fmaxf(fmaxf(ul, ml), dl);
  cm[1] = fmaxf(fmaxf(u.x, x.x), d.x);
  cm[2] = fmaxf(fmaxf(u.y, x.y), d.y);
  cm[3] = fmaxf(fmaxf(u.z, x.z), d.z);
  cm[4] = fmaxf(fmaxf(u.w, x.w), d.w);
  cm[5] = fmaxf(fmaxf(ur, mr), dr);
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const float xv = xs[cc];
    if (xv >= thr_f) {
      const float m = fmaxf(fmaxf(cm[cc], cm[cc + 1]), cm[cc + 2]);
      bool peak = (xv >= m);
      // x < m can still tie after the sigmoid (saturation, sub-ulp gap): the reference compares sigmoid values
      if (!peak && (xv > 4.0f || m < -80.0f || (m - xv) < 1e-3f)) peak = sigmoid_tie(xv, m);
      if (peak) sd_push<MODE>(sm, xv, flat + cc);
    }
  }
}

// Threshold scan of stream rows [p0, p1) (one frame): every consumer thread takes 128-bit strips from the ring,
// four in flight; only strips whose maximum reaches the threshold are examined.
template <int MODE>
__device__ __forceinline__ void sd_pass(const SdArgs& a, const SdSmem& sm, int ring_mask, int spr_shift, int p0, int p1) {
  const int tid = threadIdx.x;
  const int W = a.W, spr = W >> 2;
  const int n = (p1 - p0) * spr;
  const float NI = TAUV_NEG_INF;
#pragma unroll 1
  for (int base = 0; base < n; base += 4 * kSdNC) {
    const float thr_f = *reinterpret_cast<volatile float*>(&sm.ctx->thr_f);
    float4 v[4];
    int pr[4], pc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int s = base + u * kSdNC + tid;
      const int r = spr_shift >= 0 ? (s >> spr_shift) : (s / spr);
      pr[u] = p0 + r;
      pc[u] = (s - r * spr) << 2;
      v[u] = (s < n) ? *reinterpret_cast<const float4*>(sm.ring + (size_t)(pr[u] & ring_mask) * W + pc[u])
                     : make_float4(NI, NI, NI, NI);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float m = fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w));
      if (m >= thr_f && base + u * kSdNC + tid < n) sd_examine<MODE>(a, sm, ring_mask, pr[u], pc[u], v[u], thr_f);
    }
  }
}

// ---- threshold maintenance ------------------------------------------------------------------------------------------
// Warp 0: highest window bin b with count(bins >= b) >= k (-1: fewer than k binned).  At most 32 x 32 bins below the
// highest occupied one are visited; counts only grow, so a bin found from a slightly stale view is still valid.
__device__ __forceinline__ int sd_scan_bin(const SdSmem& sm, int k) {
  const int lane = threadIdx.x & 31;
  const int maxbin = (int)*reinterpret_cast<volatile uint32_t*>(&sm.ctx->maxbin);
  uint32_t acc = 0;
  int found = -1;
  for (int it = 0; it < 64 && found < 0; ++it) {
    const int bin = maxbin - it * 32 - lane;
    const uint32_t v = bin >= 0 ? *reinterpret_cast<volatile uint32_t*>(&sm.bins[bin]) : 0u;
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

// Lowest FINAL key that a candidate counted in window bin `bin` or above can have (composite with index bits 0).
template <int MODE>
__device__ __forceinline__ unsigned long long sd_bin_floor(int bin) {
  const float edge = cl_window_edge(bin);
  // SIGMOID_PEAK: the bins count logits, the list holds their sigmoids, whose last-bit wobble the guard band covers
  const float lowest = MODE == TAUV_TOPK_SIGMOID_PEAK ? sigmoid_ref(edge) * (1.0f - 4e-5f) : edge;
  const unsigned long long kt = (unsigned long long)float_to_key(lowest) << 32;
  return kt == 0ull ? 1ull : kt;
}

// Warp 0 (after a pass barrier): rescan the bins when enough new candidates arrived and raise the cheap filter.
template <int MODE>
__device__ __forceinline__ void sd_rescan(const SdArgs& a, const SdSmem& sm, bool force) {
  SdCtx* const ctx = sm.ctx;
  if (ctx->safe) return;
  const int cnt = ctx->count;
  const int every = a.k >= 8 ? a.k / 8 : 1;
  if (cnt < a.k || (!force && cnt - ctx->last_scan < every)) return;
  const int found = sd_scan_bin(sm, a.k);
  if ((threadIdx.x & 31) == 0) {
    ctx->last_scan = cnt;
    ctx->found_bin = found;
    if (found >= 0) {
      const float edge = cl_window_edge(found);  // at least k candidates of this frame have logit/value >= edge
      float t = TAUV_NEG_INF;
      if (MODE == TAUV_TOPK_SIGMOID_PEAK) {
        const uint32_t key = reject_key_for_score(sigmoid_ref(edge));
        if (key) t = key_to_float(key);
      } else {
        t = edge;
      }
      if (t > ctx->thr_f) ctx->thr_f = t;
    }
  }
  __syncwarp();
}

// All consumers: keep the list entries that satisfy `keep` (order not preserved across threads' slices, which is
// fine: the keys carry their own order).  Every thread holds its slice in registers, so the in-place writes cannot
// overtake unread entries.  New length -> ctx->count.
template <class Keep>
__device__ __forceinline__ void sd_compact(const SdSmem& sm, Keep keep) {
  SdCtx* const ctx = sm.ctx;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(ctx->count, kSdCap);
  unsigned long long mine[kSdSlice];
  int nm = 0;
#pragma unroll
  for (int j = 0; j < kSdSlice; ++j) {
    const int i = tid * kSdSlice + j;
    unsigned long long c = 0ull;
    if (i < n) c = sm.list[i];
    const bool kp = (i < n) && keep(c);
    mine[j] = kp ? c : 0ull;
    nm += kp ? 1 : 0;
  }
  int incl = nm;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) ctx->wsum[warp] = incl;
  sd_sync();  // every slice is in registers; warp sums are visible
  int off = incl - nm, tot = 0;
#pragma unroll
  for (int w = 0; w < kSdCW; ++w) {
    if (w < warp) off += ctx->wsum[w];
    tot += ctx->wsum[w];
  }
#pragma unroll
  for (int j = 0; j < kSdSlice; ++j)
    if (mine[j] != 0ull) sm.list[off++] = mine[j];
  sd_sync();
  if (tid == 0) ctx->count = tot;
  sd_sync();
}

// All consumers: exact prune of the list to its top-k (radix select) and, once k entries exist, an exact composite
// threshold (the k-th key: later entries at or below it cannot be in the frame's top-k) plus the matching cheap filter.
template <int MODE>
__device__ __noinline__ void sd_prune_exact(const SdArgs& a, const SdSmem& sm) {
  SdCtx* const ctx = sm.ctx;
  const int n = min(ctx->count, kSdCap);
  const unsigned long long* list = sm.list;
  auto load = [&](int i) { return list[i]; };
  const unsigned long long T = block_kth_largest<kSdNC, decltype(load), SdSync>(load, n, a.k, sm.hist, ctx->sel);
  sd_compact(sm, [&](unsigned long long c) { return c >= T && c != 0ull; });
  if (threadIdx.x == 0 && ctx->count >= a.k && T > ctx->T) {
    ctx->T = T - 1ull;  // (push rejects c <= ctx->T; T itself is in the list already and cannot come again)
    const float v = key_to_float(composite_key(T));
    float t = TAUV_NEG_INF;
    if (MODE == TAUV_TOPK_SIGMOID_PEAK) {
      const uint32_t key = reject_key_for_score(v);
      if (key) t = key_to_float(key);
    } else {
      t = v;
    }
    if (t > ctx->thr_f) ctx->thr_f = t;
  }
  sd_sync();
}

// Descending bitonic sort of 256*E keys held E per thread (element e of thread t is index e*256 + t).  Strides >= 256
// are exchanges inside a thread, strides < 32 shuffles; only strides 32..128 go through shared memory (buf: 256*E
// keys) and the consumer barrier.
template <int E>
__device__ __forceinline__ void sd_sort_desc(unsigned long long (&x)[E], unsigned long long* buf) {
  const int t = threadIdx.x;
  constexpr int N = kSdNC * E;
#pragma unroll 1
  for (int size = 2; size <= N; size <<= 1) {
#pragma unroll 1
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= kSdNC) {
#pragma unroll
        for (int se = E / 2; se >= 1; se >>= 1) {  // stride in elements of one thread (compile-time after unrolling)
          if (stride == se * kSdNC) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
              if ((e & se) == 0) {
                const int i = e * kSdNC + t;
                const bool desc = (i & size) == 0;
                const unsigned long long lo = x[e], hi = x[e | se];
                const bool sw = desc ? (lo < hi) : (lo > hi);
                x[e] = sw ? hi : lo;
                x[e | se] = sw ? lo : hi;
              }
            }
          }
        }
      } else if (stride >= 32) {
#pragma unroll
        for (int e = 0; e < E; ++e) buf[e * kSdNC + t] = x[e];
        sd_sync();
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int i = e * kSdNC + t;
          const unsigned long long y = buf[i ^ stride];
          const bool desc = (i & size) == 0, lower = (i & stride) == 0;
          const bool take_max = lower == desc;
          x[e] = (take_max == (y > x[e])) ? y : x[e];
        }
        sd_sync();
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int i = e * kSdNC + t;
          const unsigned long long y = __shfl_xor_sync(0xffffffffu, x[e], stride);
          const bool desc = (i & size) == 0, lower = (i & stride) == 0;
          const bool take_max = lower == desc;
          x[e] = (take_max == (y > x[e])) ? y : x[e];
        }
      }
    }
  }
}

// One arrival at the frame's ticket; returns how many runs had arrived before.  The word carries the launch's epoch, so
// whatever an earlier launch (or nobody) left in the workspace counts as zero: no memset in front of the kernel.
__device__ __forceinline__ uint32_t sd_ticket_arrive(unsigned long long* w, uint32_t epoch) {
  unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(w);
  while (true) {
    const unsigned long long neu = ((uint32_t)(old >> 32) == epoch) ? old + 1ull : (((unsigned long long)epoch << 32) | 1ull);
    const unsigned long long prev = atomicCAS(w, old, neu);
    if (prev == old) return (uint32_t)neu - 1u;
    old = prev;
  }
}

__device__ __forceinline__ int sd_owner(const SdArgs& a, long long row) {  // CTA whose range holds a global row
  return (int)(((row + 1) * a.G + a.rows_total - 1) / a.rows_total) - 1;
}

// All consumers of the CTA that completed a frame: merge the frame's rows, rank, write the packed outputs.
template <int MODE>
__device__ __noinline__ void sd_merge_emit(const SdArgs& a, const SdSmem& sm, int frame, int n_runs) {
  SdCtx* const ctx = sm.ctx;
  const int tid = threadIdx.x;
  const int k = a.k;
  unsigned long long* pool = sm.list;  // the run's list is in the table already
  const unsigned long long* rows = a.cand + (size_t)frame * a.tbl_rows * a.row_cap;
  const int* cnts = a.cand_count + (size_t)frame * a.tbl_rows;
  if (tid == 0) {
    ctx->total = 0;
    ctx->nge = 0;
  }
  for (int i = tid; i < k; i += kSdNC) sm.flags[i] = 0u;
  sd_sync();
  // (rows were written by other CTAs before their ticket arrival: read them through L2)
  int part = 0;
  for (int r = tid; r < n_runs; r += kSdNC) part += __ldcg(cnts + r);
  if (part) atomicAdd(&ctx->total, part);
  sd_sync();
  const int total = ctx->total;
  int npos;  // entries of the ranked output that are real candidates
  if (total <= kSdCap) {
    if (tid == 0) ctx->base = 0;
    sd_sync();
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < n_runs; r += kSdCW) {
      const int c = __ldcg(cnts + r);
      if (c == 0) continue;
      int base = 0;
      if (lane == 0) base = atomicAdd(&ctx->base, c);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (int i = lane; i < c; i += 32) pool[base + i] = __ldcg(rows + (size_t)r * a.row_cap + i);
    }
    sd_sync();
    if (total > 4 * kSdNC) {  // (many runs per frame: small batches) cut to the exact top-k first
      if (tid == 0) ctx->count = total;
      sd_sync();
      auto load = [&](int i) { return pool[i]; };
      const unsigned long long T = block_kth_largest<kSdNC, decltype(load), SdSync>(load, total, k, sm.hist, ctx->sel);
      sd_compact(sm, [&](unsigned long long c) { return c >= T; });
    }
  } else {
    // does not fit: exact k-th key straight from the table (slot i is valid iff (i % row_cap) < count of its row)
    const int nslots = n_runs * a.row_cap;
    auto load = [&](int i) -> unsigned long long {
      const int r = i / a.row_cap;
      return (i - r * a.row_cap) < __ldcg(cnts + r) ? __ldcg(rows + i) : 0ull;
    };
    const unsigned long long T = block_kth_largest<kSdNC, decltype(load), SdSync>(load, nslots, k, sm.hist, ctx->sel);
    if (tid == 0) ctx->base = 0;
    sd_sync();
    for (int i = tid; i < nslots; i += kSdNC) {
      const unsigned long long c = load(i);
      if (c >= T && c != 0ull) pool[atomicAdd(&ctx->base, 1)] = c;
    }
    sd_sync();
  }
  // (after a cut the pool holds min(k, total) keys)
  int m = total;
  if (total > kSdCap) m = ctx->base;
  else if (total > 4 * kSdNC) m = ctx->count;
  npos = min(m, k);
  sd_sync();
  // sort the pool (<= 1024 keys) descending in registers, ranked keys back to pool[0, m)
  if (m <= 2 * kSdNC) {
    unsigned long long x[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) x[e] = (e * kSdNC + tid < m) ? pool[e * kSdNC + tid] : 0ull;
    sd_sync();
    sd_sort_desc<2>(x, pool);
#pragma unroll
    for (int e = 0; e < 2; ++e) pool[e * kSdNC + tid] = x[e];
  } else {
    unsigned long long x[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) x[e] = (e * kSdNC + tid < m) ? pool[e * kSdNC + tid] : 0ull;
    sd_sync();
    sd_sort_desc<4>(x, pool);
#pragma unroll
    for (int e = 0; e < 4; ++e) pool[e * kSdNC + tid] = x[e];
  }
  sd_sync();
  // ranked outputs
  {
    const BoxArgs& g = a.box;
    const uint32_t hw_elems = (uint32_t)(a.H * a.W);
    int my_ge = 0;
    for (int r = tid; r < npos; r += kSdNC) {
      const unsigned long long c = pool[r];
      const uint32_t flat = composite_idx(c);
      const float sc = key_to_float(composite_key(c));
      const uint32_t lab = flat / hw_elems;
      const uint32_t rem = flat - lab * hw_elems;
      const int iy = (int)(rem / (uint32_t)a.W), ix = (int)(rem - (uint32_t)iy * (uint32_t)a.W);
      const long long slot = (long long)frame * k + r;
      a.out_index[slot * 2 + 0] = iy;
      a.out_index[slot * 2 + 1] = ix;
      a.out_label[slot] = lab;
      a.out_score[slot] = sc;
      if (g.enabled) {
        box_one(g, frame, slot, iy, ix);
        if (!(sc < g.thr)) ++my_ge;
      }
      if (flat < (uint32_t)k) sm.flags[flat] = 1u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_ge += __shfl_xor_sync(0xffffffffu, my_ge, o);
    if ((tid & 31) == 0 && my_ge) atomicAdd(&ctx->nge, my_ge);
  }
  sd_sync();
  if (MODE == TAUV_TOPK_SIGMOID_PEAK && npos < k) {
    // Dense stable top-k semantics: the remaining slots are zero-valued cells in ascending flat index, skipping the
    // selected peaks.  At most npos of the first k cells are selected peaks, so cells [0, k) always suffice.
    const BoxArgs& g = a.box;
    const int lane = tid & 31, warp = tid >> 5;
    const int need = k - npos;
    const long long hw_elems = (long long)a.H * a.W;
    int base = 0;
    for (int start = 0; start < k && base < need; start += kSdNC) {
      const int i = start + tid;
      const bool freec = (i < k) && (sm.flags[i] == 0u);
      const unsigned bal = __ballot_sync(0xffffffffu, freec);
      if (lane == 0) ctx->wsum[warp] = __popc(bal);
      sd_sync();
      int pos = base + __popc(bal & ((1u << lane) - 1u));
      int tot = 0;
      for (int w = 0; w < kSdCW; ++w) {
        if (w < warp) pos += ctx->wsum[w];
        tot += ctx->wsum[w];
      }
      if (freec && pos < need) {
        const int r = npos + pos;
        const long long lab = i / hw_elems;
        const long long rem = i - lab * hw_elems;
        const int iy = (int)(rem / a.W), ix = (int)(rem - (long long)iy * a.W);
        const long long slot = (long long)frame * k + r;
        a.out_index[slot * 2 + 0] = iy;
        a.out_index[slot * 2 + 1] = ix;
        a.out_label[slot] = lab;
        a.out_score[slot] = 0.0f;
        if (g.enabled) box_one(g, frame, slot, iy, ix);
      }
      base += tot;
      sd_sync();
    }
  }
  if (a.box.enabled && tid == 0) {  // entries before the first score < threshold (ranked scores descend; fillers score 0)
    int cnt = ctx->nge;
    if (MODE == TAUV_TOPK_SIGMOID_PEAK && npos < k && !(0.0f < a.box.thr)) cnt += k - npos;
    a.box.count[frame] = cnt;
  }
  sd_sync();
}

// All consumers, end of a run: prune to k + a handful, hand the survivors to the table, arrive at the frame's ticket,
// and — for the run that completes the frame — merge and emit.
template <int MODE>
__device__ __noinline__ void sd_run_end(const SdArgs& a, const SdSmem& sm, int frame) {
  SdCtx* const ctx = sm.ctx;
  const int tid = threadIdx.x;
  if (!ctx->safe) {
    // one last scan of the (complete) histogram: the bin that holds the run's k-th best; drop what lies below its edge
    if (tid < 32) {
      const int bin = sd_scan_bin(sm, a.k);
      if (tid == 0) ctx->keyT = bin >= 0 ? sd_bin_floor<MODE>(bin) : 1ull;  // fewer than k binned: keep every key
    }
    sd_sync();
    const unsigned long long kt = ctx->keyT;
    sd_compact(sm, [&](unsigned long long c) { return c >= kt; });
  }
  if (ctx->safe || ctx->count > a.row_cap) sd_prune_exact<MODE>(a, sm);  // (ties / plateaus / safe mode: exactly <= k)
  // the run's row of the candidate table
  const long long f0 = (long long)frame * a.rows_frame;
  const int first = sd_owner(a, f0), last = sd_owner(a, f0 + a.rows_frame - 1);
  const int n_runs = last - first + 1;
  const int row = (int)blockIdx.x - first;
  const int n = ctx->count;
  unsigned long long* out = a.cand + ((size_t)frame * a.tbl_rows + row) * a.row_cap;
  for (int i = tid; i < n; i += kSdNC) out[i] = sm.list[i];
  if (tid == 0) a.cand_count[(size_t)frame * a.tbl_rows + row] = n;
  __threadfence();
  sd_sync();
  if (tid == 0) {
    const uint32_t before = sd_ticket_arrive(a.ticket + frame, a.epoch);
    __threadfence();
    ctx->is_last = (before == (uint32_t)(n_runs - 1));
  }
  sd_sync();
  if (ctx->is_last) sd_merge_emit<MODE>(a, sm, frame, n_runs);
}

// All consumers: start of a run
__device__ __forceinline__ void sd_run_begin(const SdArgs& a, const SdSmem& sm, long long frame_row0) {
  SdCtx* const ctx = sm.ctx;
  const int tid = threadIdx.x;
  for (int i = tid; i < kClBins / 4; i += kSdNC) reinterpret_cast<uint4*>(sm.bins)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    ctx->T = 0ull;
    ctx->thr_f = TAUV_NEG_INF;
    ctx->count = 0;
    ctx->flags = 0;
    ctx->maxbin = 0u;
    ctx->safe = 0;
    ctx->last_scan = 0;
    ctx->found_bin = -1;
    ctx->frame_row0 = frame_row0;
  }
  sd_sync();
}

// All consumers: stream rows [p0, p1) of the current run through the filter, in passes that end with a barrier and
// the list housekeeping.  While no threshold exists (and always in safe mode) a pass is short enough that it cannot
// overflow the list.
template <int MODE>
__device__ __forceinline__ void sd_rows(const SdArgs& a, const SdSmem& sm, int ring_mask, int spr_shift, int p0, int p1) {
  SdCtx* const ctx = sm.ctx;
  const int tid = threadIdx.x;
  int p = p0;
  while (p < p1) {
    const int count0 = ctx->count;
    const bool bounded = ctx->safe || (ctx->thr_f == TAUV_NEG_INF && ctx->T == 0ull);
    int rows = p1 - p;
    if (bounded) rows = min(rows, max(1, (kSdCap - count0) / a.W));
    sd_sync();  // (everybody has read the pass geometry before anybody pushes)
    sd_pass<MODE>(a, sm, ring_mask, spr_shift, p, p + rows);
    sd_sync();
    if (ctx->flags & 2) {
      // The list overflowed (plateaus, or a threshold that lets too much through).  Drop this pass's pushes, switch the
      // run to safe mode (the bins have counted the dropped pushes: they are not read again), prune exactly, redo.
      sd_sync();
      if (tid == 0) {
        ctx->count = count0;
        ctx->flags = 0;
        ctx->safe = 1;
      }
      sd_sync();
      sd_prune_exact<MODE>(a, sm);
      continue;
    }
    if (tid < 32) sd_rescan<MODE>(a, sm, false);
    sd_sync();
    if (ctx->count > kSdSoft) {
      if (!ctx->safe && ctx->found_bin >= 0) {
        // cheap prune: below the lower edge of the bin that holds the k-th best nothing can matter
        if (tid < 32) sd_rescan<MODE>(a, sm, true);
        sd_sync();
        const unsigned long long kt = sd_bin_floor<MODE>(ctx->found_bin);
        sd_compact(sm, [&](unsigned long long c) { return c >= kt; });
      }
      if (ctx->count > kSdSoft) sd_prune_exact<MODE>(a, sm);
    }
    p += rows;
  }
}

template <int MODE>
__global__ void __launch_bounds__(kSdThreads, 1) stream_decode_kernel(const __grid_constant__ SdArgs a) {
  const SdSmem sm = sd_carve(a);
  SdCtx* const ctx = sm.ctx;
  const int tid = threadIdx.x;
  const int W = a.W, CR = a.chunk_rows, S = a.stages;
  const int ring_mask = CR * S - 1;

  // this CTA's rows and what it loads: one halo row on either side unless the range starts / ends on a plane edge
  const long long own0 = a.rows_total * blockIdx.x / a.G, own1 = a.rows_total * (blockIdx.x + 1) / a.G;
  const int lead = (own0 > 0 && own0 % a.H != 0) ? 1 : 0;
  const int trail = (own1 < a.rows_total && own1 % a.H != 0) ? 1 : 0;
  const long long load_row0 = own0 - lead;
  const int n_own = (int)(own1 - own0);
  const int n_load = n_own + lead + trail;
  const int n_chunks = (n_load + CR - 1) / CR;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_fence_init();
    ctx->load_row0 = load_row0;
  }
  __syncthreads();

  if (tid >= kSdNC) {
    // ---- producer: one lane keeps the ring full; a slot is refilled as soon as the consumers release it
    if (tid == kSdNC) {
      const uint64_t pol = sd_policy_evict_first();
      for (int c = 0; c < n_chunks; ++c) {
        const int slot = c & (S - 1);
        if (c >= S) mbar_wait(&sm.empty[slot], (uint32_t)((c / S - 1) & 1));
        const int rows = min(CR, n_load - c * CR);
        const uint32_t bytes = (uint32_t)rows * (uint32_t)W * 4u;
        mbar_expect_tx(&sm.full[slot], bytes);
        sd_bulk_g2s(sm.ring + (size_t)slot * CR * W, a.hm + (size_t)(load_row0 + (long long)c * CR) * W, bytes,
                    &sm.full[slot], pol);
      }
    }
    return;
  }

  // ---- consumers
  int spr_shift = -1;
  {
    const int spr = W >> 2;
    if ((spr & (spr - 1)) == 0) spr_shift = 31 - __clz(spr);
  }
  if (n_own <= 0) return;
  int done = lead;                  // stream positions [lead, lead + n_own) are this CTA's rows
  const int own_end = lead + n_own;
  int frame = (int)(own0 / a.rows_frame);
  sd_run_begin(a, sm, (long long)frame * a.rows_frame);
#pragma unroll 1
  for (int c = 0; c < n_chunks; ++c) {
    mbar_wait(&sm.full[c & (S - 1)], (uint32_t)((c / S) & 1));
    const int avail = min((c + 1) * CR, n_load);
    // a row can be tested once the row below it has landed (or does not exist)
    const int limit = (avail == n_load) ? own_end : min(own_end, avail - 1);
    while (done < limit) {
      const long long frame_end = (long long)(frame + 1) * a.rows_frame - load_row0;  // stream position of the next frame
      const int seg_end = (int)min((long long)limit, frame_end);
      sd_rows<MODE>(a, sm, ring_mask, spr_shift, done, seg_end);
      done = seg_end;
      if ((long long)done == frame_end || done == own_end) {
        sd_run_end<MODE>(a, sm, frame);
        if (done < own_end) {
          ++frame;
          sd_run_begin(a, sm, (long long)frame * a.rows_frame);
        }
      }
    }
    // rows of chunk c-1 are no longer needed once every row up to the last-but-one of chunk c is done
    sd_sync();
    if (tid == 0 && c >= 1) sd_mbar_arrive(&sm.empty[(c - 1) & (S - 1)]);
  }
}

}  // namespace tauv
