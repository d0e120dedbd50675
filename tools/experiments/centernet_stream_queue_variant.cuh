// centernet_stream.cuh — the CenterNet decode as ONE persistent streaming kernel (round 2).  Included by
// centernet_decode.cu after the shared helpers (keys, sigmoid_tie, reject_key_for_score, window bins, box_one).
//
// Replaces, in one launch (reference file:line under src/tauv_vision/centernet/model/):
//   decode.py:182 sigmoid · :239-252 heatmap_nms (3x3, plateaus survive) · :255-279 heatmap_detect (joint top-k) ·
//   :204-234 per-detection gather + box arithmetic (:87-88 / :65 for decode_keypoints).
//
// Why this shape (measured, profiles/r2_stream_bench_v*.txt): a B200 streams a read-once 335 MB tensor at
// 6.2-6.4 TB/s through a cp.async.bulk (1-D TMA) shared-memory ring with one CTA per SM and ~100 KB per SM in
// flight — as fast as register loads, but the bytes in flight do not depend on what the consuming warps are doing,
// the 3x3 neighbourhood of every cell is already in shared memory (no re-reads: DRAM traffic = algorithmic bytes),
// and an L2 evict-first policy on the copies keeps the previous kernel's dirty lines from stalling the stream.
//
// Work split: the B*C*H rows of the batch are cut into G equal contiguous ranges, one per CTA (G = SMs): every SM
// moves the same number of bytes whatever B is.  A range crosses at most a few frame boundaries; the part of a range
// inside one frame is a RUN.
//
// Roles inside a CTA (18 warps):
//   * producer (1 lane): keeps the ring of 4 x 32 KB chunks full; a slot is refilled when all filter warps released it.
//   * 16 filter warps, INDEPENDENT of each other in steady state (no block barrier, no shared counters with return
//     values): each takes every 16th group of 32 consecutive 128-bit strips of a chunk, compares the strip maximum
//     with the run's rejection threshold (a handful of instructions per strip), and only for strips that pass runs
//     the 3x3 test, evaluates the sigmoid of the peaks and appends their 64-bit composite keys (score key << 32 |
//     ~flat index: plain descending order = score desc, index asc — the order the reference's own KAT asserts,
//     decode.py:327-339) to the warp's PRIVATE sub-list; the peaks' logits are counted in a shared 2048-bin histogram
//     with fire-and-forget shared-memory reductions.  (The first version of this kernel had 8 warps share one list
//     and meet at four block barriers per chunk: every step was a chain of dependent instructions at ~5 cycles each
//     with two warps per scheduler, 4.9 us per 32 KB chunk instead of the 0.73 us the HBM stream allows.)
//   * manager warp: whenever enough new candidates were counted, rescans the histogram for the bin that holds the
//     k-th best candidate of the run so far and raises the rejection threshold (the k-th best candidate so far bounds
//     the frame's k-th best from below) — asynchronously, nobody waits for it.
// At the end of a run the survivors (k + a handful) go to a small global table; the CTA that completes a frame's last
// run (an epoch-stamped ticket per frame: no memset, no second launch) merges the frame's rows, sorts them, gathers
// size/offset/depth through the strided views and writes the packed outputs.
//
// Exactness under ties: everything that decides order is done on the final keys (the sigmoid VALUES, like the
// reference).  The cheap filter works on logits with a guard band (reject_key_for_score) so that it never rejects a
// logit whose sigmoid could tie with the k-th score.  A sub-list that fills up (plateaus, no usable threshold) is
// pruned by its own warp — first below the histogram's floor, then exactly to its top-k by a bitwise search, which
// also gives the warp an exact 64-bit composite threshold — so an all-equal map costs time but never correctness.
#pragma once

namespace tauv {

constexpr int kSdFW = 16;                      // filter warps
constexpr int kSdNF = kSdFW * 32;              // filter threads (threadIdx.x < kSdNF)
constexpr int kSdMgrWarp = kSdFW;              // warp 16: threshold manager
constexpr int kSdLW = kSdFW;                   // warps that own a candidate sub-list
constexpr int kSdQCap = 2048;                  // queue of hot strips waiting for their 3x3 test (entries; a power of two)
constexpr int kSdQHigh = kSdQCap - kSdNF;      // above this many outstanding entries a warp tests its hot strips itself
constexpr int kSdProdWarp = kSdFW + 1;         // warp 17: one lane issues the bulk copies
constexpr int kSdNA = (kSdFW + 1) * 32;        // filter + manager threads (joint barriers)
constexpr int kSdThreads = (kSdFW + 2) * 32;
constexpr int kSdListCap = 8192;               // candidate entries in shared memory, all sub-lists together (at most)
constexpr int kSdSubMax = kSdListCap / kSdFW;  // private sub-list of a list warp: 512 entries (k <= 256) or 256 (k <= 128)
constexpr int kSdMaxK = kSdSubMax / 2;         // a warp can always prune its own sub-list to k and have room again
constexpr int kSdMaxW = 1024;
constexpr int kSdMaxStages = 16;
constexpr int kSdRingPad = 128;
constexpr int kSdU = 4;                        // 128-bit strips per thread and filter iteration
constexpr int kSdPend = 160;                   // per filter warp: peaks waiting for their sigmoid + push (flushed 32 at a time;
                                               // one group of 32 strips can add 128 to the 31 left over)

using SdSyncAll = SyncNamed<1, kSdNA>;         // filter warps + manager
using SdSyncF = SyncNamed<2, kSdNF>;           // filter warps only
#ifdef TAUV_SD_DEBUG
__device__ long long g_sd_stamp[32];
#define SD_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_sd_stamp[i] = t_; } } while (0)
#else
#define SD_STAMP(i) do { } while (0)
#endif
__device__ __forceinline__ void sd_sync_all() { SdSyncAll::sync(); }
__device__ __forceinline__ void sd_sync_f() { SdSyncF::sync(); }

struct SdArgs {
  const float* hm;
  int B, C, H, W, k;
  int G;                  // CTAs (every one owns rows [i*R/G, (i+1)*R/G))
  long long rows_total;   // R = B*C*H
  int rows_frame;         // C*H
  int chunk_rows, stages; // rows per bulk copy, ring slots
  int off_list, off_bins, off_flags, off_pend, off_queue, off_bars, off_ctx;  // byte offsets of the shared-memory regions
  int sub_cap;            // entries of a filter warp's private sub-list (>= 2k; a power of two)
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
  unsigned long long keyT;       // scratch: broadcast of an exact prune threshold
  int found_bin;                 // window bin that holds the k-th best candidate counted so far (-1: fewer than k)
  int pad0;
  unsigned long long warpT[kSdLW]; // exact composite threshold of a list warp (0: none): later entries <= T cannot matter
  long long load_row0;           // global row held at stream position 0
  long long frame_row0;          // global row of the current frame's first row
  uint32_t thr_key;              // cheap filter: order-preserving key of a logit (SIGMOID_PEAK) / value (RAW); 0: none
  uint32_t maxbin;               // highest occupied window bin
  uint32_t pushed;               // candidates counted in the bins so far (this run)
  int req;                       // run-end request: ordinal + 1 of the run the filter warps have finished
  int is_last;
  int total;
  int nge;
  int base;
  int wsum[kSdLW];
  uint32_t sel[8];
  int npend[kSdLW];              // peaks waiting in a list warp's pending buffer
  uint32_t q_tail, q_claim, q_done;  // queue of hot strips: entries reserved / claimed for testing / tested so far
  // where each filter warp stands in the CTA's range (kept here between runs)
  struct { int c, slot, round, done; } st[kSdFW];
};

__device__ __forceinline__ uint64_t sd_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void sd_bulk_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
      "l"(src_gmem), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void sd_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sd_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(100000u)
      : "memory");
}
__device__ __forceinline__ uint32_t sd_lds_u32_volatile(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 sd_lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float sd_lds1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// Shared-memory carve-up (dynamic): ring | sub-lists | bins (= radix histogram once a run's bins are dead) | flags |
// mbarriers | ctx.  The host computes the byte offsets (SdArgs::off_*) so that device code reaches a region with one
// constant-bank load and an add; a struct of pointers handed around by reference would live in local memory.
struct SdLayout {
  int off_list, off_bins, off_flags, off_pend, off_queue, off_bars, off_ctx;
  size_t total;
};
__host__ __device__ inline SdLayout sd_layout(int chunk_rows, int stages, int W, int sub_cap) {
  SdLayout l;
  // every slot holds its chunk plus one halo row on either side; kSdRingPad bytes in front so that the word left of the
  // first slot's first cell is addressable (the 3x3 test loads its neighbourhood unconditionally)
  size_t o = kSdRingPad + (size_t)stages * (chunk_rows + 2) * W * 4;
  l.off_list = (int)o;
  o += (size_t)kSdLW * sub_cap * 8;
  l.off_bins = (int)o;
  o += (size_t)kClBins * 4;
  l.off_flags = (int)o;
  o += (size_t)kSdMaxK * 4;
  l.off_pend = (int)o;
  o += (size_t)kSdLW * kSdPend * 8;
  l.off_queue = (int)o;
  o += (size_t)kSdQCap * 8;
  l.off_bars = (int)o;
  o += 2 * kSdMaxStages * 8;
  l.off_ctx = (int)o;
  o += sizeof(SdCtx);
  l.total = o + 128;
  return l;
}
__host__ __device__ inline size_t sd_smem_bytes(int chunk_rows, int stages, int W, int sub_cap) {
  return sd_layout(chunk_rows, stages, W, sub_cap).total;
}
__device__ __forceinline__ unsigned char* sd_smem() {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  return smem_raw;
}
__device__ __forceinline__ float* sd_ring() { return reinterpret_cast<float*>(sd_smem()); }
__device__ __forceinline__ unsigned long long* sd_list(const SdArgs& a) { return reinterpret_cast<unsigned long long*>(sd_smem() + a.off_list); }
__device__ __forceinline__ uint32_t* sd_bins(const SdArgs& a) { return reinterpret_cast<uint32_t*>(sd_smem() + a.off_bins); }
__device__ __forceinline__ uint32_t* sd_flags(const SdArgs& a) { return reinterpret_cast<uint32_t*>(sd_smem() + a.off_flags); }
__device__ __forceinline__ uint2* sd_pend(const SdArgs& a) { return reinterpret_cast<uint2*>(sd_smem() + a.off_pend); }
__device__ __forceinline__ uint64_t* sd_full(const SdArgs& a) { return reinterpret_cast<uint64_t*>(sd_smem() + a.off_bars); }
__device__ __forceinline__ uint64_t* sd_empty(const SdArgs& a) { return sd_full(a) + kSdMaxStages; }
__device__ __forceinline__ uint2* sd_queue(const SdArgs& a) { return reinterpret_cast<uint2*>(sd_smem() + a.off_queue); }
__device__ __forceinline__ SdCtx* sd_ctx(const SdArgs& a) { return reinterpret_cast<SdCtx*>(sd_smem() + a.off_ctx); }

// Lowest FINAL key that a candidate counted in window bin `bin` or above can have (composite with index bits 0).
template <int MODE>
__device__ __forceinline__ unsigned long long sd_bin_floor(int bin) {
  const float edge = cl_window_edge(bin);
  // SIGMOID_PEAK: the bins count logits, the lists hold their sigmoids, whose last-bit wobble the guard band covers
  const float lowest = MODE == TAUV_TOPK_SIGMOID_PEAK ? sigmoid_ref(edge) * (1.0f - 4e-5f) : edge;
  const unsigned long long kt = (unsigned long long)float_to_key(lowest) << 32;
  return kt == 0ull ? 1ull : kt;
}

// ---- a filter warp's own list housekeeping (warp-convergent) ----------------------------------------------------------
// In-place compaction of the warp's sub-list by a predicate (single warp, in order: writes never pass unread entries).
template <class Keep>
__device__ __forceinline__ int sd_warp_compact(unsigned long long* sub, int n, Keep keep) {
  const int lane = threadIdx.x & 31;
  int out = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    unsigned long long c = 0ull;
    if (i < n) c = sub[i];
    const bool kp = (i < n) && keep(c);
    const unsigned bal = __ballot_sync(0xffffffffu, kp);
    __syncwarp();
    if (kp) sub[out + __popc(bal & ((1u << lane) - 1u))] = c;
    out += __popc(bal);
    __syncwarp();
  }
  return out;
}

// The sub-list is (nearly) full.  First drop what the histogram already rules out; if that does not free half of the
// list, prune exactly to the warp's own top-k (the frame's top-k is a subset of the union of the warps' top-k's) by a
// bitwise search for the k-th largest key, which also becomes this warp's exact push threshold.  Returns the new length.
template <int MODE>
__device__ __noinline__ int sd_warp_prune(const SdArgs& a, int cnt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  SdCtx* const ctx = sd_ctx(a);
  unsigned long long* sub = sd_list(a) + (size_t)warp * a.sub_cap;
  const int fb = *reinterpret_cast<volatile int*>(&ctx->found_bin);
  if (fb >= 0) {  // nothing below the lower edge of the bin that holds the k-th best candidate can matter
    const unsigned long long fl = sd_bin_floor<MODE>(fb);
    cnt = sd_warp_compact(sub, cnt, [&](unsigned long long c) { return c >= fl; });
  }
  if (cnt <= a.sub_cap / 2) return cnt;
  // k-th largest of cnt (> sub_cap/2 >= k) distinct keys, one bit at a time from the top
  unsigned long long T = 0ull;
  for (int bit = 63; bit >= 0; --bit) {
    const unsigned long long cand = T | (1ull << bit);
    int n = 0;
    for (int i = lane; i < cnt; i += 32) n += (sub[i] >= cand) ? 1 : 0;
    n = __reduce_add_sync(0xffffffffu, n);
    if (n >= a.k) T = cand;
  }
  cnt = sd_warp_compact(sub, cnt, [&](unsigned long long c) { return c >= T; });
  if (lane == 0) {
    ctx->warpT[warp] = T;  // (T itself is in the list; keys are distinct, so nothing equal can come again)
    // raise the shared cheap filter from this warp's k-th score
    uint32_t key = 0;
    if (MODE == TAUV_TOPK_SIGMOID_PEAK) key = reject_key_for_score(key_to_float(composite_key(T)));
    else key = composite_key(T);
    if (key) atomicMax(&ctx->thr_key, key);
  }
  __syncwarp();
  return cnt;
}

// Push the first n (<= 32) waiting peaks of this warp, all lanes at once: sigmoid, composite key, a slot in the warp's
// private sub-list; then move the rest of the waiting list (npend - n entries) to its front.  Warp-convergent.
// Returns the sub-list's new length.
template <int MODE>
__device__ __forceinline__ int sd_flush(const SdArgs& a, int n, int npend, int cnt) {
  SdCtx* const ctx = sd_ctx(a);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint2* const pend = sd_pend(a) + warp * kSdPend;
  unsigned long long* const sub = sd_list(a) + (size_t)warp * a.sub_cap;
  bool ok = lane < n;
  unsigned long long c = 0ull;
  if (ok) {
    const uint2 e = pend[lane];
    const float x = __uint_as_float(e.x);
    uint32_t key;
    if (MODE == TAUV_TOPK_SIGMOID_PEAK) {
      const float s = sigmoid_ref(x);
      ok = s > 0.0f;  // underflowed to 0: zero-valued cells are supplied by the filler, like non-peaks
      key = float_to_key(s);
    } else {
      key = float_to_key(x);
    }
    c = make_composite(key, e.y);
    ok = ok && c > ctx->warpT[warp];
  }
  // the rest of the waiting list moves down by n
  for (int i0 = 0; i0 < npend - n; i0 += 32) {  // (warp-uniform trip count)
    const int i = i0 + lane;
    uint2 t = make_uint2(0u, 0u);
    if (i < npend - n) t = pend[i + n];
    __syncwarp();
    if (i < npend - n) pend[i] = t;
    __syncwarp();
  }
  unsigned bal = __ballot_sync(0xffffffffu, ok);
  if (bal == 0u) return cnt;
  if (cnt + __popc(bal) > a.sub_cap) {
    cnt = sd_warp_prune<MODE>(a, cnt);
    ok = ok && c > ctx->warpT[warp];
    bal = __ballot_sync(0xffffffffu, ok);
    if (bal == 0u) return cnt;
  }
  if (ok) sub[cnt + __popc(bal & ((1u << lane) - 1u))] = c;
  return cnt + __popc(bal);
}

// The rare part of a filter iteration, for one group of 32 strips (warp-convergent call; `hot`: this lane's strip
// passed the threshold scan).  Hot lanes run the 3x3 test on their strip — slot row `srow` (1-based: row 0 and the
// last row of a slot are the halo rows loaded with the chunk, so the rows above and below are always in the slot; rows
// outside the plane do not exist: -inf padding, decode.py:245-250) — and append the peaks at or above the threshold
// to the warp's pending buffer, counting them in the shared histogram at once (the threshold must not lag behind what
// waits here).  The expensive part (sigmoid, keys, list) runs later on full warps (sd_flush).  RAW: every cell at or
// above the threshold is a candidate.  Returns the warp's sub-list length.
template <int MODE, bool DENSE = false>
__device__ __forceinline__ int sd_hot(const SdArgs& a, bool hot, uint32_t mid, int col, int fr, float thr_f, int cnt) {
  SdCtx* const ctx = sd_ctx(a);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint2* const pend = sd_pend(a) + warp * kSdPend;
  unsigned dense_mask = 0u;
  float dense_x[4] = {0.f, 0.f, 0.f, 0.f};
  if (hot) {
    const int W = a.W;
    const float4 x = sd_lds4(mid);  // mid: shared-memory address of the strip
    const float xs[4] = {x.x, x.y, x.z, x.w};
    unsigned mask = 0u;
    if (MODE != TAUV_TOPK_SIGMOID_PEAK) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) mask |= (xs[cc] >= thr_f) ? (1u << cc) : 0u;
    } else {
      const uint32_t rowb = (uint32_t)W * 4u;
      const int y = (a.H & (a.H - 1)) == 0 ? (fr & (a.H - 1)) : fr % a.H;
      const float NI = TAUV_NEG_INF;
      const bool hl = col > 0, hr = col + 4 < W;
      float4 u = make_float4(NI, NI, NI, NI), d = u;
      float ul = NI, ur = NI, dl = NI, dr = NI;
      if (y > 0) {
        u = sd_lds4(mid - rowb);
        if (hl) ul = sd_lds1(mid - rowb - 4);
        if (hr) ur = sd_lds1(mid - rowb + 16);
      }
      if (y + 1 < a.H) {
        d = sd_lds4(mid + rowb);
        if (hl) dl = sd_lds1(mid + rowb - 4);
        if (hr) dr = sd_lds1(mid + rowb + 16);
      }
      const float ml = hl ? sd_lds1(mid - 4) : NI, mr = hr ? sd_lds1(mid + 16) : NI;
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
          if (peak) mask |= 1u << cc;
        }
      }
    }
    if (!DENSE && mask) {
      // (one shared-memory atomic per lane that found something: candidates are rare once a threshold exists, and the
      // warp's counter is nobody else's)
      const int n = __popc(mask);
      int slot = atomicAdd(&ctx->npend[warp], n);
      atomicAdd(&ctx->pushed, (uint32_t)n);
      const uint32_t flat = (uint32_t)fr * (uint32_t)W + (uint32_t)col;
      uint32_t* const bins = sd_bins(a);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        if (mask & (1u << cc)) {
          const float xv = xs[cc];
          pend[slot++] = make_uint2(__float_as_uint(xv), flat + cc);
          if (MODE != TAUV_TOPK_SIGMOID_PEAK || xv > -80.0f) {
            const int bin = cl_window_bin(float_to_key(xv));
            if (bin >= 0) {
              atomicAdd(&bins[bin], 1u);
              if ((uint32_t)bin > *reinterpret_cast<volatile uint32_t*>(&ctx->maxbin)) atomicMax(&ctx->maxbin, (uint32_t)bin);
            }
          }
        }
      }
    }
    if (DENSE) dense_mask = mask;
    if (DENSE) {
      dense_x[0] = xs[0]; dense_x[1] = xs[1]; dense_x[2] = xs[2]; dense_x[3] = xs[3];
    }
  }
  if (DENSE) {
    // (most lanes have a strip: slots from a warp prefix sum instead of 32 atomics on one counter)
    const int n = __popc(dense_mask);
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int np0 = *reinterpret_cast<volatile int*>(&ctx->npend[warp]);
    int slot = np0 + incl - n;
    if (dense_mask) {
      const uint32_t flat = (uint32_t)fr * (uint32_t)a.W + (uint32_t)col;
      uint32_t* const bins = sd_bins(a);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        if (dense_mask & (1u << cc)) {
          const float xv = dense_x[cc];
          pend[slot++] = make_uint2(__float_as_uint(xv), flat + cc);
          if (MODE != TAUV_TOPK_SIGMOID_PEAK || xv > -80.0f) {
            const int bin = cl_window_bin(float_to_key(xv));
            if (bin >= 0) {
              atomicAdd(&bins[bin], 1u);
              if ((uint32_t)bin > *reinterpret_cast<volatile uint32_t*>(&ctx->maxbin)) atomicMax(&ctx->maxbin, (uint32_t)bin);
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0 && total) {
      ctx->npend[warp] = np0 + total;
      atomicAdd(&ctx->pushed, (uint32_t)total);
    }
  }
  __syncwarp();
  int np = *reinterpret_cast<volatile int*>(&ctx->npend[warp]);
  while (np >= 32) {
    cnt = sd_flush<MODE>(a, 32, np, cnt);
    np -= 32;
    if (lane == 0) ctx->npend[warp] = np;
    __syncwarp();
  }
  return cnt;
}

// Threshold scan of this warp's share of n consecutive 128-bit strips that start at shared-memory address `base`
// (rows of one frame in one slot; strip 0 = column 0 of slot row srow0 = frame row fr0): the warp takes every kSdFW-th
// group of 32 strips, kSdU groups in flight; while no threshold exists (bootstrap) only one, so that the first threshold
// is used as early as possible.
template <int MODE>
__device__ __forceinline__ int sd_filter(const SdArgs& a, uint32_t slot_u32, int slot_idx, uint32_t empty_u32, uint32_t base, int n,
                                         int srow0, int fr0, int spr_shift, int cnt) {
  const int tid = threadIdx.x;
  const uint32_t thr_addr = smem_u32(sd_smem()) + (uint32_t)a.off_ctx + (uint32_t)offsetof(SdCtx, thr_key);
  const float NI = TAUV_NEG_INF;
  int sb = 0;
#pragma unroll 1
  while (sb < n) {  // (warp-uniform trip count: the body votes)
    const uint32_t tk = sd_lds_u32_volatile(thr_addr);
    const float thr_f = tk ? key_to_float(tk) : NI;
    const int nu = tk ? kSdU : 1;
    const int s0 = sb + tid;
    unsigned hotmask = 0u;
#pragma unroll
    for (int u = 0; u < kSdU; ++u) {
      const int s = s0 + u * kSdNF;
      if (u < nu && s < n) {
        const float4 v = sd_lds4(base + (uint32_t)s * 16u);
        const float m = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        hotmask |= (m >= thr_f) ? (1u << u) : 0u;
      }
    }
    if (tk == 0u) {
      // bootstrap (no threshold yet: every strip is hot, whole warps at a time): test in line
      if (__any_sync(0xffffffffu, hotmask != 0u)) {
        const int spr = a.W >> 2;
#pragma unroll 1
        for (int u = 0; u < kSdU; ++u) {  // (a runtime loop: one copy of the 3x3 test, nothing of the scan live in it)
          const bool h = (hotmask >> u) & 1u;
          if (!__any_sync(0xffffffffu, h)) continue;
          const int s = s0 + u * kSdNF;
          const int r = spr_shift >= 0 ? (s >> spr_shift) : (s / spr);
          const int col = (s - r * spr) << 2;
          cnt = sd_hot<MODE>(a, h, slot_u32 + (uint32_t)(srow0 + r) * (uint32_t)a.W * 4u + (uint32_t)col * 4u, col, fr0 + r, thr_f, cnt);
        }
      }
    } else if (__any_sync(0xffffffffu, hotmask != 0u)) {
      // steady state (a few lanes per iteration): queue the strips.  Whichever warp next reaches the end of its chunk
      // share tests up to 32 queued strips at once (sd_drain) — full warps instead of one or two lanes per streaming
      // warp — and this warp goes on streaming.  The slot cannot go back to the producer before its queued strips are
      // tested: every entry is one pending transaction on the slot's `empty` barrier.  Entry: {shared-memory address
      // of the strip, frame row | column/4 << 20 | slot << 28}.
      SdCtx* const ctx = sd_ctx(a);
      uint2* const queue = sd_queue(a);
      const int spr = a.W >> 2, lane = tid & 31;
      const uint32_t outstanding = *reinterpret_cast<volatile uint32_t*>(&ctx->q_tail) - *reinterpret_cast<volatile uint32_t*>(&ctx->q_done);
      if (outstanding > (uint32_t)kSdQHigh) {
        // (the queue is nearly full — a map without a usable threshold: test in line, like the bootstrap)
#pragma unroll 1
        for (int u = 0; u < kSdU; ++u) {
          const bool h = (hotmask >> u) & 1u;
          if (!__any_sync(0xffffffffu, h)) continue;
          const int s = s0 + u * kSdNF;
          const int r = spr_shift >= 0 ? (s >> spr_shift) : (s / spr);
          const int col = (s - r * spr) << 2;
          cnt = sd_hot<MODE>(a, h, slot_u32 + (uint32_t)(srow0 + r) * (uint32_t)a.W * 4u + (uint32_t)col * 4u, col, fr0 + r, thr_f, cnt);
        }
      } else {
#pragma unroll 1
        for (int u = 0; u < kSdU; ++u) {
          const bool h = (hotmask >> u) & 1u;
          const unsigned bal = __ballot_sync(0xffffffffu, h);
          if (bal == 0u) continue;
          uint32_t base = 0;
          if (lane == 0) {
            asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(empty_u32), "r"((uint32_t)__popc(bal)) : "memory");
            base = atomicAdd(&ctx->q_tail, (uint32_t)__popc(bal));
          }
          base = __shfl_sync(0xffffffffu, base, 0);
          if (h) {
            const int s = s0 + u * kSdNF;
            const int r = spr_shift >= 0 ? (s >> spr_shift) : (s / spr);
            const int c4 = s - r * spr;
            queue[(base + __popc(bal & ((1u << lane) - 1u))) & (kSdQCap - 1)] =
                make_uint2(slot_u32 + (uint32_t)(srow0 + r) * (uint32_t)a.W * 4u + (uint32_t)c4 * 16u,
                           (uint32_t)(fr0 + r) | ((uint32_t)c4 << 20) | ((uint32_t)slot_idx << 28));
          }
        }
      }
    }
    sb += nu * kSdNF;
  }
  return cnt;
}

// Test queued hot strips, up to 32 at a time, until none are left unclaimed (any filter warp, at the end of its share of
// a chunk).  Warp-convergent.  Every tested entry completes one transaction on its slot's `empty` barrier.
template <int MODE>
__device__ __forceinline__ int sd_drain(const SdArgs& a, int cnt) {
  SdCtx* const ctx = sd_ctx(a);
  const int lane = threadIdx.x & 31;
  uint2* const queue = sd_queue(a);
  const uint32_t smem0 = smem_u32(sd_smem());
  const uint32_t thr_addr = smem0 + (uint32_t)a.off_ctx + (uint32_t)offsetof(SdCtx, thr_key);
  const uint32_t empty0 = smem0 + (uint32_t)a.off_bars + (uint32_t)kSdMaxStages * 8u;
#pragma unroll 1
  while (true) {
    uint32_t first = 0, n = 0;
    if (lane == 0) {
      uint32_t claim = *reinterpret_cast<volatile uint32_t*>(&ctx->q_claim);
      while (true) {
        const uint32_t tail = *reinterpret_cast<volatile uint32_t*>(&ctx->q_tail);
        n = min(tail - claim, 32u);
        if (n == 0u) break;
        const uint32_t prev = atomicCAS(&ctx->q_claim, claim, claim + n);
        if (prev == claim) break;
        claim = prev;
      }
      first = claim;
    }
    n = __shfl_sync(0xffffffffu, n, 0);
    if (n == 0u) return cnt;
    first = __shfl_sync(0xffffffffu, first, 0);
    uint2 e = make_uint2(0u, 0u);
    if (lane < (int)n) {
      volatile uint2* q = queue + ((first + lane) & (kSdQCap - 1));
      do {  // (the producer lane is between its reservation and its store)
        e.x = q->x;
      } while (e.x == 0u);
      e.y = q->y;
      q->x = 0u;
    }
    __syncwarp();
    const uint32_t tk = sd_lds_u32_volatile(thr_addr);
    cnt = sd_hot<MODE, true>(a, lane < (int)n, e.x, (int)((e.y >> 20) & 0xffu) << 2, (int)(e.y & 0xfffffu),
                             tk ? key_to_float(tk) : TAUV_NEG_INF, cnt);
    if (lane < (int)n)
      asm volatile("mbarrier.complete_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(empty0 + (e.y >> 28) * 8u), "r"(1u) : "memory");
    if (lane == 0) atomicAdd(&ctx->q_done, n);
    __syncwarp();
  }
}

// A filter warp streams its share of one run: from where it stands (ctx->st) to the end of the frame or of the CTA's
// range, chunk by chunk as they land.  Chunk c holds the CTA's rows [c*CR, (c+1)*CR) in slot rows 1.., with the row
// before in slot row 0 and the row after behind them, so every chunk is tested on its own and its slot is released as
// soon as this warp is through with it.  run_end: position (row of the CTA's range) where the run ends; fr_off: frame
// row of position 0.  Returns the length of the warp's sub-list (the pending peaks are flushed before it returns).
template <int MODE>
__device__ __forceinline__ int sd_stream_run(const SdArgs& a, int n_own, int run_end, int fr_off) {
  SdCtx* const ctx = sd_ctx(a);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int CR = a.chunk_rows, S = a.stages, W = a.W;
  const uint32_t smem0 = smem_u32(sd_smem());
  const uint32_t full_u32 = smem0 + (uint32_t)a.off_bars;
  const uint32_t slot_bytes = (uint32_t)(CR + 2) * (uint32_t)W * 4u;
  const int n_chunks = (n_own + CR - 1) / CR;
  int spr_shift = -1;
  {
    const int spr = W >> 2;
    if ((spr & (spr - 1)) == 0) spr_shift = 31 - __clz(spr);
  }
  int c = ctx->st[warp].c, slot = ctx->st[warp].slot, round = ctx->st[warp].round, done = ctx->st[warp].done;
  int cnt = 0;
#pragma unroll 1
  while (c < n_chunks) {
    sd_mbar_wait(full_u32 + (uint32_t)slot * 8u, (uint32_t)round & 1u);
    const int chunk_end = min((c + 1) * CR, n_own);
    const int seg_end = min(chunk_end, run_end);
    if (done < seg_end) {
      const uint32_t slot_u32 = smem0 + (uint32_t)kSdRingPad + (uint32_t)slot * slot_bytes;
      const int srow0 = done - c * CR + 1;
      cnt = sd_filter<MODE>(a, slot_u32, slot, smem_u32(sd_empty(a) + slot), slot_u32 + (uint32_t)srow0 * (uint32_t)W * 4u,
                            (seg_end - done) * (W >> 2), srow0, fr_off + done, spr_shift, cnt);
      done = seg_end;
    }
    if (done < chunk_end) break;  // the run ends in the middle of this chunk: the next run carries on from here
    __syncwarp();
    if (lane == 0) sd_mbar_arrive(sd_empty(a) + slot);  // this warp is through with the slot
    cnt = sd_drain<MODE>(a, cnt);                        // ... and tests what is queued
    ++c;
    if (++slot == S) {
      slot = 0;
      ++round;
    }
    if (done == run_end) break;
  }
  if (lane == 0) {
    ctx->st[warp].c = c;
    ctx->st[warp].slot = slot;
    ctx->st[warp].round = round;
    ctx->st[warp].done = done;
  }
  cnt = sd_drain<MODE>(a, cnt);  // (the last warp to get here finds everything the run queued)
  const int np = *reinterpret_cast<volatile int*>(&ctx->npend[warp]);
  if (np > 0) cnt = sd_flush<MODE>(a, np, np, cnt);
  __syncwarp();
  if (lane == 0) ctx->npend[warp] = 0;
  __syncwarp();
  return cnt;
}

// ---- manager warp ---------------------------------------------------------------------------------------------------
// Highest window bin b with count(bins >= b) >= k (-1: fewer than k binned).  Counts only grow, so a bin found from a
// slightly stale view is still valid.
__device__ __forceinline__ int sd_scan_bin(const SdArgs& a) {
  const int lane = threadIdx.x & 31;
  const uint32_t* bins = sd_bins(a);
  const int maxbin = (int)*reinterpret_cast<volatile uint32_t*>(&sd_ctx(a)->maxbin);
  uint32_t acc = 0;
  int found = -1;
  for (int it = 0; it < 64 && found < 0; ++it) {
    const int bin = maxbin - it * 32 - lane;
    const uint32_t v = bin >= 0 ? *reinterpret_cast<const volatile uint32_t*>(&bins[bin]) : 0u;
    uint32_t pre = v;  // inclusive prefix over lanes (lane 0 = highest bin)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, acc + pre >= (uint32_t)a.k);
    if (hit) found = maxbin - it * 32 - (__ffs(hit) - 1);
    acc += __shfl_sync(0xffffffffu, pre, 31);
    if (maxbin - (it + 1) * 32 < 0) break;
  }
  return found;
}

// Key of a logit (value) x_c below which nothing can reach the k-th best score, given that at least k candidates have
// logits (values) >= edge.  SIGMOID_PEAK: every x < x_c must have sigmoid(x) strictly below sigmoid(edge) by a relative
// 2e-5 (the guard band reject_key_for_score keeps against the last-bit wobble of expf and against ties after the
// sigmoid); since d sigmoid / sigmoid = (1 - s) dx, a margin of 5e-5 (1 + e^x) does it, and 1e-3 max(1, |x|) as well
// for x <= 0.  Far out (saturation, denormal scores) the exact routine decides.
template <int MODE>
__device__ __forceinline__ uint32_t sd_reject_key(float edge) {
  if (MODE != TAUV_TOPK_SIGMOID_PEAK) return float_to_key(edge);
  if (!(edge < 12.0f) || !(edge > -60.0f)) return reject_key_for_score(sigmoid_ref(edge));
  const float margin = fmaxf(1e-3f * fmaxf(1.0f, fabsf(edge)), 5e-5f * (1.0f + __expf(edge) * 1.01f));
  return float_to_key(edge - margin);
}

// Rescan the bins and raise the cheap filter; remember the bin (the prune floor is derived from it when needed).
template <int MODE>
__device__ __forceinline__ void sd_rescan(const SdArgs& a) {
  SdCtx* const ctx = sd_ctx(a);
  const int found = sd_scan_bin(a);
  if ((threadIdx.x & 31) == 0 && found > ctx->found_bin) {
    const uint32_t key = sd_reject_key<MODE>(cl_window_edge(found));  // at least k candidates have logit/value >= edge
    if (key) atomicMax(&ctx->thr_key, key);
    *reinterpret_cast<volatile int*>(&ctx->found_bin) = found;
  }
  __syncwarp();
}

// Descending bitonic sort of NT*E keys held E per thread (element e of thread t is index e*NT + t).  Strides >= NT
// are exchanges inside a thread, strides < 32 shuffles; only strides 32..NT/2 go through shared memory (buf: NT*E keys)
// and the group's barrier.
template <int NT, int E, class Sync>
__device__ __forceinline__ void sd_sort_desc(unsigned long long (&x)[E], unsigned long long* buf) {
  const int t = threadIdx.x;
  constexpr int N = NT * E;
#pragma unroll 1
  for (int size = 2; size <= N; size <<= 1) {
#pragma unroll 1
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= NT) {
#pragma unroll
        for (int se = E / 2; se >= 1; se >>= 1) {  // stride in elements of one thread (compile-time after unrolling)
          if (stride == se * NT) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
              if ((e & se) == 0) {
                const int i = e * NT + t;
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
        for (int e = 0; e < E; ++e) buf[e * NT + t] = x[e];
        Sync::sync();
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int i = e * NT + t;
          const unsigned long long y = buf[i ^ stride];
          const bool desc = (i & size) == 0, lower = (i & stride) == 0;
          const bool take_max = lower == desc;
          x[e] = (take_max == (y > x[e])) ? y : x[e];
        }
        Sync::sync();
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int i = e * NT + t;
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
// `guess`: the word as read earlier (saves the first round trip when nobody arrived in between).
__device__ __forceinline__ uint32_t sd_ticket_arrive(unsigned long long* w, uint32_t epoch, unsigned long long guess) {
  unsigned long long old = guess;
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

// Filter warps of the CTA that completed a frame: merge the frame's rows, sort, write the packed outputs.
template <int MODE>
__device__ __forceinline__ void sd_merge_emit(const SdArgs& a, int frame, int n_runs) {
  SdCtx* const ctx = sd_ctx(a);
  const int tid = threadIdx.x;
  const int k = a.k;
  unsigned long long* pool = sd_list(a);  // the run's candidates are in the table already
  uint32_t* flags = sd_flags(a);
  uint32_t* hist = sd_bins(a);            // (the run's bins are dead)
  const unsigned long long* rows = a.cand + (size_t)frame * a.tbl_rows * a.row_cap;
  const int* cnts = a.cand_count + (size_t)frame * a.tbl_rows;
  const int nslots = n_runs * a.row_cap;
  const int pool_cap = kSdLW * a.sub_cap;
  int m;  // keys in the pool
  if (nslots <= pool_cap) {
    // all rows at once, whatever their counts (one round trip through L2: the rows were written by other CTAs before
    // their ticket arrival), invalid slots as 0; then squeeze the zeros out while sorting
    for (int i = tid; i < nslots; i += kSdNF) {
      const int r = i / a.row_cap;
      const unsigned long long c = __ldcg(rows + (size_t)r * a.row_cap + (i - r * a.row_cap));
      pool[i] = (i - r * a.row_cap) < __ldcg(cnts + r) ? c : 0ull;
    }
    for (int i = tid; i < k; i += kSdNF) flags[i] = 0u;
    if (tid == 0) {
      ctx->nge = 0;
      ctx->base = 0;
    }
    sd_sync_f();
    SD_STAMP(5);
    m = nslots;
    if (nslots > 2 * kSdNF) {  // (many runs per frame: small batches) cut to the exact top-k first
      auto load = [&](int i) { return pool[i]; };
      const unsigned long long T = block_kth_largest<kSdNF, decltype(load), SdSyncF>(load, nslots, k, hist, ctx->sel);
      // survivors to the front: every thread holds its slice in registers, so in-place writes cannot pass unread entries
      constexpr int SL = kSdListCap / kSdNF;  // (pool_cap <= kSdListCap)
      unsigned long long mine[SL];
      int nm = 0;
#pragma unroll
      for (int j = 0; j < SL; ++j) {
        const int i = tid * SL + j;
        const unsigned long long c = i < nslots ? pool[i] : 0ull;
        const bool kp = c >= T && c != 0ull;
        mine[j] = kp ? c : 0ull;
        nm += kp ? 1 : 0;
      }
      sd_sync_f();
      int off = 0;
      if (nm) off = atomicAdd(&ctx->base, nm);
#pragma unroll
      for (int j = 0; j < SL; ++j)
        if (mine[j] != 0ull) pool[off++] = mine[j];
      sd_sync_f();
      m = ctx->base;
    }
  } else {
    // does not fit: exact k-th key straight from the table (slot i is valid iff (i % row_cap) < count of its row)
    for (int i = tid; i < k; i += kSdNF) flags[i] = 0u;
    if (tid == 0) {
      ctx->nge = 0;
      ctx->base = 0;
    }
    sd_sync_f();
    auto load = [&](int i) -> unsigned long long {
      const int r = i / a.row_cap;
      return (i - r * a.row_cap) < __ldcg(cnts + r) ? __ldcg(rows + i) : 0ull;
    };
    const unsigned long long T = block_kth_largest<kSdNF, decltype(load), SdSyncF>(load, nslots, k, hist, ctx->sel);
    for (int i = tid; i < nslots; i += kSdNF) {
      const unsigned long long c = load(i);
      if (c >= T && c != 0ull) pool[atomicAdd(&ctx->base, 1)] = c;
    }
    sd_sync_f();
    m = ctx->base;
  }
  // sort the pool (<= 1024 slots) descending in registers; zeros (invalid slots) sink to the end
  int npos;
  {
    unsigned long long x[2];
    if (m <= kSdNF) {
      unsigned long long y[1];
      y[0] = tid < m ? pool[tid] : 0ull;
      sd_sync_f();
      sd_sort_desc<kSdNF, 1, SdSyncF>(y, pool);
      x[0] = y[0];
      x[1] = 0ull;
    } else {
#pragma unroll
      for (int e = 0; e < 2; ++e) x[e] = (e * kSdNF + tid < m) ? pool[e * kSdNF + tid] : 0ull;
      sd_sync_f();
      sd_sort_desc<kSdNF, 2, SdSyncF>(x, pool);
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) pool[e * kSdNF + tid] = x[e];
  }
  sd_sync_f();
  SD_STAMP(6);
  {
    // npos = number of non-zero keys among the first k of the sorted pool (they are a prefix)
    int lo = 0, hi = min(k, 2 * kSdNF);
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (pool[mid] != 0ull) lo = mid + 1; else hi = mid;
    }
    npos = lo;
  }
  // ranked outputs
  {
    const BoxArgs& g = a.box;
    const uint32_t hw_elems = (uint32_t)(a.H * a.W);
    int my_ge = 0;
    for (int r = tid; r < npos; r += kSdNF) {
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
      if (flat < (uint32_t)k) flags[flat] = 1u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_ge += __shfl_xor_sync(0xffffffffu, my_ge, o);
    if ((tid & 31) == 0 && my_ge) atomicAdd(&ctx->nge, my_ge);
  }
  SD_STAMP(7);
  sd_sync_f();
  SD_STAMP(8);
  if (MODE == TAUV_TOPK_SIGMOID_PEAK && npos < k) {
    // Dense stable top-k semantics: the remaining slots are zero-valued cells in ascending flat index, skipping the
    // selected peaks.  At most npos of the first k cells are selected peaks, so cells [0, k) always suffice (k <= 256
    // <= the filter threads: one round).
    const BoxArgs& g = a.box;
    const int lane = tid & 31, warp = tid >> 5;
    const int need = k - npos;
    const long long hw_elems = (long long)a.H * a.W;
    const int i = tid;
    const bool freec = (i < k) && (flags[i] == 0u);
    const unsigned bal = __ballot_sync(0xffffffffu, freec);
    if (lane == 0) ctx->wsum[warp] = __popc(bal);
    sd_sync_f();
    int pos = __popc(bal & ((1u << lane) - 1u));
    for (int ww = 0; ww < warp; ++ww) pos += ctx->wsum[ww];
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
    sd_sync_f();
  }
  if (a.box.enabled && tid == 0) {  // entries before the first score < threshold (ranked scores descend; fillers score 0)
    int cnt = ctx->nge;
    if (MODE == TAUV_TOPK_SIGMOID_PEAK && npos < k && !(0.0f < a.box.thr)) cnt += k - npos;
    a.box.count[frame] = cnt;
  }
  sd_sync_f();
}

// Filter warps, end of a run: prune every sub-list to what can still matter (the latest
// floor of the histogram), hand the survivors to the table, arrive at the frame's ticket, and — for the run that
// completes the frame — merge and emit (filter warps).
template <int MODE>
__device__ __forceinline__ void sd_run_end(const SdArgs& a, int frame, int cnt) {
  SdCtx* const ctx = sd_ctx(a);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long* sub = sd_list(a) + (size_t)warp * a.sub_cap;
  // floor of the bin that holds the run's k-th best (fewer than k binned: keep everything)
  unsigned long long kt = ctx->found_bin >= 0 ? sd_bin_floor<MODE>(ctx->found_bin) : 1ull;
  auto count_ge = [&](unsigned long long t) {
    int n = 0;
    for (int i = lane; i < cnt; i += 32) n += (sub[i] >= t) ? 1 : 0;
    return __reduce_add_sync(0xffffffffu, n);
  };
  SD_STAMP(0);
  int mine = count_ge(kt);
  if (lane == 0) ctx->wsum[warp] = mine;
  sd_sync_f();
  SD_STAMP(1);
  int total = 0;
  for (int ww = 0; ww < kSdLW; ++ww) total += ctx->wsum[ww];
  if (total > a.row_cap) {
    // (ties / plateaus / a stale floor) exact k-th key over the union of the sub-lists (a padded [warp][sub_cap] table)
    sd_sync_f();
    if (lane == 0) ctx->wsum[warp] = cnt;
    sd_sync_f();
    if (warp < kSdFW) {
      const unsigned long long* list = sd_list(a);
      const int* wc = ctx->wsum;
      const int sc = a.sub_cap;
      auto load = [&](int i) -> unsigned long long { return (i & (sc - 1)) < wc[i / sc] ? list[i] : 0ull; };
      const unsigned long long T = block_kth_largest<kSdNF, decltype(load), SdSyncF>(load, kSdLW * sc, a.k, sd_bins(a), ctx->sel);
      if (tid == 0) ctx->keyT = T;
    }
    sd_sync_f();
    const unsigned long long T = ctx->keyT;
    kt = T > kt ? T : kt;
    mine = count_ge(kt);
    sd_sync_f();
    if (lane == 0) ctx->wsum[warp] = mine;
    sd_sync_f();
    total = 0;
    for (int ww = 0; ww < kSdLW; ++ww) total += ctx->wsum[ww];
  }
  int off = 0;
  for (int ww = 0; ww < warp; ++ww) off += ctx->wsum[ww];
  // the run's row of the candidate table
  const long long f0 = (long long)frame * a.rows_frame;
  const int first = sd_owner(a, f0), last = sd_owner(a, f0 + a.rows_frame - 1);
  const int n_runs = last - first + 1;
  const int row = (int)blockIdx.x - first;
  unsigned long long* out = a.cand + ((size_t)frame * a.tbl_rows + row) * a.row_cap;
  for (int i0 = 0; i0 < cnt; i0 += 32) {
    const int i = i0 + lane;
    unsigned long long c = 0ull;
    if (i < cnt) c = sub[i];
    const bool kp = (i < cnt) && c >= kt;
    const unsigned bal = __ballot_sync(0xffffffffu, kp);
    if (kp) out[off + __popc(bal & ((1u << lane) - 1u))] = c;
    off += __popc(bal);
  }
  SD_STAMP(2);
  if (tid == 0) a.cand_count[(size_t)frame * a.tbl_rows + row] = total;
  __threadfence();
  sd_sync_f();
  SD_STAMP(3);
  if (tid == 0) {
    const uint32_t before =
        sd_ticket_arrive(a.ticket + frame, a.epoch, *reinterpret_cast<volatile unsigned long long*>(a.ticket + frame));
    __threadfence();
    ctx->is_last = (before == (uint32_t)(n_runs - 1));
  }
  sd_sync_f();
  SD_STAMP(4);
  if (ctx->is_last) {
    if (warp < kSdFW) sd_merge_emit<MODE>(a, frame, n_runs);
    sd_sync_f();
  }
}

// Filter warps: the shared state of the next run (the manager is parked at the run-begin barrier)
__device__ __forceinline__ void sd_run_reset(const SdArgs& a, int frame) {
  SdCtx* const ctx = sd_ctx(a);
  const int tid = threadIdx.x;
  uint32_t* bins = sd_bins(a);
  for (int i = tid; i < kClBins / 4; i += kSdNF) reinterpret_cast<uint4*>(bins)[i] = make_uint4(0, 0, 0, 0);
  if (tid < kSdLW) ctx->warpT[tid] = 0ull;
  if (tid == 0) {
    ctx->found_bin = -1;
    ctx->thr_key = 0u;
    ctx->maxbin = 0u;
    ctx->pushed = 0u;
    ctx->frame_row0 = (long long)frame * a.rows_frame;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ticket + frame));  // (the run's arrival will want the word near)
  }
}

template <int MODE>
__global__ void __launch_bounds__(kSdThreads, 1) stream_decode_kernel(const __grid_constant__ SdArgs a) {
  SdCtx* const ctx = sd_ctx(a);
  uint64_t* const full = sd_full(a);
  uint64_t* const empty = sd_empty(a);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int W = a.W, CR = a.chunk_rows, S = a.stages;

  // this CTA's rows
  const long long own0 = a.rows_total * blockIdx.x / a.G, own1 = a.rows_total * (blockIdx.x + 1) / a.G;
  const int n_own = (int)(own1 - own0);
  const int n_chunks = (n_own + CR - 1) / CR;
  const int frame0 = (int)(own0 / a.rows_frame);
  const int n_runs_cta = n_own > 0 ? (int)((own1 - 1) / a.rows_frame) - frame0 + 1 : 0;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kSdFW);  // every filter warp is through with the slot (+ one transaction per queued strip)
    }
    mbar_fence_init();
    ctx->req = 0;
    ctx->q_tail = 0u;
    ctx->q_claim = 0u;
    ctx->q_done = 0u;
  }
  for (int i = tid; i < kSdQCap; i += kSdThreads) sd_queue(a)[i] = make_uint2(0u, 0u);
  if (tid < kSdFW) {
    ctx->npend[tid] = 0;
    ctx->st[tid].c = 0;
    ctx->st[tid].slot = 0;
    ctx->st[tid].round = 0;
    ctx->st[tid].done = 0;
  }
  __syncthreads();
  if (n_own <= 0) return;

  if (warp == kSdProdWarp) {
    // ---- producer: one lane keeps the ring full; a slot is refilled as soon as all filter warps released it.  A chunk
    // is loaded with the row before and the row after it (where the tensor has them), so neighbouring chunks overlap by
    // two rows: the second read of a row comes from L2, and no slot ever depends on another
    if (lane == 0) {
      const uint64_t pol = sd_policy_evict_first();
      const uint32_t ring_u32 = smem_u32(sd_ring());
      const uint32_t rowb = (uint32_t)W * 4u;
      int slot = 0;
      uint32_t round = 0;
      for (int c = 0; c < n_chunks; ++c) {
        if (round > 0) mbar_wait(&empty[slot], (round - 1) & 1);
        const long long g0 = own0 + (long long)c * CR;                            // first row of the chunk
        const long long g1 = own0 + min((long long)(c + 1) * CR, (long long)n_own);  // one past its last row
        const long long l0 = g0 > 0 ? g0 - 1 : g0, l1 = g1 < a.rows_total ? g1 + 1 : g1;
        const uint32_t bytes = (uint32_t)(l1 - l0) * rowb;
        mbar_expect_tx(&full[slot], bytes);
        sd_bulk_g2s(ring_u32 + (uint32_t)kSdRingPad + (uint32_t)slot * (uint32_t)(CR + 2) * rowb + (uint32_t)(l0 - (g0 - 1)) * rowb,
                    a.hm + (size_t)l0 * W, bytes, smem_u32(&full[slot]), pol);
        if (++slot == S) {
          slot = 0;
          ++round;
        }
      }
    }
    return;
  }

  if (warp == kSdMgrWarp) {
    // ---- manager: while the filter warps stream a run, rescan the histogram whenever enough new candidates were
    // counted and raise the rejection threshold.  Nobody waits for it.
    const int every = a.k >= 32 ? a.k / 16 : 2;
    for (int run = 0; run < n_runs_cta; ++run) {
      sd_sync_all();  // run begin (the filter warps have reset the shared state)
      uint32_t last_scan = 0;
      while (*reinterpret_cast<volatile int*>(&ctx->req) != run + 1) {
        const uint32_t pushed = *reinterpret_cast<volatile uint32_t*>(&ctx->pushed);
        if (pushed >= (uint32_t)a.k && (last_scan == 0 || pushed - last_scan >= (uint32_t)every)) {
          sd_rescan<MODE>(a);
          last_scan = pushed;
        } else if (last_scan != 0) {
          __nanosleep(64);
        }
      }
      sd_rescan<MODE>(a);  // (whatever arrived since the last scan)
      sd_sync_all();        // every filter warp has finished the run's rows; the floor is final
    }
    return;
  }

  // ---- filter warps: per run — reset the shared state, stream, hand over
#pragma unroll 1
  for (int run = 0; run < n_runs_cta; ++run) {
    const int frame = frame0 + run;
    sd_run_reset(a, frame);
    sd_sync_all();  // run begin
    const long long f0 = (long long)frame * a.rows_frame;
    const int run_end = (int)min((long long)n_own, f0 + a.rows_frame - own0);  // position where the run ends
#ifdef TAUV_SD_DEBUG
    long long tq0, tq1, tq2, tq3;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq0));
#endif
    const int cnt = sd_stream_run<MODE>(a, n_own, run_end, (int)(own0 - f0));
#ifdef TAUV_SD_DEBUG
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq1));
#endif
    if (lane == 0) *reinterpret_cast<volatile int*>(&ctx->req) = run + 1;
    sd_sync_all();  // all rows of the run are filtered and tested, and the manager's last scan is published
#ifdef TAUV_SD_DEBUG
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq2));
#endif
    sd_run_end<MODE>(a, frame, cnt);
#ifdef TAUV_SD_DEBUG
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq3));
    if (tid == 0 && (blockIdx.x % 21) == 0)
      printf("[sd] cta %d run %d/%d: stream %lld ns, barrier %lld ns, run_end %lld ns (is_last %d)\n", blockIdx.x, run, n_runs_cta,
             tq1 - tq0, tq2 - tq1, tq3 - tq2, ctx->is_last);
    if (tid == 0 && blockIdx.x == 0)
      printf("[sd] cta 0 stamps: count %lld write %lld fence+sync %lld ticket %lld | load %lld sort %lld emit %lld sync %lld\n",
             g_sd_stamp[1] - g_sd_stamp[0], g_sd_stamp[2] - g_sd_stamp[1], g_sd_stamp[3] - g_sd_stamp[2], g_sd_stamp[4] - g_sd_stamp[3],
             g_sd_stamp[5] - g_sd_stamp[4], g_sd_stamp[6] - g_sd_stamp[5], g_sd_stamp[7] - g_sd_stamp[6], g_sd_stamp[8] - g_sd_stamp[7]);
#endif
  }
}

}  // namespace tauv
