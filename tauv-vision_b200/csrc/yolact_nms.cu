// yolact_nms.cu — YOLACT confidence, top-k and Fast NMS on sm_100a.
//
// Replaces (reference file:line under src/tauv_vision/yolact/model/nms.py):
//   :9-10   softmax(classification)[..., 1:].max(-1)       (full [B,N,C1] softmax materialised today)
//   :12-17  full-N descending sort, first top_k of frame 0
//   :19-22  iou_matrix of the top_k boxes, triu(1), column max
//   :24-27  keep = (iou_max <= thr) & (conf >= thr); idx[keep]
// and the decode of only the top_k priors (boxes.py:55-61) for the fused detect entry point.
//
// K1 scores_kernel   : one warp per class row, lanes across classes, coalesced 4-byte loads straight from
//                      HBM (rows are 324 B, never 16-B aligned), shuffle max / sum.  Reads cls exactly once.
// K2 nms_frame_kernel: one CTA per frame: radix-select the top_k (confidence desc, prior asc), bitonic
//                      sort, boxes to shared memory, 32x32 IoU tiles -> per-column suppression bitmask
//                      (__ballot + atomicOr), ordered compaction of the survivors.
#include "common.cuh"
#include "yolact_common.cuh"

namespace tauv {

constexpr int kNmsThreads = 1024;
constexpr int kNmsMaxTopK = 4096;
constexpr int kNmsRegs = 20;    // scores per thread held in registers by the one-pass ranking (N <= 20480)
constexpr int kNmsCand = 1024;  // its candidate list (aliases the 8 KB radix histogram)

// ---- K1 ----------------------------------------------------------------------------------------
// NCH = ceil(C1/32) register-resident chunks per lane (C1 <= 128); NCH = 0 -> generic two-pass loop.
template <int NCH, bool ARGMAX>
__global__ void __launch_bounds__(256) scores_kernel(const float* __restrict__ cls, long long rows, int C1,
                                                     float* __restrict__ score, int32_t* __restrict__ argmax_all) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  // a warp owns blocks of 32 consecutive rows so that the results leave as one coalesced store
  for (long long blk = warp; blk * 32 < rows; blk += nwarps) {
    const long long row0 = blk * 32;
    const int nrows = (int)min(32LL, rows - row0);
    float my_score = 0.f;
    int my_arg = 0;
    for (int r = 0; r < nrows; ++r) {
      const float* x = cls + (row0 + r) * C1;
      float m = TAUV_NEG_INF, mfg = TAUV_NEG_INF;
      float best_v = TAUV_NEG_INF;
      int best_i = 0x7fffffff;
      float sum = 0.f;
      if (NCH > 0) {
        float v[NCH > 0 ? NCH : 1];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int j = lane + 32 * c;
          v[c] = (j < C1) ? __ldg(x + j) : TAUV_NEG_INF;
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int j = lane + 32 * c;
          m = fmaxf(m, v[c]);
          if (j >= 1) mfg = fmaxf(mfg, v[c]);
          if (ARGMAX && j < C1 && v[c] > best_v) { best_v = v[c]; best_i = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          mfg = fmaxf(mfg, __shfl_xor_sync(0xffffffffu, mfg, o));
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int j = lane + 32 * c;
          if (j < C1) sum = __fadd_rn(sum, expf(__fsub_rn(v[c], m)));
        }
      } else {
        for (int j = lane; j < C1; j += 32) {
          const float t = __ldg(x + j);
          m = fmaxf(m, t);
          if (j >= 1) mfg = fmaxf(mfg, t);
          if (ARGMAX && t > best_v) { best_v = t; best_i = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          mfg = fmaxf(mfg, __shfl_xor_sync(0xffffffffu, mfg, o));
        }
        for (int j = lane; j < C1; j += 32) sum = __fadd_rn(sum, expf(__fsub_rn(__ldg(x + j), m)));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
      // softmax is monotone per row: max_j>=1 softmax_j = exp(max_fg - max) / sum
      const float s = __fdiv_rn(expf(__fsub_rn(mfg, m)), sum);
      if (ARGMAX) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
          const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
          if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
        }
      }
      if (lane == r) {
        my_score = s;
        my_arg = best_i;
      }
    }
    if (lane < nrows) {
      score[row0 + lane] = my_score;
      if (ARGMAX) argmax_all[row0 + lane] = my_arg;
    }
  }
}

// K1' scores_tile_kernel: the same result from a different mapping (C1 <= 256).  A warp pulls 32 consecutive class rows
// (one contiguous 32*C1*4-byte block) into its own shared-memory tile with 4-byte cp.async (128 contiguous bytes per
// warp instruction, no alignment requirement; rows of 81 floats are never 16-byte aligned), then LANE = ROW: every lane
// walks its own row at an odd word stride (conflict-free), so the max / sum-of-exp reductions are plain sequential
// loops with no shuffles — a third of the instructions of the lanes-across-classes version, which was issue-bound.
constexpr int kScoreTileWarps = 4;

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool ARGMAX>
__global__ void __launch_bounds__(kScoreTileWarps * 32) scores_tile_kernel(const float* __restrict__ cls, long long rows,
                                                                            int C1, int stride,
                                                                            float* __restrict__ score,
                                                                            int32_t* __restrict__ argmax_all) {
  extern __shared__ __align__(16) float s_tiles[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* tile = s_tiles + (size_t)wib * 32 * stride;
  const long long warp = (long long)blockIdx.x * kScoreTileWarps + wib;
  const long long nwarps = (long long)gridDim.x * kScoreTileWarps;
  for (long long blk = warp; blk * 32 < rows; blk += nwarps) {
    const long long row0 = blk * 32;
    const int nrows = (int)min(32LL, rows - row0);
    const int n_el = nrows * C1;
    const float* src = cls + row0 * C1;
    if (stride == C1 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      // odd C1 needs no padding, so the block is one linear copy: 16-byte chunks (a quarter of the instructions)
      const int n16 = n_el >> 2;
      for (int q = lane; q < n16; q += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(tile + 4 * q)), "l"(src + 4 * q) : "memory");
      for (int e = (n16 << 2) + lane; e < n_el; e += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(tile + e)), "l"(src + e) : "memory");
    } else {
      // element e of the block -> tile[(e / C1) * stride + e % C1], kept incrementally (e advances by 32)
      int r = lane / C1, j = lane - r * C1;
      const int dr = 32 / C1, dj = 32 - dr * C1;
      for (int e = lane; e < n_el; e += 32) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(tile + r * stride + j)), "l"(src + e) : "memory");
        r += dr;
        j += dj;
        if (j >= C1) { j -= C1; ++r; }
      }
    }
    // (an L2 prefetch of the warp's next block here was measured: 81.9 -> 84.0 us, not kept — enough warps are in flight)
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    if (lane < nrows) {
      const float* x = tile + lane * stride;
      float m = x[0], mfg = TAUV_NEG_INF;
      int best_i = 0;
      for (int c = 1; c < C1; ++c) {
        const float v = x[c];
        if (ARGMAX && v > m) best_i = c;  // first maximum wins, like torch.argmax
        m = fmaxf(m, v);
        mfg = fmaxf(mfg, v);
      }
      // sum of exp(x - m) in four independent chains.  ex2.approx on x*log2(e): <= 2 ulp, i.e. ~2e-7 relative on the
      // score against the 1e-5 the parity tests allow (the reference's own softmax reduces in a different order too).
      const float L2E = 1.4426950408889634f;
      const float ml = m * L2E;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int c = 0;
      for (; c + 4 <= C1; c += 4) {
        s0 += ex2_fast(fmaf(x[c], L2E, -ml));
        s1 += ex2_fast(fmaf(x[c + 1], L2E, -ml));
        s2 += ex2_fast(fmaf(x[c + 2], L2E, -ml));
        s3 += ex2_fast(fmaf(x[c + 3], L2E, -ml));
      }
      for (; c < C1; ++c) s0 += ex2_fast(fmaf(x[c], L2E, -ml));
      const float sum = (s0 + s1) + (s2 + s3);
      // softmax is monotone per row: max_j>=1 softmax_j = exp(max_fg - max) / sum
      score[row0 + lane] = __fdiv_rn(ex2_fast(fmaf(mfg, L2E, -ml)), sum);
      if (ARGMAX) argmax_all[row0 + lane] = best_i;
    }
    __syncwarp();  // the tile is overwritten by the next block
  }
}

// ---- K2 ----------------------------------------------------------------------------------------
struct NmsArgs {
  const float* score;      // [B,N]
  const float4* box;       // decoded [B,N,4] or NULL
  const float4* enc;       // [B,N,4] (when box == NULL)
  const float4* anchor;    // [1 or B,N,4]
  int anchor_batch;
  float v0, v1;
  int N, top_k;
  float iou_thr, conf_thr;
  float q_lo, q_hi;        // iou_thr * (1 -+ 1e-6): an approximate quotient outside this band decides the test; 0 = off
  const float* cls;        // [B,N,C1] (only for keep_class)
  int C1;
  int32_t* keep_class;     // [B,top_k] argmax over all classes of the kept priors, or NULL
  int64_t* keep;           // [B,top_k]
  int32_t* n_keep;         // [B]
  float4* keep_box;        // [B,top_k,4] or NULL
  float* keep_score;       // [B,top_k] or NULL
  long long* trace;        // debug (-DTAUV_DEBUG, tools/detect_probe.py): clock64 of frame 0 at the phase boundaries
};

#ifdef TAUV_DEBUG
#define NMS_TRACE(i) do { if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[i] = clock64(); } while (0)
static long long* g_nms_trace = nullptr;
#else
#define NMS_TRACE(i) do { } while (0)
#endif

__global__ void __launch_bounds__(kNmsThreads, 1) nms_frame_kernel(NmsArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const int N = a.N;
  const int K = min(a.top_k, N);
  int p2 = 32;  // (at least one full tile of the pair test)
  while (p2 < K) p2 <<= 1;
  float4* bxs = reinterpret_cast<float4*>(smem_raw);                                // [p2] boxes of the ranked priors
  unsigned long long* sel = reinterpret_cast<unsigned long long*>(bxs + p2);       // [p2]
  float* cor = reinterpret_cast<float*>(sel + p2);                                  // [5][p2] SoA corners+area
  uint32_t* hist = reinterpret_cast<uint32_t*>(cor + 5 * (size_t)p2);               // [2048]
  uint32_t* sup = hist;                                                              // reused: [p2/32]
  __shared__ uint32_t ctl[8];
  __shared__ int s_wsum[kNmsThreads / 32];
  const float* sc = a.score + (size_t)b * N;

  asm volatile("griddepcontrol.wait;" ::: "memory");  // (the scores of the launch before; no-op without the attribute)
  NMS_TRACE(0);
  if (tid == 0) ctl[5] = 0;
  for (int i = tid; i < p2; i += kNmsThreads) sel[i] = 0ull;
  __syncthreads();
  auto load = [&](int i) -> unsigned long long { return make_composite(float_to_key(sc[i]), (uint32_t)i); };
  // ---- top K of the frame's N confidences, ranked (confidence desc, prior index asc)
  // Fast path (the shapes of the YOLACT head): every thread keeps its <= 20 scores in registers, ONE pass over the data.
  //   threshold: thread maxima are sorted inside each warp (shuffles only); the ceil(K/32)-th largest maximum of a warp
  //   has that many maxima above it, so the smallest of those 32 values has >= K scores at or above it — and, the
  //   maxima being the top ~5 % of the scores, not many more (~2K);
  //   the candidates at or above it go to a 1024-entry list and are ranked there.
  // The exact radix select over all N scores (the general path below) needs five passes and was a third of the kernel.
  bool ranked = false;
  if (N >= kNmsThreads && N <= kNmsThreads * kNmsRegs && K <= kNmsCand / 2) {
    unsigned long long* cand = reinterpret_cast<unsigned long long*>(hist);  // [kNmsCand] (the histogram is not in use)
    __shared__ uint32_t s_wthr[kNmsThreads / 32];
    float v[kNmsRegs];
#pragma unroll
    for (int u = 0; u < kNmsRegs; ++u) {
      const int i = u * kNmsThreads + tid;
      v[u] = __ldg(sc + min(i, N - 1));  // (unpredicated: there are only seven predicate registers)
    }
    // All twenty loads are in flight before the first use: left to itself the compiler issues them in pairs between
    // the key arithmetic (short live ranges), the warp stalls on the first use, and the pass costs 10 k cycles of
    // serialised L2 latency.  One empty asm that names every value cannot be placed before the last load.
    static_assert(kNmsRegs == 20, "the operand list below names twenty registers");
    asm volatile(""
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                   "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]), "+f"(v[16]),
                   "+f"(v[17]), "+f"(v[18]), "+f"(v[19]));
    NMS_TRACE(1);
    // from here on the registers hold the order-preserving keys; the threshold only needs the 32-bit keys (a tie on
    // the key at the threshold admits a few more candidates, never fewer than K)
    uint32_t mk = 0u;
#pragma unroll
    for (int u = 0; u < kNmsRegs; ++u) {
      const uint32_t k = float_to_key(v[u]);
      v[u] = __uint_as_float(k);
      mk = u * kNmsThreads + tid < N ? max(mk, k) : mk;
    }
    cand[tid] = 0ull;
    static_assert(kNmsCand == kNmsThreads, "one candidate slot per thread");
    uint32_t mx = mk;
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {  // warp bitonic sort, descending: lane 0 ends with the largest
      const bool desc = size == 32 || (lane & size) == 0;
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        const uint32_t y = __shfl_xor_sync(0xffffffffu, mx, stride);
        mx = (((lane & stride) == 0) == desc) ? max(mx, y) : min(mx, y);
      }
    }
    // Threshold.  The ceil(K/32)-th largest maximum of a warp has that many maxima of its own warp at or above it, so the
    // SMALLEST of these 32 values has >= K scores at or above it — but also about 3K (measured: 572 candidates for
    // K = 200), and the ranking below is quadratic in that.  So every warp counts, in the other warps' sorted maxima,
    // how many are at or above ITS value, and the LARGEST value with a count >= K is taken (~1.2K candidates).
    __shared__ uint32_t s_max[32 * 33];  // warp w's sorted maxima at [w * 33 ...] (odd stride: the searches below do not collide)
    const uint32_t wthr = __shfl_sync(0xffffffffu, mx, (K + 31) / 32 - 1);
    s_max[warp * 33 + lane] = mx;
    __syncthreads();
    {
      const uint32_t* run = s_max + lane * 33;
      int c = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1)
        if (run[c + step - 1] >= wthr) c += step;
      if (run[c] >= wthr) ++c;
      const int total = __reduce_add_sync(0xffffffffu, c);
      if (lane == 0) s_wthr[warp] = total >= K ? wthr : 0u;
    }
    __syncthreads();
    NMS_TRACE(2);
    const uint32_t T = __reduce_max_sync(0xffffffffu, s_wthr[lane]);
    // append without atomics (640 warp-level adds to one shared counter serialise for ~13 k cycles) and without votes:
    // every thread notes which of its scores qualify, one block-wide prefix over the counts gives it a private range
    // of the list (the order inside the list does not matter: it is ranked next)
    uint32_t tm = 0u;
#pragma unroll
    for (int u = 0; u < kNmsRegs; ++u)
      if (u * kNmsThreads + tid < N && __float_as_uint(v[u]) >= T) tm |= 1u << u;
    const unsigned my_cnt = (unsigned)__popc(tm);
    unsigned incl = my_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_wsum[warp] = (int)incl;
    __syncthreads();
    unsigned base = incl - my_cnt, n_cand = 0;
    {
      const unsigned w = (unsigned)s_wsum[lane];
      unsigned wincl = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, wincl, o);
        if (lane >= o) wincl += y;
      }
      n_cand = __shfl_sync(0xffffffffu, wincl, 31);
      base += __shfl_sync(0xffffffffu, wincl - w, warp);
    }
    if (n_cand <= (unsigned)kNmsCand) {
#pragma unroll
      for (int u = 0; u < kNmsRegs; ++u)
        if ((tm >> u) & 1u) cand[base++] = make_composite(__float_as_uint(v[u]), (uint32_t)(u * kNmsThreads + tid));
    }
    __syncthreads();
    NMS_TRACE(3);
#ifdef TAUV_DEBUG
    if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[8] = n_cand;
#endif
    if (n_cand <= (unsigned)kNmsCand) {  // (a plateau of equal scores can overflow the list: general path)
      // Rank without a block-wide sort (its 10-15 shared-memory steps cost two 32-warp barriers each: 13 k cycles).
      // Every warp sorts its own 32 candidates with shuffles; a candidate's rank is then its position in its own run
      // plus, for every other run, the number of entries above it — a 6-probe binary search per run, all runs
      // independent of each other.  Composites are distinct, so ranks are; the first K land in sel[].
      const int nruns = ((int)n_cand + 31) >> 5;
      unsigned long long x = 0ull;
      if (warp < nruns) {
        x = cand[tid];
#pragma unroll
        for (int size = 2; size <= 32; size <<= 1) {
          const bool desc = size == 32 || (lane & size) == 0;
#pragma unroll
          for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, stride);
            const bool take_max = ((lane & stride) == 0) == desc;
            x = (take_max == (y > x)) ? y : x;
          }
        }
        cand[tid] = x;
      }
      __syncthreads();
      if (warp < nruns && x != 0ull) {  // (a composite is never 0: its low word is the complement of a prior index)
        int rank = lane;
        for (int r0 = 0; r0 < nruns; r0 += 4) {  // four searches at a time: their probes are independent loads
          const unsigned long long* run[4];
          int c[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            run[q] = cand + min(r0 + q, kNmsCand / 32 - 1) * 32;
            c[q] = 0;
          }
#pragma unroll
          for (int step = 16; step > 0; step >>= 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (run[q][c[q] + step - 1] > x) c[q] += step;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (run[q][c[q]] > x) ++c[q];
            rank += (r0 + q == warp || r0 + q >= nruns) ? 0 : c[q];
          }
        }
        if (rank < K) sel[rank] = x;
      }
      ranked = true;
    }
    __syncthreads();  // (cand aliases hist / sup)
    if (tid == 0) ctl[5] = 0;
    __syncthreads();
  }
  if (!ranked) {
    const unsigned long long T = block_kth_largest<kNmsThreads>(load, N, K, hist, ctl);
    for (int i0 = 0; i0 < N; i0 += 4 * kNmsThreads) {  // (four independent loads in flight; one atomic per warp and strip)
      unsigned long long c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kNmsThreads + tid;
        c[u] = i < N ? load(i) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool take = i0 + u * kNmsThreads + tid < N && c[u] >= T;
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        if (bal) {
          unsigned base = 0;
          if (lane == 0) base = atomicAdd(&ctl[5], (unsigned)__popc(bal));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (take) sel[base + __popc(bal & ((1u << lane) - 1u))] = c[u];
        }
      }
    }
    __syncthreads();
    block_bitonic_sort_desc_reg<kNmsThreads>(sel, p2);
  }

  NMS_TRACE(4);
  // the class rows of the ranked priors are wanted at the very end (argmax of the kept ones) and left the L2 long ago:
  // ask for them now, so that the DRAM / TLB latency of ~K scattered rows runs under the phases in between
  if (a.keep_class) {
    const size_t row_bytes = (size_t)a.C1 * 4;
    for (int r = warp; r < K; r += kNmsThreads / 32) {
      const char* row = reinterpret_cast<const char*>(a.cls + ((size_t)b * N + composite_idx(sel[r])) * a.C1);
      const size_t off = (size_t)lane * 128;
      if (off < row_bytes + 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + min(off, row_bytes - 4)));
    }
  }
  // boxes of the ranked priors -> corners in shared memory (decoded on the fly in detect mode)
  for (int r = tid; r < K; r += kNmsThreads) {
    const uint32_t idx = composite_idx(sel[r]);
    float4 bx;
    if (a.box) {
      bx = a.box[(size_t)b * N + idx];
    } else {
      const float4 an = a.anchor[(a.anchor_batch == 1 ? 0 : (size_t)b * N) + idx];
      bx = decode_one(a.enc[(size_t)b * N + idx], an, a.v0, a.v1);
    }
    const Corners c = to_corners(bx);
    // (a NaN corner poisons the area: the pair test then takes the IEEE path, whose result is NaN either way)
    const bool nan_corner = c.y0 != c.y0 || c.x0 != c.x0 || c.y1 != c.y1 || c.x1 != c.x1;
    cor[r] = c.y0; cor[p2 + r] = c.x0; cor[2 * p2 + r] = c.y1; cor[3 * p2 + r] = c.x1;
    cor[4 * p2 + r] = nan_corner ? __int_as_float(0x7fc00000) : c.area;
    bxs[r] = bx;
  }
  const int T32 = (K + 31) >> 5;
  for (int i = tid; i < T32; i += kNmsThreads) sup[i] = 0u;
  __syncthreads();
  NMS_TRACE(5);

  // Fast NMS: column j is suppressed iff some earlier-ranked i < j has iou(i,j) > thr (nms.py:19-24; a NaN
  // IoU also suppresses because `NaN <= thr` is False).  Tiles (I <= J) of 32x32 pairs, one warp per tile.
  const int ntiles = T32 * (T32 + 1) / 2;
  for (int t = warp; t < ntiles; t += kNmsThreads / 32) {
    // unrank t -> (J, I) with I <= J, tiles enumerated column by column
    int J = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((J + 1) * (J + 2) / 2 <= t) ++J;
    while (J * (J + 1) / 2 > t) --J;
    const int I = t - J * (J + 1) / 2;
    const int j = J * 32 + lane;
    const int jc = min(j, K - 1);  // (lanes past the last column work on a valid one and are masked at the vote)
    Corners cj;
    cj.y0 = cor[jc]; cj.x0 = cor[p2 + jc]; cj.y1 = cor[2 * p2 + jc]; cj.x1 = cor[3 * p2 + jc]; cj.area = cor[4 * p2 + jc];
    // `fl(inter / union) <= thr` without the division (its ~60 instructions, needed by some lane on almost every step,
    // were a third of the kernel): inter * rcp.approx(union) is within 2^-22 of the real quotient, so outside a band of
    // 1e-6 around thr it decides the rounded quotient's test too.  Pairs inside the band, and every special case (NaN
    // corners carry a NaN area; empty, tiny, huge or infinite unions), are noted and take the IEEE division afterwards.
    // The 32 rows of the tile are straight-line code (no branch between them: their loads and reciprocals overlap).
    const bool band = a.q_hi > 0.0f;
    bool s = false;
    unsigned exact = 0u;
#pragma unroll 8
    for (int t = 0; t < 32; ++t) {
      const int i = I * 32 + t;  // (< p2: rows past K hold stale values and are masked by i < j < K)
      const float y0 = cor[i], x0 = cor[p2 + i], y1 = cor[2 * p2 + i], x1 = cor[3 * p2 + i], ar = cor[4 * p2 + i];
      const float ih = fmaxf(__fsub_rn(fminf(y1, cj.y1), fmaxf(y0, cj.y0)), 0.0f);
      const float iw = fmaxf(__fsub_rn(fminf(x1, cj.x1), fmaxf(x0, cj.x0)), 0.0f);
      const float inter = __fmul_rn(ih, iw);
      const float uni = __fsub_rn(__fadd_rn(ar, cj.area), inter);
      float rcp;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(uni));
      const float q = __fmul_rn(inter, rcp);
      const bool live = i < j;  // strictly upper triangle
      const bool fast = band && uni > 1e-30f && uni < 1e30f;
      const bool above = fast && q > a.q_hi, below = fast && q < a.q_lo;
      s = s || (live && above);
      exact |= (live && !above && !below ? 1u : 0u) << t;
    }
    while (exact) {  // (rare; lane-dependent trip count: no warp-level operation inside)
      const int i = I * 32 + __ffs(exact) - 1;
      exact &= exact - 1u;
      Corners ci;
      ci.y0 = cor[i]; ci.x0 = cor[p2 + i]; ci.y1 = cor[2 * p2 + i]; ci.x1 = cor[3 * p2 + i]; ci.area = cor[4 * p2 + i];
      s = s || !(iou_pair(ci, cj) <= a.iou_thr);
    }
    s = s && j < K;
    const unsigned bal = __ballot_sync(0xffffffffu, s);
    if (lane == 0 && bal) atomicOr(&sup[J], bal);
  }
  __syncthreads();

  NMS_TRACE(6);
  // ordered compaction of the survivors (the corner arrays are dead: their space lists the kept prior indices)
  uint32_t* kidx = reinterpret_cast<uint32_t*>(cor);
  int base = 0;
  const float conf_thr = a.conf_thr;
  for (int start = 0; start < K; start += kNmsThreads) {
    const int r = start + tid;
    bool kp = false;
    unsigned long long c = 0ull;
    if (r < K) {
      c = sel[r];
      const float conf = key_to_float(composite_key(c));
      kp = !((sup[r >> 5] >> (r & 31)) & 1u) && (conf >= conf_thr);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, kp);
    if (lane == 0) s_wsum[warp] = __popc(bal);
    __syncthreads();
    int pos = base + __popc(bal & ((1u << lane) - 1u));
    int tot;
    {
      static_assert(kNmsThreads / 32 == 32, "one lane per warp count");
      const int w = s_wsum[lane];
      int wincl = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, wincl, o);
        if (lane >= o) wincl += y;
      }
      pos += __shfl_sync(0xffffffffu, wincl - w, warp);
      tot = __shfl_sync(0xffffffffu, wincl, 31);
    }
    if (kp) {
      a.keep[(size_t)b * a.top_k + pos] = (int64_t)composite_idx(c);
      if (a.keep_score) a.keep_score[(size_t)b * a.top_k + pos] = key_to_float(composite_key(c));
      if (a.keep_box) a.keep_box[(size_t)b * a.top_k + pos] = bxs[r];
      kidx[pos] = composite_idx(c);
    }
    base += tot;
    __syncthreads();
  }
  if (tid == 0) a.n_keep[b] = base;
  NMS_TRACE(7);
  // argmax over ALL classes of the kept priors (yolact_node.py:129 / evaluate_batch.py:93-95): one warp per kept prior,
  // eight priors' rows in flight per warp (the rows left the L2 long ago: one DRAM round trip instead of eight)
  if (a.keep_class) {
    const int C1 = a.C1;
    const float* cls_b = a.cls + (size_t)b * N * C1;
    int32_t* kc = a.keep_class + (size_t)b * a.top_k;
    if (C1 <= 96) {
      for (int r0 = warp; r0 < base; r0 += 8 * (kNmsThreads / 32)) {
        float x[8][3];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int r = r0 + q * (kNmsThreads / 32);
          const float* row = cls_b + (size_t)kidx[min(r, base - 1)] * C1;
#pragma unroll
          for (int t = 0; t < 3; ++t) x[q][t] = lane + 32 * t < C1 ? __ldg(row + lane + 32 * t) : TAUV_NEG_INF;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int r = r0 + q * (kNmsThreads / 32);
          if (r >= base) break;  // (warp-uniform)
          float best_v = TAUV_NEG_INF;
          int best_i = 0x7fffffff;
#pragma unroll
          for (int t = 0; t < 3; ++t)
            if (x[q][t] > best_v) { best_v = x[q][t]; best_i = lane + 32 * t; }
          // (best_v is never NaN: `>` does not admit one) first maximum across the lanes with two warp reductions
          const uint32_t key = float_to_key(best_v);
          const uint32_t top = __reduce_max_sync(0xffffffffu, key);
          const int first = __reduce_min_sync(0xffffffffu, key == top ? best_i : 0x7fffffff);
          if (lane == 0) kc[r] = first == 0x7fffffff ? 0 : first;
        }
      }
    } else {
      for (int r = warp; r < base; r += kNmsThreads / 32) {
        const float* row = cls_b + (size_t)kidx[r] * C1;
        float best_v = TAUV_NEG_INF;
        int best_i = 0x7fffffff;
        for (int j = lane; j < C1; j += 32) {
          const float t = row[j];
          if (t > best_v) { best_v = t; best_i = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
          const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
          if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
        }
        if (lane == 0) kc[r] = best_i == 0x7fffffff ? 0 : best_i;
      }
    }
  }
  NMS_TRACE(9);
}

static int launch_scores(const float* cls, long long rows, int C1, float* score, int32_t* argmax_all, cudaStream_t st) {
  if (C1 <= 256 && !debug_env("TAUV_SCORES_OLD")) {
    const int stride = C1 | 1;
    const size_t smem = (size_t)kScoreTileWarps * 32 * stride * 4;
    long long blocks = (rows + 32 * kScoreTileWarps - 1) / (32 * kScoreTileWarps);
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (argmax_all) {
      TAUV_CUDA(ensure_dynamic_smem((const void*)(scores_tile_kernel<true>), smem));
      scores_tile_kernel<true><<<(unsigned)blocks, kScoreTileWarps * 32, smem, st>>>(cls, rows, C1, stride, score, argmax_all);
    } else {
      TAUV_CUDA(ensure_dynamic_smem((const void*)(scores_tile_kernel<false>), smem));
      scores_tile_kernel<false><<<(unsigned)blocks, kScoreTileWarps * 32, smem, st>>>(cls, rows, C1, stride, score, argmax_all);
    }
    TAUV_LAUNCH_CHECK("scores_tile_kernel");
    return 0;
  }
  const long long blocks_needed = (rows + 32 * 8 - 1) / (32 * 8);  // 8 warps per CTA, 32 rows per warp
  long long blocks = blocks_needed;
  const long long cap = (long long)num_sms() * 8 * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const int nch = (C1 + 31) / 32;
#define TAUV_SCORES(NCH)                                                                                       \
  do {                                                                                                          \
    if (argmax_all) scores_kernel<NCH, true><<<(unsigned)blocks, 256, 0, st>>>(cls, rows, C1, score, argmax_all); \
    else scores_kernel<NCH, false><<<(unsigned)blocks, 256, 0, st>>>(cls, rows, C1, score, argmax_all);          \
  } while (0)
  switch (nch) {
    case 1: TAUV_SCORES(1); break;
    case 2: TAUV_SCORES(2); break;
    case 3: TAUV_SCORES(3); break;
    case 4: TAUV_SCORES(4); break;
    default: TAUV_SCORES(0); break;
  }
#undef TAUV_SCORES
  TAUV_LAUNCH_CHECK("scores_kernel");
  return 0;
}

static size_t nms_smem(int K) {
  int p2 = 32;
  while (p2 < K) p2 <<= 1;
  return (size_t)p2 * 16 + (size_t)p2 * 8 + (size_t)p2 * 5 * 4 + kRadixBins * 4;
}

static int run_nms(const float* cls, const float* box, const float* enc, const float* anchor, int anchor_batch, float v0,
                   float v1, int B, int N, int C1, int n_frames, int top_k, float iou_thr, float conf_thr,
                   int64_t* keep, int32_t* n_keep, float* keep_box, float* keep_score, int32_t* keep_class, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
  TAUV_REQUIRE(cls && keep && n_keep, TAUV_E_NULL, "cls/keep/n_keep must not be NULL");
  TAUV_REQUIRE(B > 0 && N > 0 && C1 >= 2 && top_k > 0, TAUV_E_SHAPE, "bad shape B=%d N=%d C1=%d top_k=%d", B, N, C1, top_k);
  TAUV_REQUIRE(n_frames > 0 && n_frames <= B, TAUV_E_SHAPE, "n_frames=%d must be in [1,%d]", n_frames, B);
  TAUV_REQUIRE(top_k <= kNmsMaxTopK, TAUV_E_UNSUPPORTED, "top_k=%d exceeds the built-in limit %d", top_k, kNmsMaxTopK);
  const size_t need = align_up((size_t)n_frames * N * 4, 256);
  TAUV_REQUIRE(ws && (uintptr_t)ws % 256 == 0 && ws_bytes >= need, TAUV_E_WORKSPACE, "workspace %zu < required %zu", ws_bytes, need);
  float* score = reinterpret_cast<float*>(ws);
  if (int e = launch_scores(cls, (long long)n_frames * N, C1, score, nullptr, st)) return e;
  NmsArgs a;
  a.score = score; a.box = (const float4*)box; a.enc = (const float4*)enc; a.anchor = (const float4*)anchor;
  a.anchor_batch = anchor_batch; a.v0 = v0; a.v1 = v1; a.N = N; a.top_k = top_k; a.iou_thr = iou_thr; a.conf_thr = conf_thr;
  a.q_lo = a.q_hi = 0.0f;
  if (iou_thr > 1e-30f && iou_thr < 1e30f) {  // (normal, positive: the band below is well inside the float range)
    a.q_lo = (float)((double)iou_thr * (1.0 - 1e-6));
    a.q_hi = (float)((double)iou_thr * (1.0 + 1e-6));
  }
  a.cls = cls; a.C1 = C1; a.keep_class = keep_class;
  a.keep = keep; a.n_keep = n_keep; a.keep_box = (float4*)keep_box; a.keep_score = keep_score;
#ifdef TAUV_DEBUG
  a.trace = g_nms_trace;
#else
  a.trace = nullptr;
#endif
  const size_t smem = nms_smem(top_k < N ? top_k : N);
  TAUV_CUDA(ensure_dynamic_smem((const void*)(nms_frame_kernel), smem));
  {
    // programmatic dependent launch: the CTAs are set up while the scores kernel drains and wait (griddepcontrol.wait,
    // first thing in the kernel) for its completion and memory flush — the launch latency leaves the critical path
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_frames);
    cfg.blockDim = dim3(kNmsThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TAUV_CUDA(cudaLaunchKernelEx(&cfg, nms_frame_kernel, a));
  }
  TAUV_LAUNCH_CHECK("nms_frame_kernel");
  return 0;
}

}  // namespace tauv

using namespace tauv;

extern "C" int tauv_yolact_scores(const float* cls, int B, int N, int C1, float* score, int32_t* argmax_all,
                                  tauv_stream_t stream) {
  TAUV_REQUIRE(cls && score, TAUV_E_NULL, "cls/score must not be NULL");
  TAUV_REQUIRE(B > 0 && N > 0 && C1 >= 2, TAUV_E_SHAPE, "bad shape B=%d N=%d C1=%d", B, N, C1);
  return launch_scores(cls, (long long)B * N, C1, score, argmax_all, (cudaStream_t)stream);
}

extern "C" size_t tauv_yolact_nms_workspace_bytes(int B, int N, int C1, int top_k) {
  (void)C1; (void)top_k;
  if (B <= 0 || N <= 0) return 0;
  return align_up((size_t)B * N * 4, 256);
}

extern "C" int tauv_yolact_fast_nms(const float* cls, const float* box, int B, int N, int C1, int n_frames, int top_k,
                                    float iou_threshold, float confidence_threshold, int64_t* keep, int32_t* n_keep,
                                    void* workspace, size_t workspace_bytes, tauv_stream_t stream) {
  TAUV_REQUIRE(box, TAUV_E_NULL, "box must not be NULL");
  TAUV_REQUIRE((uintptr_t)box % 16 == 0, TAUV_E_ALIGN, "box must be 16-byte aligned");
  return run_nms(cls, box, nullptr, nullptr, 1, 0.f, 0.f, B, N, C1, n_frames, top_k, iou_threshold,
                 confidence_threshold, keep, n_keep, nullptr, nullptr, nullptr, workspace, workspace_bytes,
                 (cudaStream_t)stream);
}

extern "C" int tauv_yolact_detect(const float* cls, const float* enc, const float* anchor, int B, int N, int C1,
                                  int anchor_batch, float v0, float v1, int top_k, float iou_threshold,
                                  float confidence_threshold, int64_t* keep, int32_t* n_keep, float* keep_box,
                                  float* keep_score, int32_t* keep_class, void* workspace, size_t workspace_bytes,
                                  tauv_stream_t stream) {
  TAUV_REQUIRE(enc && anchor, TAUV_E_NULL, "enc/anchor must not be NULL");
  TAUV_REQUIRE(anchor_batch == 1 || anchor_batch == B, TAUV_E_SHAPE, "anchor batch %d must be 1 or %d", anchor_batch, B);
  TAUV_REQUIRE((uintptr_t)enc % 16 == 0 && (uintptr_t)anchor % 16 == 0 && (uintptr_t)keep_box % 16 == 0, TAUV_E_ALIGN,
               "box tensors must be 16-byte aligned");
  return run_nms(cls, nullptr, enc, anchor, anchor_batch, v0, v1, B, N, C1, B, top_k, iou_threshold,
                 confidence_threshold, keep, n_keep, keep_box, keep_score, keep_class, workspace, workspace_bytes,
                 (cudaStream_t)stream);
}

#ifdef TAUV_DEBUG
// Debug hook for tools/detect_probe.py (only in -DTAUV_DEBUG builds; process-global, not thread-safe).
extern "C" void tauv_debug_nms_trace(long long* buf) { tauv::g_nms_trace = buf; }
#endif
