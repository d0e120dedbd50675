// centernet_decode.cu — CenterNet head decode on sm_100a.
//
// Replaces (reference file:line under src/tauv_vision/centernet/model/):
//   decode.py:182      F.sigmoid(prediction.heatmap)
//   decode.py:239-252  heatmap_nms   (3x3 max-pool == test)
//   decode.py:255-279  heatmap_detect (joint top-k over C*H*W per frame)
//   decode.py:204-234  per-detection gather + box arithmetic
//
// Pipeline (2 launches for a whole batch):
//   K1 tile_topk_kernel : one CTA per "item" (a band of rows of one (frame, class) plane).  The
//      band streams HBM -> shared through a ring of 1-D bulk copies (cp.async.bulk + mbarrier);
//      peaks are found on the logits (sigmoid is monotone), the sigmoid is evaluated only for
//      survivors, and the item's top-k candidates (64-bit composite keys) go to a small global
//      candidate table.  The logits are read from HBM exactly once; nothing dense is written.
//   K2 merge_kernel     : one CTA per frame selects the frame's top-k from its items' candidates
//      (radix select + bitonic sort), fills zero-score slots like a dense stable top-k would,
//      and (optionally) gathers size/offset/depth and does the box arithmetic.
#include "common.cuh"

namespace tauv {

constexpr int kTileThreads = 256;
constexpr int kStages = 6;          // ring slots
constexpr int kChunkElems = 2048;   // target elements per ring slot
constexpr int kItemElems = 16384;   // target elements per item
constexpr int kMergeThreads = 1024;
constexpr int kMaxK = 4096;

struct TopkPlan {
  int rows_per_chunk;   // R
  int slot_elems;       // R*W rounded up to a multiple of 4
  int rows_per_item;
  int items_per_plane;
  int items_per_frame;  // C * items_per_plane
  int cap;              // candidate-list capacity (entries)
  int soft;             // prune when the list grows beyond this
  int bulk;             // 1: bulk-copy ring, 0: plain loads
  size_t smem_bytes;
  size_t cand_bytes, count_bytes, thr_bytes;
};

static int make_plan(int B, int C, int H, int W, int k, const void* ptr, TopkPlan* p) {
  p->bulk = (W % 4 == 0) && ((uintptr_t)ptr % 16 == 0);
  int R = kChunkElems / W;
  if (R < 1) R = 1;
  if (R > H) R = H;
  p->rows_per_chunk = R;
  p->slot_elems = (int)align_up((size_t)R * W, 4);
  // item: whole chunks, about kItemElems elements, but enough items to fill the machine
  int rows_item = (int)align_up((size_t)((kItemElems + W - 1) / W), (size_t)R);
  if (rows_item < R) rows_item = R;
  const long long planes = (long long)B * C;
  const int sms = num_sms();
  while (rows_item > R && planes * ((H + rows_item - 1) / rows_item) < 3LL * sms) {
    int next = (int)align_up((size_t)(rows_item / 2), (size_t)R);
    if (next >= rows_item) break;
    rows_item = next;
  }
  if (rows_item > H) rows_item = H;
  p->rows_per_item = rows_item;
  p->items_per_plane = (H + rows_item - 1) / rows_item;
  p->items_per_frame = C * p->items_per_plane;
  // The list is pruned to its top-k as soon as it holds more than 2k entries, which also gives the item (and,
  // through the per-frame published threshold, every other item of the frame) a rejection threshold early.
  p->soft = 2 * k;
  p->cap = p->soft + p->slot_elems;
  size_t smem = align_up((size_t)kStages * p->slot_elems * sizeof(float), 16);
  smem += (size_t)p->cap * 8 + kRadixBins * 4 + kStages * 8 + 64 + (size_t)(R + 4) * 2 * sizeof(int);
  p->smem_bytes = smem;
  p->cand_bytes = align_up((size_t)B * p->items_per_frame * (size_t)k * 8, 256);
  p->count_bytes = align_up((size_t)B * p->items_per_frame * 4, 256);
  p->thr_bytes = align_up((size_t)B * 4, 256);
  return 0;
}

// ----------------------------------------------------------------------------------------------
// K1
// ----------------------------------------------------------------------------------------------
struct TileArgs {
  const float* hm;
  int B, C, H, W, k;
  int R, slot_elems, rows_per_item, items_per_plane;
  int cap, soft;
  unsigned long long* cand;  // [B*items_per_frame][k]
  int* cand_count;           // [B*items_per_frame]
  uint32_t* frame_thr;       // [B] published rejection threshold (order-preserving key of a logit / value), 0 = none
};

// x < m can still tie after the sigmoid (saturation, or a sub-ulp gap): the reference compares sigmoid values
// (decode.py:252), so those rare cases are decided on the sigmoids themselves.
__device__ __noinline__ bool sigmoid_tie(float x, float m) { return sigmoid_ref(x) == sigmoid_ref(m); }

// A logit x_c such that every x < x_c has sigmoid(x) strictly below the score s_k (with a few-ulp guard for the
// last-bit wobble of expf).  Returns the order-preserving key of x_c, or 0 when no such logit is found cheaply
// (saturated scores) — then nothing is rejected up front and the exact prune alone bounds the list.
__device__ uint32_t reject_key_for_score(float s_k) {
  if (!(s_k > 0.0f) || !(s_k < 1.0f)) return 0u;
  const float x0 = logf(__fdiv_rn(s_k, __fsub_rn(1.0f, s_k)));
  float margin = 1e-3f * fmaxf(1.0f, fabsf(x0));
  for (int t = 0; t < 4; ++t, margin *= 8.0f) {
    const float xc = x0 - margin;
    if (sigmoid_ref(xc) < s_k * (1.0f - 4e-6f)) return float_to_key(xc);
  }
  return 0u;
}

template <int MODE, bool BULK>
__global__ void __launch_bounds__(kTileThreads) tile_topk_kernel(TileArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int W = a.W, H = a.H, R = a.R;
  float* ring = reinterpret_cast<float*>(smem_raw);
  size_t off = align_up((size_t)kStages * a.slot_elems * sizeof(float), 16);
  unsigned long long* list = reinterpret_cast<unsigned long long*>(smem_raw + off);
  off += (size_t)a.cap * 8;
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + off);
  off += kRadixBins * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + off);
  off += kStages * 8;
  uint32_t* ctl = reinterpret_cast<uint32_t*>(smem_raw + off);  // [0..3] select ctl, [4] count, [5] emit, [6] n_conv
  off += 64;
  int* rowtab = reinterpret_cast<int*>(smem_raw + off);          // [2][R+4] ring offsets of rows ra-1 .. rb
  __shared__ unsigned long long s_thr;                           // push filter (see push())
  __shared__ int s_wsum[kTileThreads / 32], s_base;

  // Block -> item mapping interleaves the frames (consecutive blocks work on different frames), so that the
  // first items of EVERY frame finish early and publish a threshold the frame's remaining items can use.
  const int items_per_frame = a.C * a.items_per_plane;
  const int frame = blockIdx.x % a.B;
  const int item_in_frame = blockIdx.x / a.B;
  const int item = frame * items_per_frame + item_in_frame;
  const int ip = item_in_frame % a.items_per_plane;
  const int c_in_frame = item_in_frame / a.items_per_plane;
  const long long plane = (long long)frame * a.C + c_in_frame;
  const int r0 = ip * a.rows_per_item;
  const int r1 = min(H, r0 + a.rows_per_item);
  const int lo = max(r0 - 1, 0);
  const int hi = min(r1 + 1, H);
  const int nchunks = (hi - lo + R - 1) / R;
  const float* base = a.hm + (size_t)plane * H * W;
  const uint32_t plane_flat0 = (uint32_t)c_in_frame * (uint32_t)(H * W);

  if (tid == 0) {
    if (BULK) {
      for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
      mbar_fence_init();
    }
    ctl[4] = 0;
    ctl[6] = 0;
    s_thr = (unsigned long long)(*reinterpret_cast<volatile uint32_t*>(a.frame_thr + frame)) << 32;
  }
  __syncthreads();

  auto issue_chunk = [&](int j) {  // thread 0 only
    const int row = lo + j * R;
    const int rows = min(R, hi - row);
    const uint32_t bytes = (uint32_t)rows * W * 4u;
    uint64_t* bar = &bars[j % kStages];
    mbar_expect_tx(bar, bytes);
    bulk_g2s(ring + (size_t)(j % kStages) * a.slot_elems, base + (size_t)row * W, bytes, bar);
  };
  if (BULK && tid == 0) {
    const int pre = min(nchunks, kStages);
    for (int j = 0; j < pre; ++j) issue_chunk(j);
  }

  auto row_off = [&](int r) -> int {  // ring offset (in floats) of plane row r, -1 when the row does not exist
    if (r < 0 || r >= H) return -1;
    const int q = r - lo;
    const int ch = q / R;
    return (ch % kStages) * a.slot_elems + (q - ch * R) * W;
  };
  auto fill_rowtab = [&](int j) {  // rows ra-1 .. rb of step j
    const int ra = max(lo + j * R, r0);
    const int rb = min(min(lo + (j + 1) * R, hi), r1);
    int* tab = rowtab + (j & 1) * (R + 4);
    for (int i = tid; i < rb - ra + 2; i += kTileThreads) tab[i] = row_off(ra - 1 + i);
  };
  fill_rowtab(0);
  __syncthreads();  // table of step 0 visible to everyone (later tables ride on the end-of-step barrier)

  int* count_p = reinterpret_cast<int*>(&ctl[4]);
  int my_end = 0;  // highest (slot+1) this thread produced in the current step
  unsigned long long thr = 0ull;
  float thr_f = TAUV_NEG_INF;

  // A list entry is a 64-bit "pre-composite": order-preserving key of the VALUE in the high word, ~flat index in
  // the low word.  In RAW mode that already is the final sort key.  In SIGMOID_PEAK mode the value is the logit;
  // the sigmoid (and the final key) is applied later, densely, by convert().  Entries below `thr` are provably
  // outside the frame's top-k and never enter the list.
  auto push = [&](float x, uint32_t flat) {
    const unsigned long long pre = make_composite(float_to_key(x), flat);
    if (pre < thr) return;
    const int slot = atomicAdd(count_p, 1);
    list[slot] = pre;  // capacity is guaranteed by the prune policy (cap = soft + chunk elements)
    my_end = max(my_end, slot + 1);
  };

  // logit pre-composites [n_conv, n) -> score composites (0 for a sigmoid that underflowed to 0: zero-valued cells
  // are supplied by the merge kernel's filler, like non-peaks)
  auto convert = [&](int n) {
    if (MODE == TAUV_TOPK_SIGMOID_PEAK) {
      for (int i = (int)ctl[6] + tid; i < n; i += kTileThreads) {
        const unsigned long long pre = list[i];
        const float s = sigmoid_ref(key_to_float(composite_key(pre)));
        list[i] = (s > 0.0f) ? (((unsigned long long)float_to_key(s) << 32) | (pre & 0xffffffffull)) : 0ull;
      }
      __syncthreads();
    }
  };
  // after a select with threshold T over score composites: new push filter + publication
  auto publish = [&](unsigned long long T, int n_kept) {  // thread 0 only
    if (n_kept < a.k) return;
    uint32_t key;
    if (MODE == TAUV_TOPK_SIGMOID_PEAK) {
      key = reject_key_for_score(key_to_float(composite_key(T)));
      if (key == 0u) return;
      const unsigned long long t = (unsigned long long)key << 32;
      if (t > s_thr) s_thr = t;
    } else {
      key = composite_key(T);
      if (T > s_thr) s_thr = T;
    }
    atomicMax(a.frame_thr + frame, key);
  };

  for (int j = 0; j < nchunks; ++j) {
    // ---- make chunk j (and j+1: the row below the last row of j) resident ----
    if (BULK) {
      mbar_wait(&bars[j % kStages], (uint32_t)((j / kStages) & 1));
      if (j + 1 < nchunks) mbar_wait(&bars[(j + 1) % kStages], (uint32_t)(((j + 1) / kStages) & 1));
    } else {
      __syncthreads();
      for (int jj = max(j - 1, 0); jj <= min(j + 1, nchunks - 1); ++jj) {
        const int row = lo + jj * R;
        const int n = min(R, hi - row) * W;
        float* dst = ring + (size_t)(jj % kStages) * a.slot_elems;
        const float* src = base + (size_t)row * W;
        for (int i = tid; i < n; i += kTileThreads) dst[i] = __ldg(src + i);
      }
      __syncthreads();
    }
    // what other items of this frame have published meanwhile: the load is issued here and consumed at the end
    // of the step (folded into s_thr for the next step), so its L2 latency hides behind the step's work
    uint32_t pub_key = 0u;
    if (tid == 0) pub_key = *reinterpret_cast<volatile uint32_t*>(a.frame_thr + frame);
    thr = s_thr;
    thr_f = key_to_float(composite_key(thr));
    if (composite_key(thr) == 0u) thr_f = TAUV_NEG_INF;
    my_end = 0;
    const int ra = max(lo + j * R, r0);
    const int rb = min(min(lo + (j + 1) * R, hi), r1);
    const int* tab = rowtab + (j & 1) * (R + 4);  // tab[i] = ring offset of row ra-1+i

    if (BULK) {
      // ---- vector path: one float4 strip per task; only strips holding a value >= thr are examined further ----
      const int S = W >> 2;
      const int ntasks = (rb - ra) * S;
      int ri = tid / S, cs = tid - ri * S;         // task = (row ra+ri, strip cs)
      const int dr = kTileThreads / S, dc = kTileThreads - dr * S;
#pragma unroll 1
      for (int task = tid; task < ntasks; task += kTileThreads) {
        const int o1 = tab[ri + 1];
        const int col = cs << 2;
        const float4 x = *reinterpret_cast<const float4*>(ring + o1 + col);
        const float mx = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
        if (mx >= thr_f) {
          const uint32_t flat = plane_flat0 + (uint32_t)((ra + ri) * W + col);
          if (MODE == TAUV_TOPK_SIGMOID_PEAK) {
            // 3x3 neighbourhood of the strip: columns col-1 .. col+4 of rows r-1, r, r+1 (-inf outside the plane)
            const int o0 = tab[ri], o2 = tab[ri + 2];
            float cm[6];  // column-wise max over the three rows
            {
              const bool hl = col > 0, hr = col + 4 < W;
              float4 u = make_float4(TAUV_NEG_INF, TAUV_NEG_INF, TAUV_NEG_INF, TAUV_NEG_INF), d = u;
              float ul = TAUV_NEG_INF, ur = TAUV_NEG_INF, dl = TAUV_NEG_INF, dright = TAUV_NEG_INF;
              if (o0 >= 0) {
                u = *reinterpret_cast<const float4*>(ring + o0 + col);
                if (hl) ul = ring[o0 + col - 1];
                if (hr) ur = ring[o0 + col + 4];
              }
              if (o2 >= 0) {
                d = *reinterpret_cast<const float4*>(ring + o2 + col);
                if (hl) dl = ring[o2 + col - 1];
                if (hr) dright = ring[o2 + col + 4];
              }
              const float ml = hl ? ring[o1 + col - 1] : TAUV_NEG_INF;
              const float mr = hr ? ring[o1 + col + 4] : TAUV_NEG_INF;
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
                if (peak) push(xv, flat + cc);
              }
            }
          } else {
            if (x.x >= thr_f) push(x.x, flat);
            if (x.y >= thr_f) push(x.y, flat + 1);
            if (x.z >= thr_f) push(x.z, flat + 2);
            if (x.w >= thr_f) push(x.w, flat + 3);
          }
        }
        ri += dr;
        cs += dc;
        if (cs >= S) { cs -= S; ++ri; }
      }
    } else {
      // ---- scalar path (any W / unaligned base): one element per task ----
      const int n = (rb - ra) * W;
      for (int t = tid; t < n; t += kTileThreads) {
        const int ri = t / W;
        const int col = t - ri * W;
        const int o1 = tab[ri + 1];
        const float xv = ring[o1 + col];
        if (!(xv >= thr_f)) continue;
        const uint32_t flat = plane_flat0 + (uint32_t)((ra + ri) * W + col);
        if (MODE == TAUV_TOPK_SIGMOID_PEAK) {
          float m = TAUV_NEG_INF;
          for (int dy = 0; dy < 3; ++dy) {
            const int o = tab[ri + dy];
            if (o < 0) continue;
            for (int dc2 = -1; dc2 <= 1; ++dc2) {
              const int c2 = col + dc2;
              if (c2 >= 0 && c2 < W) m = fmaxf(m, ring[o + c2]);
            }
          }
          bool peak = (xv >= m);
          if (!peak && (xv > 4.0f || m < -80.0f || (m - xv) < 1e-3f)) peak = sigmoid_tie(xv, m);
          if (peak) push(xv, flat);
        } else {
          push(xv, flat);
        }
      }
    }

    // ---- end of step: everyone is done with chunk j-1; prune if the list is getting full ----
    if (j + 1 < nchunks) fill_rowtab(j + 1);
    if (tid == 0) {
      const unsigned long long pub = (unsigned long long)pub_key << 32;
      if (pub > s_thr) s_thr = pub;
    }
    const int over = __syncthreads_or(my_end > a.soft);
    if (BULK && tid == 0 && j >= 1 && j - 1 + kStages < nchunks) issue_chunk(j - 1 + kStages);
    if (over) {
      const int n = *count_p;
      convert(n);
      const unsigned long long T =
          block_kth_largest<kTileThreads>([&](int i) { return list[i]; }, n, a.k, hist, ctl);
      // stable in-place compaction, kTileThreads entries per round (write index <= read index)
      if (tid == 0) s_base = 0;
      __syncthreads();
      for (int start = 0; start < n; start += kTileThreads) {
        const int i = start + tid;
        unsigned long long c = 0ull;
        bool keep = false;
        if (i < n) {
          c = list[i];
          keep = (c >= T) && (c != 0ull);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int lane = tid & 31, warp = tid >> 5;
        if (lane == 0) s_wsum[warp] = __popc(bal);
        __syncthreads();
        int pos = s_base + __popc(bal & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) pos += s_wsum[w];
        if (keep) list[pos] = c;
        __syncthreads();
        if (tid == 0) {
          int tot = 0;
          for (int w = 0; w < kTileThreads / 32; ++w) tot += s_wsum[w];
          s_base += tot;
        }
        __syncthreads();
      }
      if (tid == 0) {
        *count_p = s_base;
        ctl[6] = (uint32_t)s_base;  // everything kept is converted
        publish(T, s_base);
      }
      __syncthreads();
    }
  }

  // ---- emit the item's top-k ----
  __syncthreads();
  const int n = *count_p;
  convert(n);
  const unsigned long long T =
      block_kth_largest<kTileThreads>([&](int i) { return list[i]; }, n, a.k, hist, ctl);
  if (tid == 0) ctl[5] = 0;
  __syncthreads();
  unsigned long long* out = a.cand + (size_t)item * a.k;
  for (int i = tid; i < n; i += kTileThreads) {
    const unsigned long long c = list[i];
    if (c >= T && c != 0ull) out[atomicAdd(&ctl[5], 1u)] = c;
  }
  __syncthreads();
  if (tid == 0) {
    a.cand_count[item] = (int)ctl[5];
    publish(T, (int)ctl[5]);
  }
}

// ----------------------------------------------------------------------------------------------
// K2: per-frame merge (+ optional box decode)
// ----------------------------------------------------------------------------------------------
struct BoxArgs {
  int enabled;
  const float* size; long long ss[4];
  const float* offset; long long os[4];
  const float* depth; long long ds[3];
  int mode, ratio, in_h, in_w, out_h, out_w;
  float thr;
  double* yx; float* hw; float* depth_out; int* count;
};

__device__ __forceinline__ void box_one(const BoxArgs& g, int b, long long slot, int iy, int ix) {
  const float h = g.size[b * g.ss[0] + iy * g.ss[1] + ix * g.ss[2]];
  const float w = g.size[b * g.ss[0] + iy * g.ss[1] + ix * g.ss[2] + g.ss[3]];
  g.hw[slot * 2 + 0] = h;
  g.hw[slot * 2 + 1] = w;
  if (g.mode == TAUV_BOX_DECODE) {
    // decode.py:214-215: Python doubles
    const float oy = g.offset[b * g.os[0] + iy * g.os[1] + ix * g.os[2]];
    const float ox = g.offset[b * g.os[0] + iy * g.os[1] + ix * g.os[2] + g.os[3]];
    g.yx[slot * 2 + 0] = __ddiv_rn(__dadd_rn(__dmul_rn((double)g.ratio, (double)iy), (double)oy), (double)g.in_h);
    g.yx[slot * 2 + 1] = __ddiv_rn(__dadd_rn(__dmul_rn((double)g.ratio, (double)ix), (double)ox), (double)g.in_w);
  } else {
    // decode.py:87-88: int64 tensor / int -> fp32 true divide, then float()
    g.yx[slot * 2 + 0] = (double)__fdiv_rn((float)iy, (float)g.out_h);
    g.yx[slot * 2 + 1] = (double)__fdiv_rn((float)ix, (float)g.out_w);
  }
  if (g.depth != nullptr && g.depth_out != nullptr) {
    const float d = g.depth[b * g.ds[0] + iy * g.ds[1] + ix * g.ds[2]];
    float inv = __fdiv_rn(1.0f, sigmoid_ref(d));
    if (g.mode == TAUV_BOX_DECODE) inv = __fsub_rn(inv, 1.0f);  // decode.py:324
    g.depth_out[slot] = inv;
  }
}

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
  uint32_t* flags = hist;                                                                     // reused: [k]
  __shared__ uint32_t ctl[8];
  __shared__ int s_total, s_first_below, s_pool_n;
  __shared__ int s_wsum[kMergeThreads / 32];

  const unsigned long long* fc = cand + (size_t)b * items_per_frame * k;
  const int* cnt = cand_count + (size_t)b * items_per_frame;
  const int nslots = items_per_frame * k;

  if (tid == 0) {
    s_total = 0;
    s_first_below = k;
    s_pool_n = 0;
    ctl[5] = 0;
  }
  for (int i = tid; i < p2; i += kMergeThreads) sel[i] = 0ull;
  __syncthreads();
  int part = 0;
  for (int i = tid; i < items_per_frame; i += kMergeThreads) part += cnt[i];
  if (part) atomicAdd(&s_total, part);
  __syncthreads();
  const int total = s_total;

  unsigned long long T = 1ull;  // total <= k: everything valid (non-zero) is selected
  if (total <= pool_cap) {
    // usual case (items reject most of their candidates against the frame threshold): pull the frame's
    // candidates into shared memory once, one warp per item, then select there
    for (int it = warp; it < items_per_frame; it += kMergeThreads / 32) {
      const int c = cnt[it];
      if (c == 0) continue;
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_pool_n, c);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (int i = lane; i < c; i += 32) pool[base + i] = fc[(size_t)it * k + i];
    }
    __syncthreads();
    if (total > k) T = block_kth_largest<kMergeThreads>([&](int i) { return pool[i]; }, total, k, hist, ctl);
    for (int i = tid; i < total; i += kMergeThreads) {
      const unsigned long long c = pool[i];
      if (c >= T && c != 0ull) sel[atomicAdd(&ctl[5], 1u)] = c;
    }
  } else {
    // candidates live in a padded [items][k] table: slot i is valid iff (i % k) < cnt[i / k]
    auto load = [&](int i) -> unsigned long long {
      const int it = i / k;
      return (i - it * k) < cnt[it] ? fc[i] : 0ull;  // 0 never beats a real composite
    };
    if (total > k) T = block_kth_largest<kMergeThreads>(load, nslots, k, hist, ctl);
    for (int i = tid; i < nslots; i += kMergeThreads) {
      const unsigned long long c = load(i);
      if (c >= T && c != 0ull) sel[atomicAdd(&ctl[5], 1u)] = c;
    }
  }
  __syncthreads();
  const int npos = (int)ctl[5];  // = min(k, total)
  block_bitonic_sort_desc<kMergeThreads>(sel, p2);

  // ---- ranked outputs ----
  const long long hw_elems = (long long)H * W;
  for (int r = tid; r < npos; r += kMergeThreads) {
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
    for (int i = tid; i < k; i += kMergeThreads) flags[i] = 0u;
    __syncthreads();
    for (int r = tid; r < npos; r += kMergeThreads) {
      const uint32_t flat = composite_idx(sel[r]);
      if (flat < (uint32_t)k) flags[flat] = 1u;
    }
    __syncthreads();
    const int need = k - npos;
    int base = 0;
    for (int start = 0; start < k && base < need; start += kMergeThreads) {
      const int i = start + tid;
      const bool freec = (i < k) && (flags[i] == 0u);
      const unsigned bal = __ballot_sync(0xffffffffu, freec);
      if (lane == 0) s_wsum[warp] = __popc(bal);
      __syncthreads();
      int pos = base + __popc(bal & ((1u << lane) - 1u));
      int tot = 0;
      for (int w = 0; w < kMergeThreads / 32; ++w) {
        if (w < warp) pos += s_wsum[w];
        tot += s_wsum[w];
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
        if (g.enabled) {
          box_one(g, b, slot, iy, ix);
          if (0.0f < g.thr) atomicMin(&s_first_below, r);
        }
      }
      base += tot;
      __syncthreads();
    }
  }
  if (g.enabled) {
    __syncthreads();
    if (tid == 0) g.count[b] = s_first_below;
  }
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

static int plan_and_check(const float* hm, int B, int C, int H, int W, int k, void* ws, size_t ws_bytes, TopkPlan* p,
                          TileArgs* a) {
  make_plan(B, C, H, W, k, hm, p);
  TAUV_REQUIRE(ws != nullptr && (uintptr_t)ws % 256 == 0, TAUV_E_WORKSPACE, "workspace must be 256-byte aligned");
  TAUV_REQUIRE(ws_bytes >= p->cand_bytes + p->count_bytes + p->thr_bytes, TAUV_E_WORKSPACE,
               "workspace %zu < required %zu", ws_bytes, p->cand_bytes + p->count_bytes + p->thr_bytes);
  TAUV_REQUIRE(p->smem_bytes <= 227 * 1024, TAUV_E_UNSUPPORTED, "tile needs %zu B shared memory", p->smem_bytes);
  a->hm = hm; a->B = B; a->C = C; a->H = H; a->W = W; a->k = k;
  a->R = p->rows_per_chunk; a->slot_elems = p->slot_elems; a->rows_per_item = p->rows_per_item;
  a->items_per_plane = p->items_per_plane; a->cap = p->cap; a->soft = p->soft;
  a->cand = reinterpret_cast<unsigned long long*>(ws);
  a->cand_count = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(ws) + p->cand_bytes);
  a->frame_thr = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(ws) + p->cand_bytes + p->count_bytes);
  const long long items = (long long)B * p->items_per_frame;
  TAUV_REQUIRE(items < (1LL << 31), TAUV_E_UNSUPPORTED, "too many items (%lld)", items);
  return 0;
}

// stage 1: per-item candidates into the workspace
static int run_stage1(const float* hm, int B, int C, int H, int W, int k, int mode, void* ws, size_t ws_bytes,
                      cudaStream_t st) {
  TopkPlan p;
  TileArgs a;
  if (int e = plan_and_check(hm, B, C, H, W, k, ws, ws_bytes, &p, &a)) return e;
  const long long items = (long long)B * p.items_per_frame;
  void (*kern)(TileArgs) = nullptr;
  if (mode == TAUV_TOPK_SIGMOID_PEAK) kern = p.bulk ? tile_topk_kernel<1, true> : tile_topk_kernel<1, false>;
  else kern = p.bulk ? tile_topk_kernel<0, true> : tile_topk_kernel<0, false>;
  TAUV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
  TAUV_CUDA(cudaMemsetAsync(a.frame_thr, 0, p.thr_bytes, st));  // key 0 = "no threshold published yet"
  kern<<<(unsigned)items, kTileThreads, p.smem_bytes, st>>>(a);
  TAUV_LAUNCH_CHECK("tile_topk_kernel");
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
  long long pool_cap = (long long)p.items_per_frame * k;
  if (pool_cap > 12288) pool_cap = 12288;
  const size_t msmem = (size_t)p2 * 8 + (size_t)pool_cap * 8 + (size_t)(kRadixBins > k ? kRadixBins : k) * 4;
  if (mode == TAUV_TOPK_SIGMOID_PEAK) {
    TAUV_CUDA(cudaFuncSetAttribute(merge_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    merge_kernel<1><<<B, kMergeThreads, msmem, st>>>(a.cand, a.cand_count, p.items_per_frame, k, H, W, (int)pool_cap,
                                                     index, label, score, box);
  } else {
    TAUV_CUDA(cudaFuncSetAttribute(merge_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    merge_kernel<0><<<B, kMergeThreads, msmem, st>>>(a.cand, a.cand_count, p.items_per_frame, k, H, W, (int)pool_cap,
                                                     index, label, score, box);
  }
  TAUV_LAUNCH_CHECK("merge_kernel");
  return 0;
}

static int run_topk(const float* hm, int B, int C, int H, int W, int k, int mode, int64_t* index, int64_t* label,
                    float* score, const BoxArgs& box, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (int e = run_stage1(hm, B, C, H, W, k, mode, ws, ws_bytes, st)) return e;
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
  return p.cand_bytes + p.count_bytes + p.thr_bytes;
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
