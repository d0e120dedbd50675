// yolact_mask_umma.cuh — tcgen05 / TMEM mask contraction (included by yolact_mask.cu).
//
//   mask[i, pix] = sigmoid( sum_p coeff[i,p] * proto[p,pix] ) * crop(box[i])      (masks.py:8-21)
//
// as D[M = 128 pixels, N = detections (<= 256)] = A[128 x 32] * B[N x 32]^T on the 5th-gen tensor cores:
//   A = prototype tile transposed to [pixel][p], bf16, K-major, SWIZZLE_64B canonical layout (a row is 32 bf16 =
//       64 B), converted from the fp32 [P][HW] map by the producer warps (coalesced 4-byte loads along the pixels,
//       16-byte swizzled stores), 2 stages;
//   B = the frame's coefficients [detection][p], same layout, written once per frame;
//       every fp32 operand is split into a bf16 pair hi + lo and the product is formed as lo*hi + hi*lo + hi*hi
//       (three MMAs into the same accumulator), so the logits are accurate to ~1e-5 instead of bf16's ~1e-2;
//   D = fp32 accumulator in tensor memory, TMEM lane = pixel, column = detection (2 x 256 columns, so the epilogue
//       of one tile overlaps the MMA of the next).
// Pixels on the TMEM lanes make the epilogue store-friendly: after tcgen05.ld a warp holds, for one detection, 32
// consecutive pixels in its 32 lanes, i.e. one fully coalesced 128-byte store per detection and warp, straight from
// registers (the mask write is >80 % of the bytes this kernel moves).
// Persistent, warp-specialised, one CTA per SM:
//   warps 0-11  epilogue : tcgen05.ld -> sigmoid + crop -> rows staged in shared memory -> one TMA tensor store per
//                          32 detections (three groups of four warps, one warp per TMEM lane quadrant)
//   warps 12-15 producer : fp32 -> bf16 hi/lo conversion of the A tile (and of B / the crop bounds at a frame change)
//   warp  16    MMA      : one elected thread issues tcgen05.mma (K = 16 per instruction) + commits
// All hand-offs are mbarriers; tcgen05.commit arrives on them when the tensor core is done with an operand.
// The contraction depth is 32, i.e. ~12 flop per byte moved: the kernel is bound by the fp32 mask WRITE, the tensor
// pipe idles most of the time by construction.
#pragma once

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)

namespace tauv {

constexpr int kUmmaP = 32;            // contraction depth this kernel is built for
constexpr int kUmmaM = 128;           // pixels per MMA tile (TMEM lanes)
constexpr int kUmmaNMax = 256;        // detections per launch (TMEM columns per accumulator stage); more: host loops
#ifndef TAUV_MASK_EPI_WARPS
#define TAUV_MASK_EPI_WARPS 12
#endif
#ifndef TAUV_MASK_PROD_WARPS
#define TAUV_MASK_PROD_WARPS 4
#endif
// (measured with tools/mask_trace.py and tools/variants_mask.sh: 8 epilogue + 4 producer warps, 2 A stages: 1777 us per
// 64 frames, epilogue-bound at 5 us per tile; 12 + 4 warps: 1279 us, now producer-bound (latency of the prototype loads);
// 12 + 4 warps with 3 A stages and 2 staging buffers per group: 1008 us; 16 epilogue warps spill: 2756 us)
constexpr int kUmmaEpiWarps = TAUV_MASK_EPI_WARPS, kUmmaProdWarps = TAUV_MASK_PROD_WARPS;
#ifndef TAUV_MASK_PREFETCH
#define TAUV_MASK_PREFETCH 1  // tiles ahead (mask writer: 0 -> 996 us, 1 -> 965, 2 -> 1035, 4 -> 1157 per 64 frames; the reducing
                              // kernel does not care: 1, 2, 3, 6 -> 808-810 us, and holding the next tile in registers: 814)
#endif
#ifndef TAUV_MASK_A_STAGES
#define TAUV_MASK_A_STAGES 2
#endif
#ifndef TAUV_MASK_STAGE_BUFS
#define TAUV_MASK_STAGE_BUFS 2
#endif
constexpr int kUmmaGroups = kUmmaEpiWarps / 4;  // an epilogue group = four warps, one per TMEM lane quadrant
constexpr int kUmmaAStages = TAUV_MASK_A_STAGES;      // prototype tiles in flight (the producers are latency-bound)
constexpr int kUmmaStageBufs = TAUV_MASK_STAGE_BUFS;  // output staging buffers per epilogue group
constexpr int kUmmaDepthChunks = (kUmmaNMax / 32 + kUmmaGroups - 1) / kUmmaGroups;  // 32-detection chunks per group
constexpr int kUmmaThreads = (kUmmaEpiWarps + kUmmaProdWarps + 1) * 32;
// The reducing kernel (fused consumer) writes nothing and is bound by its producers: twice as many of them, each
// thread converting half of a pixel row (16 of the 32 prototype planes).
#ifndef TAUV_DEPTH_PROD_WARPS
#define TAUV_DEPTH_PROD_WARPS 8
#endif
constexpr int kUmmaDepthProdWarps = TAUV_DEPTH_PROD_WARPS;
#ifndef TAUV_MASK_RAW_STAGES
#define TAUV_MASK_RAW_STAGES 3  // (with 2 A stages: writer 895 us vs 920 with 2 raw + 3 A stages; the reducing kernel does not care)
#endif
constexpr int kUmmaRawStages = TAUV_MASK_RAW_STAGES;  // fp32 prototype tiles in flight from HBM (TMA tensor loads)
constexpr int kUmmaDepthThreads = (kUmmaEpiWarps + kUmmaDepthProdWarps + 1) * 32;
template <bool kDepth>
struct UmmaRoles {
  static constexpr int kProd = kDepth ? kUmmaDepthProdWarps : kUmmaProdWarps;
  static constexpr int kThreads = kDepth ? kUmmaDepthThreads : kUmmaThreads;
};

struct UmmaSmem {
  // each operand is kept as a bf16 pair (hi, lo) with hi + lo == the fp32 value to ~2^-17
  __align__(1024) unsigned char a[kUmmaAStages][2][kUmmaM * 64];  // [stage][hi/lo] prototype tiles, 8 KB each
  __align__(1024) unsigned char b[2][kUmmaNMax * 64];     // [hi/lo] coefficients of the frame, 2 x 16 KB
  float bounds[kUmmaNMax][4];                             // crop bounds (left, right, top, bottom) per detection
  __align__(16) float zeros[kUmmaM];                      // source of the bulk zero-fill stores
  // Output staging: [epilogue group][buffer][detection of the chunk][pixel of the tile].  A warp can only read its own
  // TMEM lane quadrant (32 pixels), and 128-byte segments scattered over ~160 masks that lie 305 KB apart run HBM at
  // 1.6 TB/s (tools/store_bench.cu); 512-byte rows reach 3.9 TB/s.  So the four quadrant warps of a group assemble
  // whole 512-byte rows here and one of them hands each row to the bulk-copy engine.
  __align__(128) float stage[kUmmaGroups][kUmmaStageBufs][32][kUmmaM];  // 16 KB per buffer (one barrier per chunk)
  // fp32 prototype tiles as the TMA engine delivers them: [stage][plane][pixel of the tile], 16 KB each (round 2: the
  // producers' plain loads were the reducing kernel's bottleneck — 32 dependent-latency loads per pixel row)
  __align__(128) float raw[kUmmaRawStages][kUmmaP][kUmmaM];
  uint64_t raw_full[kUmmaRawStages], raw_empty[kUmmaRawStages];
  uint64_t a_full[kUmmaAStages], a_empty[kUmmaAStages], acc_full[2], acc_empty[2], frame_done;
  uint32_t tmem_base;
};

// ---- PTX wrappers --------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread t of the warp gets columns [c, c+32) of TMEM lane (lane_base + t)
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// One column: thread t of the warp gets column c of TMEM lane (lane_base + t).  Asynchronous: the register is valid
// after tc_wait_ld, which names it as an operand so that the compiler cannot move a use above the wait.
__device__ __forceinline__ float tc_ld_col(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return __uint_as_float(r);
}
__device__ __forceinline__ void tc_wait_ld(float& v) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v)::"memory");
}
__device__ __forceinline__ void tc_wait_ld2(float& v0, float& v1) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v0), "+f"(v1)::"memory");
}

// Shared-memory matrix descriptor for a K-major, SWIZZLE_64B operand whose rows are 64 bytes (32 bf16):
// 8-row groups are 512 B apart (stride byte offset), start address in 16-byte units, descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_k_sw64(const void* smem, uint32_t byte_offset) {
  const uint32_t addr = smem_u32(smem) + byte_offset;
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fffu);        // start address
  d |= (uint64_t)1u << 16;                       // leading byte offset (unused for swizzled K-major; canonical 1)
  d |= (uint64_t)(512u >> 4) << 32;              // stride byte offset: 8 rows x 64 B
  d |= (uint64_t)1u << 46;                       // version
  d |= (uint64_t)4u << 61;                       // layout type: SWIZZLE_64B
  return d;
}
// Instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = n (multiple of 16, <= 256), dense.
__device__ __forceinline__ uint32_t umma_idesc_bf16_m128(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kUmmaM >> 4) << 24);
}

// Byte offset of the 16-byte chunk `c` (8 bf16: k = 8c..8c+7) of row `r` in the SWIZZLE_64B K-major layout:
// Swizzle<2,4,3>: address bits [4,6) ^= address bits [7,9), i.e. chunk ^= (r >> 1) & 3 for 64-byte rows.
__device__ __forceinline__ uint32_t sw64_offset(int r, int c) { return (uint32_t)r * 64u + (uint32_t)((c ^ ((r >> 1) & 3)) << 4); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// fp32 -> (hi, lo) bf16 pairs of two values: hi = rn_bf16(x), lo = rn_bf16(x - hi)
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(x0, x1);
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  lo = pack_bf16x2(x0 - h0, x1 - h1);
}

__device__ __forceinline__ void mask_stamp(const MaskArgs& a, long long u_rel, int col) {
  if (a.trace && blockIdx.x == 0 && u_rel < 512) {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    a.trace[u_rel * 8 + col] = t;
  }
}

// detections of frame b handled by this launch (rows [m_base, m_base + 256))
__device__ __forceinline__ int frame_rows(const MaskArgs& a, int b, int m_base) {
  const int n = (a.n_keep ? a.n_keep[b] : a.n_host) - m_base;
  return n <= 0 ? 0 : (n < kUmmaNMax ? n : kUmmaNMax);
}

template <bool kDepth>
__global__ void __launch_bounds__(UmmaRoles<kDepth>::kThreads, 1) mask_umma_kernel(const __grid_constant__ MaskArgs a, int B,
                                                                    int m_base,
                                                                    const __grid_constant__ CUtensorMap out_map,
                                                                    int use_tma,
                                                                    const __grid_constant__ CUtensorMap in_map,
                                                                    int use_tma_in) {
  constexpr int kProd = UmmaRoles<kDepth>::kProd;  // producer warps of this mode
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // (round up INSIDE the shared window: pointer arithmetic on the __shared__ array keeps the address space, an integer
  // round trip does not — the compiler then emits generic LD/ST for every shared access, which cost the epilogue ~150
  // cycles per detection)
  UmmaSmem* sm = reinterpret_cast<UmmaSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = a.H * a.W;
  const int n_tiles = (HW + kUmmaM - 1) / kUmmaM;
  const long long units = (long long)B * n_tiles;  // unit = (frame, tile of 128 pixels); contiguous share per CTA
  const long long u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;
  // Per-tile index arithmetic without divisions: every role is a single dependent instruction stream per warp, and the
  // 64-bit u / n_tiles plus the pixel -> (row, column) divisions cost ~1.3 us per tile (measured: the epilogue's
  // "work" with its whole inner loop removed).  (frame, tile) advance incrementally; pix / W is a multiply-high.
  const int b_first = (int)(u0 / n_tiles), nt_first = (int)(u0 - (long long)b_first * n_tiles);
  const unsigned w_magic = (unsigned)((0x100000000ULL + (unsigned)a.W - 1) / (unsigned)a.W);   // ceil(2^32 / W)
  const bool w_magic_ok = a.W > 1 && (unsigned long long)(HW + kUmmaM) * (unsigned)a.W < 0x100000000ULL;    // exact for pix < 2^32 / W
  auto div_w = [&](int pix) { return w_magic_ok ? (int)__umulhi((unsigned)pix, w_magic) : pix / a.W; };

  if (tid == 0) {
    for (int s = 0; s < kUmmaAStages; ++s) {
      mbar_init(&sm->a_full[s], kProd * 32);
      mbar_init(&sm->a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm->acc_full[s], 1);
      mbar_init(&sm->acc_empty[s], kUmmaEpiWarps * 32);
    }
    mbar_init(&sm->frame_done, kUmmaEpiWarps * 32);
    for (int s = 0; s < kUmmaRawStages; ++s) {
      mbar_init(&sm->raw_full[s], 1);
      mbar_init(&sm->raw_empty[s], kProd * 32);
    }
    mbar_fence_init();
  }
  if (tid < kUmmaM) sm->zeros[tid] = 0.0f;
  fence_proxy_async();  // the zeros are read by the async proxy (bulk stores)
  if (warp == kUmmaEpiWarps + kProd) {  // the MMA warp owns the tensor memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;

  if (warp < kUmmaEpiWarps) {
    // ======================================= epilogue =======================================
    const int quad = warp & 3, half = warp >> 2;  // TMEM lane quadrant / epilogue group (takes chunks half, half + groups, ...)
    // bulk stores need 16-byte aligned rows: H*W % 4 == 0 and an aligned output
    const bool bulk_zero = (HW % 4 == 0) && ((uintptr_t)a.out % 16 == 0);
    uint32_t uses[2] = {0, 0};
    uint32_t nbuf = 0;  // staged chunks so far (selects the staging buffer)
    unsigned long long dsum[kUmmaDepthChunks] = {}, dcnt[kUmmaDepthChunks] = {};  // (depth mode) lane j: detection c*32+j
    uint2 pw_next = make_uint2(0u, 0u);  // (depth mode) the next tile's pooled depth of this thread's pixel
    long long pw_u = -1;
    int as = 0, rows_frame = -1, n_rows = 0;
    int b = b_first, nt = nt_first - 1;
    for (long long u = u0; u < u1; ++u) {
      if (++nt == n_tiles) {
        nt = 0;
        ++b;
      }
      const bool frame_ends = u + 1 == u1 || nt + 1 == n_tiles;  // the last tile of this CTA's run of frame b
      if (b != rows_frame) {  // one global read per frame, not per tile
        n_rows = frame_rows(a, b, m_base);
        rows_frame = b;
      }
      if (n_rows == 0) continue;
      const int pix = nt * kUmmaM + quad * 32 + lane;
      const int yy = div_w(pix);
      const float py = (float)yy, px = (float)(pix - yy * a.W);
      float* out_tile = a.out + ((size_t)b * a.top_k + m_base) * HW + nt * kUmmaM;
      float* out_pix = out_tile + quad * 32 + lane;
      float* lg_pix = a.logits ? a.logits + ((size_t)b * a.top_k + m_base) * HW + pix : nullptr;
      // image rows the tile's pixels lie in: a mask whose box misses them is all zero on this tile
      const int tile_px = min(kUmmaM, HW - nt * kUmmaM);
      const float ty0 = (float)div_w(nt * kUmmaM), ty1 = (float)div_w(nt * kUmmaM + tile_px - 1);
      // (depth mode) this pixel's pooled camera depth: sum of the millimetre readings and number of valid readings
      // among the camera pixels whose nearest prototype pixel it is
      // (loaded one tile ahead: a load issued here would be waited for in full on every tile — 1.5 us of the 2.4 us
      // tile period when the epilogue is the bottleneck and its accumulator is already waiting)
      uint2 pw = make_uint2(0u, 0u);
      if constexpr (kDepth) {
        if (pw_u == u) pw = pw_next;
        else if (pix < HW) pw = __ldg(a.pool + (size_t)b * HW + pix);
        if (u + 1 < u1) {
          const int b2 = nt + 1 == n_tiles ? b + 1 : b;
          const int pix2 = (nt + 1 == n_tiles ? 0 : nt + 1) * kUmmaM + quad * 32 + lane;
          pw_next = pix2 < HW ? __ldg(a.pool + (size_t)b2 * HW + pix2) : make_uint2(0u, 0u);
          pw_u = u + 1;
        }
      }
      mbar_wait(&sm->acc_full[as], uses[as] & 1u);
      tc_fence_after();
      if (tid == 0) mask_stamp(a, u - u0, 5);
      if constexpr (kDepth) {
        // ---- fused consumer (yolact_node.py:131,178-183): no mask is written.  A detection's pixel is "on" when it
        // lies inside the crop box and sigmoid(logit) > 0.5, i.e. logit > 0; the warp adds up the pooled depth of its
        // 32 pixels with two integer warp reductions per detection, lane j keeps the running totals of detection
        // c*32 + j, and they go to global memory once per frame.  Boxes that miss the tile's image rows are skipped.
        // the image rows / columns this warp's 32 pixels lie in: a box that misses them selects nothing here (with the
        // probe's boxes nine (warp, detection) pairs in ten are skipped; the kernel is bound by instruction issue)
        const int wp0 = nt * kUmmaM + quad * 32, wp1 = min(wp0 + 31, HW - 1);
        const int wr0 = div_w(wp0), wr1 = div_w(wp1);
        const float wy0 = (float)wr0, wy1 = (float)wr1;
        const float wx0 = wr0 == wr1 ? (float)(wp0 - wr0 * a.W) : 0.0f, wx1 = wr0 == wr1 ? (float)(wp1 - wr1 * a.W) : (float)a.W;
#pragma unroll
        for (int ci = 0; ci < kUmmaDepthChunks; ++ci) {
          const int c = half + ci * kUmmaGroups;
          if (c * 32 < n_rows && wp0 < HW) {  // warp-uniform
            bool live = false;
            const int det = c * 32 + lane;
            if (det < n_rows) {
              const float4 bd = *reinterpret_cast<const float4*>(sm->bounds[det]);
              live = !(wy1 < bd.z || wy0 > bd.w || wx1 < bd.x || wx0 > bd.y);
            }
            const unsigned live_mask = __ballot_sync(0xffffffffu, live);
            if (live_mask != 0u) {
              // Only the live detections are touched, one TMEM column each (a dynamic column address needs no register
              // indexing; the next live column is in flight while this one is reduced).  Measured on the way here:
              // testing all 32 detections of the chunk, with a branch or straight-line, costs 2 us per tile — a warp on
              // its own issues a dependent instruction every ~5 cycles, so only the instruction COUNT matters.
              const uint32_t tbase = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * kUmmaNMax + c * 32);
              const float4* bp = reinterpret_cast<const float4*>(sm->bounds[c * 32]);
              unsigned ts = 0u, tn = 0u;  // this lane's detection, this tile (fits 32 bits: checked by the host)
              // two live detections per step: their two TMEM columns are loaded together and their four warp
              // reductions issue back to back (each step is otherwise one chain of load -> test -> reduce -> select
              // latencies); an odd one out is paired with itself
              unsigned m = live_mask;
              while (m) {
                const int j0 = __ffs(m) - 1;
                m &= m - 1u;
                const bool two = m != 0u;
                const int j1 = two ? __ffs(m) - 1 : j0;
                m &= m - 1u;  // (0 & anything = 0)
                float v0 = tc_ld_col(tbase + (uint32_t)j0), v1 = tc_ld_col(tbase + (uint32_t)j1);
                tc_wait_ld2(v0, v1);
                const float4 b0 = bp[j0], b1 = bp[j1];  // broadcasts
                const int n0 = __float_as_int(px - b0.x) | __float_as_int(b0.y - px) | __float_as_int(py - b0.z) |
                               __float_as_int(b0.w - py);
                const int n1 = __float_as_int(px - b1.x) | __float_as_int(b1.y - px) | __float_as_int(py - b1.z) |
                               __float_as_int(b1.w - py);
                const bool on0 = n0 >= 0 && v0 > 0.0f, on1 = n1 >= 0 && v1 > 0.0f;
                const unsigned sx0 = __reduce_add_sync(0xffffffffu, on0 ? pw.x : 0u);
                const unsigned sn0 = __reduce_add_sync(0xffffffffu, on0 ? pw.y : 0u);
                const unsigned sx1 = __reduce_add_sync(0xffffffffu, on1 ? pw.x : 0u);
                const unsigned sn1 = __reduce_add_sync(0xffffffffu, on1 ? pw.y : 0u);
                if (lane == j0) {
                  ts = sx0;
                  tn = sn0;
                }
                if (lane == j1) {
                  ts = sx1;
                  tn = sn1;
                }
              }
              dsum[ci] += ts;
              dcnt[ci] += tn;
            }
          }
        }
        if (tid == 0) mask_stamp(a, u - u0, 6);
        if (a.trace && blockIdx.x == 0 && lane == 0 && u - u0 < 512) {  // the slowest epilogue warp of the tile: (time << 5) | warp
          long long t;
          asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
          atomicMax(reinterpret_cast<unsigned long long*>(a.trace + (u - u0) * 8 + 7), ((unsigned long long)t << 5) | (unsigned)warp);
        }
        tc_fence_before();
        mbar_arrive(&sm->acc_empty[as]);
        ++uses[as];
        as ^= 1;
        if (frame_ends) {
          // end of this CTA's run of tiles of frame b: publish (integer sums: exact and order-independent)
#pragma unroll
          for (int ci = 0; ci < kUmmaDepthChunks; ++ci) {
            const int det = (half + ci * kUmmaGroups) * 32 + lane;
            if (det < n_rows && (dsum[ci] | dcnt[ci])) {
              unsigned long long* dst = a.acc + ((size_t)b * a.top_k + m_base + det) * 2;
              atomicAdd(dst, dsum[ci]);
              atomicAdd(dst + 1, dcnt[ci]);
            }
            dsum[ci] = 0ull;
            dcnt[ci] = 0ull;
          }
          mbar_arrive(&sm->frame_done);
        }
        continue;
      }
      if constexpr (!kDepth) {
#pragma unroll 1
      for (int c = half; c * 32 < n_rows; c += kUmmaGroups) {
        // lane j looks at detection c*32+j: does its box reach the tile at all?  If not, one bulk store of zeros
        // (shared -> global, issued by a single lane of the quadrant-0 warp) replaces 4 warps x 1 store + tests.
        bool live = false;
        {
          const int det = c * 32 + lane;
          if (det < n_rows) {
            const float4 bd = *reinterpret_cast<const float4*>(sm->bounds[det]);
            live = lg_pix != nullptr || !bulk_zero || !(ty1 < bd.z || ty0 > bd.w);
            // (TMA mode writes whole 32-row boxes: rows the box misses come out as zeros from the crop arithmetic)
            if (!live && quad == 0 && !use_tma) bulk_s2g(out_tile + (size_t)det * HW, sm->zeros, (uint32_t)tile_px * 4u);
          }
        }
        const unsigned live_mask = __ballot_sync(0xffffffffu, live);
        // (uniform over the group's four warps: `live` does not depend on the quadrant)
        if (live_mask == 0u && !use_tma) continue;
        float v[32];
        tc_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * kUmmaNMax + c * 32), v);
        float* op = out_pix + (size_t)(c * 32) * HW;
        float* lp = lg_pix ? lg_pix + (size_t)(c * 32) * HW : nullptr;
        const float4* bp = reinterpret_cast<const float4*>(sm->bounds[c * 32]);
        if (bulk_zero) {
          // ---- staged path: rows of 128 pixels assembled in shared memory, one bulk store per (detection, tile)
          float (*stg)[kUmmaM] = sm->stage[half][nbuf % (uint32_t)kUmmaStageBufs];
          // Branch-free AND predicate-free on purpose.  With a vote and a branch per detection — or even just the
          // chained FSETPs of the crop test and the range fix-ups of __expf / __fdividef, which all funnel through one
          // predicate register — the 32 iterations ran strictly one after the other (~100-160 cycles each: 5-8 us of
          // epilogue per tile against 2 us for everything else).  Here the crop test is sign-bit arithmetic
          // (inclusive bounds; the "no crop" bounds are +-inf, which subtract to +inf) and the sigmoid is
          // rcp.approx(1 + ex2.approx(-x*log2 e)) (2 ulp each; saturates correctly to 0 / 1), so eight detections at a
          // time are independent straight-line code.  Rows that are not live are computed too and never stored.
#pragma unroll
          for (int j0 = 0; j0 < 32; j0 += 8) {
            if (((live_mask >> j0) & 0xffu) == 0u) {  // warp-uniform: none of these eight boxes reaches the tile
#pragma unroll
              for (int i = 0; i < 8; ++i) stg[j0 + i][quad * 32 + lane] = 0.0f;
              continue;
            }
            int keepm[8];
            float sg[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bd = bp[j0 + i];  // same address for the whole warp: a broadcast
              const int neg = __float_as_int(px - bd.x) | __float_as_int(bd.y - px) | __float_as_int(py - bd.z) |
                              __float_as_int(bd.w - py);
              keepm[i] = ~(neg >> 31);  // all ones when the pixel is inside the box
            }
            if (a.precise) {  // (uniform over the launch)
              // consumers that interpolate the mask values and threshold the result (tauv_yolact_mask_binary, bilinear)
              // need fp32-class values: rcp.approx(1 + ex2.approx(-x log2 e)), 2 ulp each, saturating to 0 / 1
#pragma unroll
              for (int i = 0; i < 8; ++i) sg[i] = __fdividef(1.0f, 1.0f + exp2f(v[j0 + i] * -1.4426950408889634f));
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                // sigmoid(x) = 0.5 + 0.5*tanh(x/2): one MUFU op (tanh.approx, |error| <= ~5e-4 on the mask value
                // against the 2.5e-3 the bf16 contraction is allowed) instead of ex2 + rcp
                float t;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(v[j0 + i] * 0.5f));
                sg[i] = fmaf(t, 0.5f, 0.5f);
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) stg[j0 + i][quad * 32 + lane] = __int_as_float(__float_as_int(sg[i]) & keepm[i]);
          }
          if (lp && pix < HW) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if ((live_mask >> j) & 1u) lp[(size_t)j * HW] = v[j];
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the bulk-copy engine
          // (before the barrier: the stores issued from the buffer the NEXT chunk will fill have read it)
          if (quad == 0) bulk_wait_read<(kUmmaStageBufs >= 2 ? kUmmaStageBufs - 2 : 0)>();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
          if (quad == 0) {
            if (use_tma) {
              // one tensor store for the whole chunk: box {128 pixels, 32 detections, 1 frame} of out[B][top_k][HW];
              // the parts of the box beyond HW or top_k are clipped by the copy engine
              if (lane == 0) {
                asm volatile(
                    "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&out_map),
                    "r"(nt * kUmmaM), "r"(m_base + c * 32), "r"(b), "r"(smem_u32(&stg[0][0]))
                    : "memory");
              }
            } else if ((live_mask >> lane) & 1u) {
              bulk_s2g(out_tile + (size_t)(c * 32 + lane) * HW, stg[lane], (uint32_t)tile_px * 4u);
            }
            bulk_commit();
          }
          ++nbuf;
          continue;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j, op += HW) {
          if ((live_mask >> j) & 1u) {  // warp-uniform
            const float4 bd = bp[j];  // same address for the whole warp: a broadcast
            const bool inside = px >= bd.x && px <= bd.y && py >= bd.z && py <= bd.w;
            float val = 0.0f;
            if (__any_sync(0xffffffffu, inside)) val = inside ? __fdividef(1.0f, 1.0f + __expf(-v[j])) : 0.0f;
            if (pix < HW) {
              *op = val;  // 32 lanes = 32 consecutive pixels of one mask: one 128-byte store
              if (lp) lp[(size_t)j * HW] = v[j];
            }
          }
        }
      }
      if (tid == 0) mask_stamp(a, u - u0, 6);
      tc_fence_before();
      mbar_arrive(&sm->acc_empty[as]);
      ++uses[as];
      as ^= 1;
      // end of this CTA's run of units of frame b: the producers may overwrite B / the crop bounds
      if (frame_ends) mbar_arrive(&sm->frame_done);
      }  // (!kDepth)
    }
    bulk_commit();
    bulk_wait<0>();  // the zero-fill stores this thread issued have completed
  } else if (warp < kUmmaEpiWarps + kProd) {
    // ======================================= producers =======================================
    const int pt = tid - kUmmaEpiWarps * 32;  // producer thread index: the pixel rows of the A tile this thread converts
    uint32_t fills[kUmmaAStages] = {}, frames = 0;
    int st = 0, cur_frame = -1, rows_frame = -1, n_rows = 0;
    int b = b_first, nt = nt_first - 1;
    // TMA loads of the fp32 prototype tiles (producer thread 0 issues them, kUmmaRawStages tiles ahead of the
    // conversion): the load cursor walks the same sequence of tiles — frames without detections skipped — as the loop
    uint32_t raw_fills[kUmmaRawStages] = {}, raw_loads[kUmmaRawStages] = {};
    int rs = 0, ls = 0, lb = b_first, lnt = nt_first - 1, l_rows_frame = -1, l_rows = 0;
    long long lu = u0;
    auto issue_next_load = [&]() {  // producer thread 0 only
      for (; lu < u1; ++lu) {
        if (++lnt == n_tiles) {
          lnt = 0;
          ++lb;
        }
        if (lb != l_rows_frame) {
          l_rows = frame_rows(a, lb, m_base);
          l_rows_frame = lb;
        }
        if (l_rows == 0) continue;
        if (raw_loads[ls] > 0) mbar_wait(&sm->raw_empty[ls], (raw_loads[ls] - 1) & 1u);  // every producer is done with it
        mbar_expect_tx(&sm->raw_full[ls], (uint32_t)sizeof(sm->raw[0]));
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                smem_u32(&sm->raw[ls][0][0])),
            "l"(&in_map), "r"(lnt * kUmmaM), "r"(lb * kUmmaP), "r"(smem_u32(&sm->raw_full[ls]))
            : "memory");
        ++raw_loads[ls];
        ls = ls + 1 == kUmmaRawStages ? 0 : ls + 1;
        ++lu;
        return;
      }
    };
    if (use_tma_in && pt == 0)
      for (int i = 0; i < kUmmaRawStages; ++i) issue_next_load();
    for (long long u = u0; u < u1; ++u) {
      if (++nt == n_tiles) {
        nt = 0;
        ++b;
      }
      if (b != rows_frame) {
        n_rows = frame_rows(a, b, m_base);
        rows_frame = b;
      }
      if (n_rows == 0) continue;
      if (b != cur_frame) {
        // B (coefficients) and the crop bounds of the frame.  The epilogue must be done with the previous frame
        // (which also means every MMA that read the old B has completed).
        if (frames > 0) mbar_wait(&sm->frame_done, (frames - 1) & 1u);
        const int n_pad = (n_rows + 15) & ~15;
        for (int i = pt; i < n_pad * 4; i += kProd * 32) {
          const int row = i >> 2, c = i & 3;  // 16-byte chunk c of detection `row`
          uint4 qh = make_uint4(0, 0, 0, 0), ql = qh;
          if (row < n_rows) {
            const size_t src_row = a.keep ? ((size_t)b * a.N + (size_t)a.keep[(size_t)b * a.top_k + m_base + row])
                                          : (size_t)(m_base + row);
            const float4 f0 = *reinterpret_cast<const float4*>(a.coeff + src_row * kUmmaP + c * 8);
            const float4 f1 = *reinterpret_cast<const float4*>(a.coeff + src_row * kUmmaP + c * 8 + 4);
            split_bf16x2(f0.x, f0.y, qh.x, ql.x);
            split_bf16x2(f0.z, f0.w, qh.y, ql.y);
            split_bf16x2(f1.x, f1.y, qh.z, ql.z);
            split_bf16x2(f1.z, f1.w, qh.w, ql.w);
          }
          *reinterpret_cast<uint4*>(sm->b[0] + sw64_offset(row, c)) = qh;
          *reinterpret_cast<uint4*>(sm->b[1] + sw64_offset(row, c)) = ql;
        }
        for (int row = pt; row < ((n_rows + 31) & ~31); row += kProd * 32) {
          float4 bd = make_float4(TAUV_NEG_INF, -TAUV_NEG_INF, TAUV_NEG_INF, -TAUV_NEG_INF);  // no crop
          if (row >= n_rows) {
            bd = make_float4(-TAUV_NEG_INF, TAUV_NEG_INF, -TAUV_NEG_INF, TAUV_NEG_INF);  // padding row: empty box -> zeros
          } else if (a.box) {
            const CropBounds cbd = crop_bounds(a.box[(size_t)b * a.top_k + m_base + row], a.H, a.W);
            bd = make_float4(cbd.left, cbd.right, cbd.top, cbd.bottom);
          }
          *reinterpret_cast<float4*>(sm->bounds[row]) = bd;
        }
        cur_frame = b;
        ++frames;
      }
      if (!use_tma_in && TAUV_MASK_PREFETCH > 0 && u + TAUV_MASK_PREFETCH < u1) {
        // The producers' cost is the latency of their 32 loads per pixel row (2.8 us per tile straight from HBM under
        // the write stream).  Holding the next tile in registers spills (the kernel is capped at 96 registers by its
        // 17 warps) and a spill waits for the load; an L2 prefetch of a later tile costs two instructions per thread:
        // 32 planes x 512 bytes = 128 lines, one (+ the straddled one: plane bases are not 128-byte aligned) each.
        int b2 = b, nt2 = nt + TAUV_MASK_PREFETCH;
        while (nt2 >= n_tiles) {
          nt2 -= n_tiles;
          ++b2;
        }
        for (int i = pt; i < kUmmaP * 4; i += kProd * 32) {
          const int p = i >> 2, q = i & 3;
          const int pix = nt2 * kUmmaM + q * 32;
          if (pix < HW) {
            const float* src = a.proto + ((size_t)b2 * kUmmaP + p) * HW + pix;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
            if (pix + 31 < HW) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + 31));
          }
        }
      }
      {
        if (fills[st] > 0) mbar_wait(&sm->a_empty[st], (fills[st] - 1) & 1u);
        if (pt == 0) mask_stamp(a, u - u0, 0);
        // A tile: pixel rows pt, pt + producers, ...; 32 prototype values each -> 4 chunks of 8 bf16 (hi and lo)
        if (use_tma_in) {
          // the fp32 tile is in shared memory ([plane][pixel]: consecutive lanes read consecutive words)
          mbar_wait(&sm->raw_full[rs], raw_fills[rs] & 1u);
          constexpr int kSplit = kProd * 32 / kUmmaM;             // producer threads per pixel row (1 or 2)
          constexpr int kPlanes = kUmmaP / (kSplit > 0 ? kSplit : 1);
          if constexpr (kSplit >= 1) {
            const int pr = pt & (kUmmaM - 1), h = pt >> 7;
            float f[kPlanes];
#pragma unroll
            for (int p = 0; p < kPlanes; ++p) f[p] = sm->raw[rs][kPlanes * h + p][pr];
#pragma unroll
            for (int c = 0; c < kPlanes / 8; ++c) {
              uint4 qh, ql;
              split_bf16x2(f[8 * c], f[8 * c + 1], qh.x, ql.x);
              split_bf16x2(f[8 * c + 2], f[8 * c + 3], qh.y, ql.y);
              split_bf16x2(f[8 * c + 4], f[8 * c + 5], qh.z, ql.z);
              split_bf16x2(f[8 * c + 6], f[8 * c + 7], qh.w, ql.w);
              *reinterpret_cast<uint4*>(sm->a[st][0] + sw64_offset(pr, (kPlanes / 8) * h + c)) = qh;
              *reinterpret_cast<uint4*>(sm->a[st][1] + sw64_offset(pr, (kPlanes / 8) * h + c)) = ql;
            }
          }
          mbar_arrive(&sm->raw_empty[rs]);
          ++raw_fills[rs];
          rs = rs + 1 == kUmmaRawStages ? 0 : rs + 1;
          if (pt == 0) issue_next_load();
        } else if constexpr (kProd * 32 == 2 * kUmmaM) {
          // two threads per pixel row: planes [16h, 16h + 16) -> chunks 2h, 2h + 1
          const int pr = pt & (kUmmaM - 1), h = pt >> 7;
          const int pix = nt * kUmmaM + pr;
          const float* src = a.proto + ((size_t)b * kUmmaP + 16 * h) * HW + pix;
          float f[16];
#pragma unroll
          for (int p = 0; p < 16; ++p) f[p] = pix < HW ? __ldg(src + (size_t)p * HW) : 0.0f;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint4 qh, ql;
            split_bf16x2(f[8 * c], f[8 * c + 1], qh.x, ql.x);
            split_bf16x2(f[8 * c + 2], f[8 * c + 3], qh.y, ql.y);
            split_bf16x2(f[8 * c + 4], f[8 * c + 5], qh.z, ql.z);
            split_bf16x2(f[8 * c + 6], f[8 * c + 7], qh.w, ql.w);
            *reinterpret_cast<uint4*>(sm->a[st][0] + sw64_offset(pr, 2 * h + c)) = qh;
            *reinterpret_cast<uint4*>(sm->a[st][1] + sw64_offset(pr, 2 * h + c)) = ql;
          }
        } else
#pragma unroll 1
        for (int pr = pt; pr < kUmmaM; pr += kProd * 32) {
          const int pix = nt * kUmmaM + pr;
          const float* src = a.proto + (size_t)b * kUmmaP * HW + pix;
          float f[kUmmaP];
#pragma unroll
          for (int p = 0; p < kUmmaP; ++p) f[p] = pix < HW ? __ldg(src + (size_t)p * HW) : 0.0f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 qh, ql;
            split_bf16x2(f[8 * c], f[8 * c + 1], qh.x, ql.x);
            split_bf16x2(f[8 * c + 2], f[8 * c + 3], qh.y, ql.y);
            split_bf16x2(f[8 * c + 4], f[8 * c + 5], qh.z, ql.z);
            split_bf16x2(f[8 * c + 6], f[8 * c + 7], qh.w, ql.w);
            *reinterpret_cast<uint4*>(sm->a[st][0] + sw64_offset(pr, c)) = qh;
            *reinterpret_cast<uint4*>(sm->a[st][1] + sw64_offset(pr, c)) = ql;
          }
        }
      }
      fence_proxy_async();  // generic-proxy writes (A, and B / bounds at a frame change) -> visible to the async proxy
      if (pt == 0) mask_stamp(a, u - u0, 1);
      mbar_arrive(&sm->a_full[st]);
      ++fills[st];
      st = st + 1 == kUmmaAStages ? 0 : st + 1;
    }
  } else if (lane == 0) {
    // ======================================= MMA issuer =======================================
    uint32_t fills[kUmmaAStages] = {}, uses[2] = {0, 0};
    int st = 0, as = 0, rows_frame = -1, n_rows = 0;
    int b = b_first, nt = nt_first - 1;
    for (long long u = u0; u < u1; ++u) {
      if (++nt == n_tiles) {
        nt = 0;
        ++b;
      }
      if (b != rows_frame) {
        n_rows = frame_rows(a, b, m_base);
        rows_frame = b;
      }
      if (n_rows == 0) continue;
      const uint32_t idesc = umma_idesc_bf16_m128((n_rows + 15) & ~15);
      mbar_wait(&sm->a_full[st], fills[st] & 1u);
      mask_stamp(a, u - u0, 2);
      if (uses[as] > 0) mbar_wait(&sm->acc_empty[as], (uses[as] - 1) & 1u);
      tc_fence_after();
      mask_stamp(a, u - u0, 3);
      const uint32_t d = tmem + (uint32_t)(as * kUmmaNMax);
#pragma unroll
      for (int k = 0; k < kUmmaP / 16; ++k) {  // K = 16 bf16 = 32 bytes per instruction, inside the 64-byte rows
        const uint64_t ah = umma_desc_k_sw64(sm->a[st][0], k * 32), al = umma_desc_k_sw64(sm->a[st][1], k * 32);
        const uint64_t bh = umma_desc_k_sw64(sm->b[0], k * 32), bl = umma_desc_k_sw64(sm->b[1], k * 32);
        tc_mma_bf16(d, al, bh, idesc, k > 0);  // small terms first
        tc_mma_bf16(d, ah, bl, idesc, 1u);
        tc_mma_bf16(d, ah, bh, idesc, 1u);
      }
      tc_commit(&sm->acc_full[as]);  // arrives when the MMAs above are complete ...
      tc_commit(&sm->a_empty[st]);   // ... and so does this one: the A stage may be refilled
      mask_stamp(a, u - u0, 4);
      ++uses[as];
      as ^= 1;
      ++fills[st];
      st = st + 1 == kUmmaAStages ? 0 : st + 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kUmmaEpiWarps + kProd) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

static bool umma_shape_ok(const MaskArgs& a) {
  return a.P == kUmmaP && (uintptr_t)a.coeff % 16 == 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// out[B][rows_per_frame][HW] fp32 as a 3-D tensor map with box {128, 32, 1}; false when the layout does not qualify
// (HW*4 not a multiple of 16, unaligned base) or the driver entry point is unavailable: the kernel then stores rows
// one by one.
static bool make_out_map(const MaskArgs& a, int B, CUtensorMap* map) {
  const long long HW = (long long)a.H * a.W;
  if (HW % 4 != 0 || (uintptr_t)a.out % 16 != 0 || a.logits != nullptr) return false;
  static EncodeTiledFn encode = nullptr;  // (idempotent lookup; benign race)
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
      (void)cudaGetLastError();
      return false;
    }
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)a.top_k, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)HW * 4, (cuuint64_t)a.top_k * (cuuint64_t)HW * 4};  // bytes, dims 1 and 2
  const cuuint32_t box[3] = {(cuuint32_t)kUmmaM, 32u, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, a.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// proto[B][P][HW] fp32 as a 2-D tensor map {HW, B*P} with box {128 pixels, 32 planes} (pixels beyond HW read as zero)
static bool make_in_map(const MaskArgs& a, int B, CUtensorMap* map) {
  const long long HW = (long long)a.H * a.W;
  if (HW % 4 != 0 || (uintptr_t)a.proto % 16 != 0 || a.P != kUmmaP) return false;
  static EncodeTiledFn encode = nullptr;  // (idempotent lookup; benign race)
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
      (void)cudaGetLastError();
      return false;
    }
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)HW, (cuuint64_t)B * kUmmaP};
  const cuuint64_t strides[1] = {(cuuint64_t)HW * 4};  // bytes, dim 1
  const cuuint32_t box[2] = {(cuuint32_t)kUmmaM, (cuuint32_t)kUmmaP};
  const cuuint32_t estr[2] = {1u, 1u};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a.proto), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int launch_mask_umma(const MaskArgs& a, int B, int max_rows, cudaStream_t st) {
  const size_t smem = sizeof(UmmaSmem) + 1024;
  const bool depth = a.acc != nullptr;
  TAUV_CUDA(ensure_dynamic_smem((const void*)(depth ? mask_umma_kernel<true> : mask_umma_kernel<false>), smem));
  const int HW = a.H * a.W;
  const long long units = (long long)B * ((HW + kUmmaM - 1) / kUmmaM);
  long long grid = num_sms();
  if (grid > units) grid = units;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const int use_tma = !depth && !debug_env("TAUV_MASK_NO_TMA") && make_out_map(a, B, &map) ? 1 : 0;
  CUtensorMap in_map;
  memset(&in_map, 0, sizeof(in_map));
  const int use_tma_in = !debug_env("TAUV_MASK_NO_TMA_IN") && make_in_map(a, B, &in_map) ? 1 : 0;
  for (int m_base = 0; m_base < max_rows; m_base += kUmmaNMax) {
    if (depth) mask_umma_kernel<true><<<(unsigned)grid, kUmmaDepthThreads, smem, st>>>(a, B, m_base, map, 0, in_map, use_tma_in);
    else mask_umma_kernel<false><<<(unsigned)grid, kUmmaThreads, smem, st>>>(a, B, m_base, map, use_tma, in_map, use_tma_in);
    TAUV_LAUNCH_CHECK("mask_umma_kernel");
  }
  return 0;
}

}  // namespace tauv
