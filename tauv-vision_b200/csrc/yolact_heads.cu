// yolact_heads.cu — the prediction heads' outputs in the layout every consumer here reads.
//
// Replaces (reference file:line under src/tauv_vision/yolact/model/):
//   prediction_head.py:111-113  classification.permute(0, 2, 3, 1).reshape(B, -1, n_classes + 1)
//   prediction_head.py:122-124  box_encoding.permute(0, 2, 3, 1).reshape(B, -1, 4)
//   prediction_head.py:137-140  mask_coeff.permute(0, 2, 3, 1).reshape(B, -1, n_prototype_masks), then tanh
//   model.py:55-58              torch.cat over the FPN levels along the prior axis
// The reference makes a transposed copy per level (the reshape of a permuted view), a tanh pass, and a concatenated copy;
// here each level's NCHW convolution output [B, CH, H_l, W_l] (CH = A * C) is transposed straight into its slice of the
// final [B, sum_l A H_l W_l, C] tensor — one read, one write — with tanh applied on the way.  The backward is the
// inverse transposition (times 1 - y^2 for tanh) into per-level NCHW gradients.
#include "common.cuh"

namespace tauv {

constexpr int kHeadLevels = 8;
constexpr int kHeadTile = 32;

struct HeadPackArgs {
  const float* lvl_in[kHeadLevels];   // forward: level inputs [B,CH,HW_l]
  float* lvl_out[kHeadLevels];        // backward: level gradients [B,CH,HW_l]
  int hw[kHeadLevels];                // H_l * W_l
  int row_base[kHeadLevels];          // sum of hw over the previous levels
  int tile_base[kHeadLevels + 1];     // prefix of the levels' tile counts
  int n_levels, CH, rows;             // rows = sum of hw
  int tanh_act;
  float* packed;                      // forward: output [B, rows * CH]
  const float* grad_packed;           // backward: incoming gradient [B, rows * CH]
  const float* y_packed;              // backward with tanh: the forward output
};

template <bool BACKWARD>
__global__ void __launch_bounds__(kHeadTile * 8) head_pack_kernel(const HeadPackArgs a) {
  __shared__ float tile[kHeadTile][kHeadTile + 1];
  int l = 0;
  while (l + 1 < a.n_levels && (int)blockIdx.x >= a.tile_base[l + 1]) ++l;
  const int hw = a.hw[l];
  const int t = blockIdx.x - a.tile_base[l];
  const int tiles_hw = (hw + kHeadTile - 1) / kHeadTile;
  const int hw0 = (t % tiles_hw) * kHeadTile, ch0 = (t / tiles_hw) * kHeadTile;
  const int b = blockIdx.y, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const size_t lvl_off = (size_t)b * a.CH * hw;
  const size_t packed_off = (size_t)b * a.rows * a.CH + (size_t)a.row_base[l] * a.CH;
  if (!BACKWARD) {
    const float* in = a.lvl_in[l] + lvl_off;
#pragma unroll
    for (int i = 0; i < kHeadTile; i += 8) {   // rows of the tile = channels, 128 contiguous bytes along hw
      const int ch = ch0 + ty + i, p = hw0 + tx;
      if (ch < a.CH && p < hw) tile[ty + i][tx] = in[(size_t)ch * hw + p];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kHeadTile; i += 8) {   // written with the channels contiguous
      const int p = hw0 + ty + i, ch = ch0 + tx;
      if (p < hw && ch < a.CH) {
        const float v = tile[tx][ty + i];
        a.packed[packed_off + (size_t)p * a.CH + ch] = a.tanh_act ? tanhf(v) : v;
      }
    }
  } else {
    float* out = a.lvl_out[l] + lvl_off;
#pragma unroll
    for (int i = 0; i < kHeadTile; i += 8) {
      const int p = hw0 + ty + i, ch = ch0 + tx;
      if (p < hw && ch < a.CH) {
        const size_t at = packed_off + (size_t)p * a.CH + ch;
        float g = a.grad_packed[at];
        if (a.tanh_act) {
          const float y = a.y_packed[at];
          g = g * (1.0f - y * y);
        }
        tile[ty + i][tx] = g;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kHeadTile; i += 8) {
      const int ch = ch0 + ty + i, p = hw0 + tx;
      if (ch < a.CH && p < hw) out[(size_t)ch * hw + p] = tile[tx][ty + i];
    }
  }
}

// The same with whole output rows per CTA (CH * (PT + 1) floats of shared memory): a CTA takes PT consecutive cells of a
// level and ALL their CH channels, so what it writes (forward) or reads (backward) in the packed tensor is one contiguous
// block of PT * CH floats — full sectors, whatever CH is (243 floats per row are never 128-byte aligned: 32 x 32 tiles
// touch five sectors per 128-byte store and waste the partial channel tile).
template <bool BACKWARD>
__global__ void __launch_bounds__(kHeadTile * 8) head_pack_rows_kernel(const HeadPackArgs a, int PT) {
  extern __shared__ float s_rows[];  // [CH][PT + 1]
  int l = 0;
  while (l + 1 < a.n_levels && (int)blockIdx.x >= a.tile_base[l + 1]) ++l;
  const int hw = a.hw[l];
  const int hw0 = (blockIdx.x - a.tile_base[l]) * PT;
  const int npix = min(PT, hw - hw0);
  const int b = blockIdx.y, tid = threadIdx.x, nthr = kHeadTile * 8;
  const int S = PT + 1;
  const size_t lvl_off = (size_t)b * a.CH * hw + hw0;
  const size_t packed_off = (size_t)b * a.rows * a.CH + (size_t)(a.row_base[l] + hw0) * a.CH;
  const int n_el = npix * a.CH;
  const int dch = nthr % a.CH, dp = nthr / a.CH;   // a step of nthr elements of the packed block, as (cell, channel)
  if (!BACKWARD) {
    const float* in = a.lvl_in[l] + lvl_off;
    for (int idx = tid; idx < a.CH * PT; idx += nthr) {   // (PT is a multiple of 32: a warp reads one channel's cells)
      const int ch = idx / PT, p = idx - ch * PT;
      if (p < npix) s_rows[ch * S + p] = in[(size_t)ch * hw + p];
    }
    __syncthreads();
    int p = tid / a.CH, ch = tid - p * a.CH;
    for (int e = tid; e < n_el; e += nthr) {
      const float v = s_rows[ch * S + p];
      a.packed[packed_off + e] = a.tanh_act ? tanhf(v) : v;
      ch += dch;
      p += dp;
      if (ch >= a.CH) { ch -= a.CH; ++p; }
    }
  } else {
    int p = tid / a.CH, ch = tid - p * a.CH;
    for (int e = tid; e < n_el; e += nthr) {
      float g = a.grad_packed[packed_off + e];
      if (a.tanh_act) {
        const float y = a.y_packed[packed_off + e];
        g = g * (1.0f - y * y);
      }
      s_rows[ch * S + p] = g;
      ch += dch;
      p += dp;
      if (ch >= a.CH) { ch -= a.CH; ++p; }
    }
    __syncthreads();
    float* out = a.lvl_out[l] + lvl_off;
    for (int idx = tid; idx < a.CH * PT; idx += nthr) {
      const int ch2 = idx / PT, p2 = idx - ch2 * PT;
      if (p2 < npix) out[(size_t)ch2 * hw + p2] = s_rows[ch2 * S + p2];
    }
  }
}

// Measured at B = 64 (tools/heads_once.py): whole rows win for narrow heads (A*C = 12: 23.2 -> 18.1 us — the 32 x 32 tiles
// are 5/8 empty there) and lose for wide ones (243: 179 -> 296 us, 96: 90 -> 132 us: more index arithmetic per element
// than the tile kernel, which is what bounds both), so only heads narrower than a tile take them.
static int head_rows_pt(int CH) { return CH < kHeadTile ? 256 : 0; }

static int head_pack_plan(HeadPackArgs* a, const int* hw, int n_levels, int B, int CH, int PT) {
  TAUV_REQUIRE(hw, TAUV_E_NULL, "pointers must not be NULL");
  TAUV_REQUIRE(n_levels > 0 && n_levels <= kHeadLevels, TAUV_E_UNSUPPORTED, "n_levels=%d outside 1..%d", n_levels, kHeadLevels);
  TAUV_REQUIRE(B > 0 && B <= 65535 && CH > 0, TAUV_E_SHAPE, "bad shape B=%d CH=%d", B, CH);
  long long rows = 0, tiles = 0;
  for (int l = 0; l < n_levels; ++l) {
    TAUV_REQUIRE(hw[l] > 0, TAUV_E_SHAPE, "level %d has %d cells", l, hw[l]);
    a->hw[l] = hw[l];
    a->row_base[l] = (int)rows;
    a->tile_base[l] = (int)tiles;
    rows += hw[l];
    tiles += PT ? (long long)((hw[l] + PT - 1) / PT)
                : (long long)((hw[l] + kHeadTile - 1) / kHeadTile) * ((CH + kHeadTile - 1) / kHeadTile);
  }
  TAUV_REQUIRE(rows * CH < (1LL << 31) && tiles < (1LL << 31), TAUV_E_UNSUPPORTED, "level sizes too large");
  a->tile_base[n_levels] = (int)tiles;
  a->n_levels = n_levels;
  a->CH = CH;
  a->rows = (int)rows;
  return 0;
}

}  // namespace tauv

using namespace tauv;

extern "C" int tauv_yolact_pack_heads(const float* const* levels, const int* hw, int n_levels, int B, int CH,
                                      int tanh_act, float* packed, tauv_stream_t stream) {
  TAUV_REQUIRE(levels && packed, TAUV_E_NULL, "pointers must not be NULL");
  HeadPackArgs a{};
  const int PT = CH > 0 ? head_rows_pt(CH) : 0;
  if (int rc = head_pack_plan(&a, hw, n_levels, B, CH, PT)) return rc;
  for (int l = 0; l < n_levels; ++l) {
    TAUV_REQUIRE(levels[l], TAUV_E_NULL, "level %d is NULL", l);
    a.lvl_in[l] = levels[l];
  }
  a.tanh_act = tanh_act;
  a.packed = packed;
  if (PT) {
    const size_t smem = (size_t)CH * (PT + 1) * sizeof(float);
    if (smem > 48 * 1024) TAUV_CUDA(ensure_dynamic_smem((const void*)head_pack_rows_kernel<false>, smem));
    head_pack_rows_kernel<false><<<dim3(a.tile_base[n_levels], B), kHeadTile * 8, smem, (cudaStream_t)stream>>>(a, PT);
  } else {
    head_pack_kernel<false><<<dim3(a.tile_base[n_levels], B), kHeadTile * 8, 0, (cudaStream_t)stream>>>(a);
  }
  TAUV_LAUNCH_CHECK("head_pack_kernel<forward>");
  return 0;
}

extern "C" int tauv_yolact_pack_heads_backward(const float* grad_packed, const float* y_packed, const int* hw,
                                               int n_levels, int B, int CH, int tanh_act, float* const* grad_levels,
                                               tauv_stream_t stream) {
  TAUV_REQUIRE(grad_packed && grad_levels && (!tanh_act || y_packed), TAUV_E_NULL, "pointers must not be NULL");
  HeadPackArgs a{};
  const int PT = CH > 0 ? head_rows_pt(CH) : 0;
  if (int rc = head_pack_plan(&a, hw, n_levels, B, CH, PT)) return rc;
  for (int l = 0; l < n_levels; ++l) {
    TAUV_REQUIRE(grad_levels[l], TAUV_E_NULL, "level %d is NULL", l);
    a.lvl_out[l] = grad_levels[l];
  }
  a.tanh_act = tanh_act;
  a.grad_packed = grad_packed;
  a.y_packed = y_packed;
  if (PT) {
    const size_t smem = (size_t)CH * (PT + 1) * sizeof(float);
    if (smem > 48 * 1024) TAUV_CUDA(ensure_dynamic_smem((const void*)head_pack_rows_kernel<true>, smem));
    head_pack_rows_kernel<true><<<dim3(a.tile_base[n_levels], B), kHeadTile * 8, smem, (cudaStream_t)stream>>>(a, PT);
  } else {
    head_pack_kernel<true><<<dim3(a.tile_base[n_levels], B), kHeadTile * 8, 0, (cudaStream_t)stream>>>(a);
  }
  TAUV_LAUNCH_CHECK("head_pack_kernel<backward>");
  return 0;
}
