// yolact_common.cuh — box arithmetic shared by the YOLACT kernels (exact reference operation order).
#pragma once
#include "common.cuh"

namespace tauv {

struct Corners {
  float y0, x0, y1, x1, area;
};

// boxes.py:15-27 (corners = centre -+ size/2) and :77-78 (area from h*w, not from the corners)
__device__ __forceinline__ Corners to_corners(float4 b) {
  Corners c;
  const float hh = __fdiv_rn(b.z, 2.0f), hw = __fdiv_rn(b.w, 2.0f);
  c.y0 = __fsub_rn(b.x, hh);
  c.x0 = __fsub_rn(b.y, hw);
  c.y1 = __fadd_rn(b.x, hh);
  c.x1 = __fadd_rn(b.y, hw);
  c.area = __fmul_rn(b.z, b.w);
  return c;
}

// torch.max / torch.min propagate NaN; fmaxf / fminf do not.
__device__ __forceinline__ float max_nan(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }
__device__ __forceinline__ float min_nan(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }
__device__ __forceinline__ float clamp_min0(float v) { return (v != v) ? v : fmaxf(v, 0.0f); }

// boxes.py:68-83
__device__ __forceinline__ float iou_pair(const Corners& a, const Corners& b) {
  const float ih = clamp_min0(__fsub_rn(min_nan(a.y1, b.y1), max_nan(a.y0, b.y0)));
  const float iw = clamp_min0(__fsub_rn(min_nan(a.x1, b.x1), max_nan(a.x0, b.x0)));
  const float inter = __fmul_rn(ih, iw);
  const float uni = __fsub_rn(__fadd_rn(a.area, b.area), inter);
  return __fdiv_rn(inter, uni);
}

__device__ __forceinline__ float4 decode_one(float4 e, float4 a, float v0, float v1) {
  float4 o;
  o.x = __fadd_rn(a.x, __fmul_rn(__fmul_rn(e.x, v0), a.z));
  o.y = __fadd_rn(a.y, __fmul_rn(__fmul_rn(e.y, v0), a.w));
  o.z = __fmul_rn(a.z, expf(__fmul_rn(e.z, v1)));
  o.w = __fmul_rn(a.w, expf(__fmul_rn(e.w, v1)));
  return o;
}

__device__ __forceinline__ float4 encode_one(float4 b, float4 a, float v0, float v1) {
  float4 o;
  o.x = __fdiv_rn(__fsub_rn(b.x, a.x), __fmul_rn(v0, a.z));
  o.y = __fdiv_rn(__fsub_rn(b.y, a.y), __fmul_rn(v0, a.w));
  o.z = __fdiv_rn(logf(__fdiv_rn(b.z, a.z)), v1);
  o.w = __fdiv_rn(logf(__fdiv_rn(b.w, a.w)), v1);
  return o;
}

// boxes.py:88-103 — inclusive crop on integer pixel coordinates
struct CropBounds {
  float left, right, top, bottom;
};
__device__ __forceinline__ CropBounds crop_bounds(float4 box, int H, int W) {
  const float by = __fmul_rn(box.x, (float)H), bx = __fmul_rn(box.y, (float)W);
  const float bh = __fmul_rn(box.z, (float)H), bw = __fmul_rn(box.w, (float)W);
  CropBounds c;
  c.left = __fsub_rn(bx, __fdiv_rn(bw, 2.0f));
  c.right = __fadd_rn(bx, __fdiv_rn(bw, 2.0f));
  c.top = __fsub_rn(by, __fdiv_rn(bh, 2.0f));
  c.bottom = __fadd_rn(by, __fdiv_rn(bh, 2.0f));
  return c;
}


}  // namespace tauv
