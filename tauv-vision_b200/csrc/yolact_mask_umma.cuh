// yolact_mask_umma.cuh — tcgen05 / TMEM mask contraction (included by yolact_mask.cu).
#pragma once

namespace tauv {

static bool umma_shape_ok(const MaskArgs&) { return false; }

static int launch_mask_umma(const MaskArgs&, int, int, cudaStream_t) {
  return fail(TAUV_E_UNSUPPORTED, "tensor-core mask kernel not built");
}

}  // namespace tauv
