/*
 * tauv_b200.h — C ABI of libtauv_b200.so: the B200 (sm_100a) detection-head hot path of
 * TAUV-Vision (CenterNet decode / target encode, YOLACT post-process / anchor matching).
 *
 * The reference (pure Python/PyTorch) has no FFI layer: callers bind module-level functions
 * by name.  This header is the boundary a maintainer binds with ctypes;
 * every entry point cites the reference function (file:line under /root/reference) it replaces.
 *
 * Conventions
 *  - All pointers are DEVICE pointers on the current CUDA device unless the name ends in _host.
 *  - fp32 unless stated; "i64" = int64_t; booleans are uint8_t (torch.bool storage).
 *  - Strides are in ELEMENTS (torch .stride()), passed wherever the reference hands us a
 *    permuted view (Prediction.size/offset/depth are NCHW tensors viewed as NHWC).
 *  - The library allocates nothing and keeps no pointer after return.  Scratch comes from the
 *    caller: ask tauv_*_workspace_bytes(), pass a device buffer of at least that size
 *    (256-byte aligned).  Work is enqueued on `stream` (a cudaStream_t); no call synchronises.
 *  - Return value: 0 = OK; < 0 = argument error (TAUV_E_*); > 0 = a cudaError_t.
 *    tauv_last_error() returns a thread-local message for the last non-zero return.
 *  - Re-entrant: no mutable state except an idempotent, lock-protected per-device cache of launch
 *    configuration (occupancy, opted-in shared-memory size); concurrent calls on different
 *    streams/devices are safe.
 *  - There is no CPU path: a device without sm_100 returns TAUV_E_ARCH.
 */
#ifndef TAUV_B200_H
#define TAUV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAUV_B200_VERSION 100 /* 0.1.0 */

enum {
  TAUV_OK = 0,
  TAUV_E_NULL = -1,      /* required pointer is NULL */
  TAUV_E_SHAPE = -2,     /* non-positive / inconsistent shape */
  TAUV_E_K_RANGE = -3,   /* k > C*H*W  (reference: RuntimeError "selected index k out of range") */
  TAUV_E_KERNEL = -4,    /* even / non-positive kernel_size (reference: AssertionError decode.py:243) */
  TAUV_E_WORKSPACE = -5, /* workspace too small or misaligned */
  TAUV_E_UNSUPPORTED = -6, /* shape beyond what the kernels are built for (see message) */
  TAUV_E_ARCH = -7,      /* device is not sm_100 */
  TAUV_E_ALIGN = -8      /* pointer not aligned as required */
};

typedef void* tauv_stream_t; /* cudaStream_t */

int tauv_version(void);
const char* tauv_last_error(void);
/* 0 when the CURRENT device can run the kernels (compute capability 10.x). */
int tauv_check_device(void);

/* ------------------------------------------------------------------------------------------
 * CenterNet
 * ---------------------------------------------------------------------------------------- */

/* heatmap_nms(heatmap, kernel_size)            — centernet/model/decode.py:239-252
 * out = (max_pool2d(in, k, stride 1, pad (k-1)/2) == in) * in, elementwise over [B,C,H,W]
 * (contiguous NCHW).  apply_sigmoid != 0 first maps in -> 1/(1+exp(-in)) (decode.py:182-183
 * fused).  k must be odd and >= 1. */
int tauv_heatmap_nms(const float* in, float* out, int B, int C, int H, int W, int kernel_size,
                     int apply_sigmoid, tauv_stream_t stream);

/* Top-k modes for tauv_heatmap_topk */
enum {
  TAUV_TOPK_RAW = 0,         /* heatmap_detect on the values as given (decode.py:255-279) */
  TAUV_TOPK_SIGMOID_PEAK = 1 /* sigmoid -> 3x3 peak suppression -> heatmap_detect, one pass
                                (decode.py:182-184 / :56-58), never materialising the map */
};

size_t tauv_heatmap_topk_workspace_bytes(int B, int C, int H, int W, int k);

/* heatmap_detect(heatmap, n_detections)        — centernet/model/decode.py:255-279
 * Joint top-k over the flattened C*H*W of every frame.  Order: score descending, ties by
 * flat index ascending (the order the reference's own KAT decode.py:327-339 asserts).
 *   index [B,k,2] i64 (y,x), label [B,k] i64, score [B,k] f32.
 * In SIGMOID_PEAK mode entries past the last positive peak have score 0 and the lowest flat
 * indices whose suppressed value is 0, as a stable top-k of the dense suppressed map gives. */
int tauv_heatmap_topk(const float* heatmap, int B, int C, int H, int W, int k, int mode,
                      int64_t* index, int64_t* label, float* score, void* workspace,
                      size_t workspace_bytes, tauv_stream_t stream);

/* Box-decode modes */
enum {
  TAUV_BOX_DECODE = 0,   /* decode():           y = (ratio*iy + offset_y)/in_h  (fp64), depth = 1/sigmoid(d) - 1
                                                 — decode.py:210-221, :319-324 */
  TAUV_BOX_KEYPOINTS = 1 /* decode_keypoints(): y = iy/out_h (fp32 divide), no offset, depth = 1/sigmoid(d)
                                                 — decode.py:65, :87-91 */
};

/* Per-detection gather + box arithmetic of decode()/decode_keypoints() for the k ranked peaks
 * of every frame.  size/offset: logical [B,H,W,2] with element strides {b,y,x,c}; depth:
 * logical [B,H,W(,1)] with strides {b,y,x} (NULL = no depth head).
 *   yx [B,k,2] f64, hw [B,k,2] f32, depth_out [B,k] f32 (may be NULL iff depth NULL),
 *   count [B] i32 = number of leading entries before the first score < score_threshold
 *   (decode.py:208-209, compared in fp32 like torch). */
int tauv_centernet_boxes(const int64_t* index, const float* score, int B, int k, int H, int W,
                         const float* size, const int64_t size_strides[4], const float* offset,
                         const int64_t offset_strides[4], const float* depth,
                         const int64_t depth_strides[3], int mode, int downsample_ratio, int in_h,
                         int in_w, int out_h, int out_w, float score_threshold, double* yx,
                         float* hw, float* depth_out, int32_t* count, tauv_stream_t stream);

/* decode(prediction, model_config, n_detections, score_threshold) — decode.py:179-236, device part:
 * the result of tauv_heatmap_topk(SIGMOID_PEAK) followed by tauv_centernet_boxes, in ONE launch for
 * k <= 256 (thread-block clusters; the workspace is then not touched), two launches otherwise. */
int tauv_centernet_decode(const float* heatmap_logits, int B, int C, int H, int W, int k,
                          const float* size, const int64_t size_strides[4], const float* offset,
                          const int64_t offset_strides[4], const float* depth,
                          const int64_t depth_strides[3], int mode, int downsample_ratio,
                          int in_h, int in_w, float score_threshold, int64_t* index,
                          int64_t* label, float* score, double* yx, float* hw, float* depth_out,
                          int32_t* count, void* workspace, size_t workspace_bytes,
                          tauv_stream_t stream);

/* Launch 1 of the two tauv_centernet_decode makes for 16-byte aligned maps with W % 4 == 0 and k <= 256, on its own
 * (bench.py times it with stream events: it is the only pass over the logits, centernet/model/decode.py:182,239-252 read
 * them three times).  Leaves in the workspace, per frame, the maximum of every block of 4 columns x 8 rows of every
 * plane and the maximum of every 32 consecutive blocks.  Returns TAUV_E_UNSUPPORTED for other shapes. */
int tauv_centernet_block_maxima(const float* heatmap_logits, int B, int C, int H, int W, int k,
                                void* workspace, size_t workspace_bytes, tauv_stream_t stream);

/* The two-launch form of tauv_centernet_decode as separate calls (profiling tools put stream events
 * between them).  stage1 fills the workspace with candidates (reads the logits once); stage2 merges
 * them per frame and does the box arithmetic.  Same workspace, same shapes, same stream for both. */
int tauv_heatmap_topk_stage1(const float* heatmap, int B, int C, int H, int W, int k, int mode,
                             void* workspace, size_t workspace_bytes, tauv_stream_t stream);
int tauv_centernet_decode_stage2(int B, int C, int H, int W, int k, const float* size,
                                 const int64_t size_strides[4], const float* offset,
                                 const int64_t offset_strides[4], const float* depth,
                                 const int64_t depth_strides[3], int mode, int downsample_ratio,
                                 int in_h, int in_w, float score_threshold, int64_t* index,
                                 int64_t* label, float* score, double* yx, float* hw,
                                 float* depth_out, int32_t* count, void* workspace,
                                 size_t workspace_bytes, tauv_stream_t stream);

/* Gather `nch` channels at the ranked peak positions from a strided [B, ..., H, W] map:
 * out[b,j,c] = src[b*sb + sel(b,j)*ssel + c*sc + iy*sy + ix*sx], where sel is label[b,j]
 * (or 0 if label is NULL).  Used for keypoint_affinity[b,label,0:2,y,x] (decode.py:121-122). */
int tauv_gather_at(const float* src, int64_t sb, int64_t ssel, int64_t sc, int64_t sy,
                   int64_t sx, int nch, const int64_t* index, const int64_t* label, int B, int k,
                   float* out, tauv_stream_t stream);

/* Backward of tauv_gather_at without a label: dst[b*sb + c*sc + iy*sy + ix*sx] = sum of grad[b,j,c] over the objects j
 * of frame b whose index[b,j] is (iy, ix), in object order (no atomics); dst must be zero-filled by the caller.  With
 * tauv_gather_at this is the per-object gather of centernet/model/loss.py:196-227 and its autograd. */
int tauv_scatter_add_at(const float* grad, const int64_t* index, int B, int k, int nch, float* dst,
                        int64_t sb, int64_t sc, int64_t sy, int64_t sx, tauv_stream_t stream);

/* The greedy keypoint -> object association of decode_keypoints — decode.py:98-135, on the device
 * (one warp per frame).  Inputs: the ranked objects of tauv_centernet_decode(mode KEYPOINTS)
 * (label [B,k] i64, yx [B,k,2] f64, count [B] i32) and the ranked keypoint peaks of
 * tauv_heatmap_topk(SIGMOID_PEAK) on the keypoint heatmap (kp_index [B,kk,2], kp_label [B,kk],
 * kp_score [B,kk]); affinity: the [B,Kp,2,H,W] head tensor with its five element strides;
 * kp_map [Kp][2] i32 = object_config.decode_keypoint_index(channel) = (object label, keypoint slot).
 * Keypoints are taken in rank order until the first with (double)score < keypoint_score_threshold
 * (the reference compares Python floats, :100-102); each goes to the free candidate with the
 * smallest |atan2(a_y,a_x) - atan2(k_y-d_y, k_x-d_x)| in doubles, first minimum wins.
 * Outputs per (frame, object rank, keypoint slot): kp_set [B,k,max_kp] u8, kp_yx [B,k,max_kp,2]
 * f32 (index / out_h, index / out_w in fp32, :115-118), kp_score_out [B,k,max_kp] f32,
 * kp_aff_out [B,k,max_kp,2] f32 (a_y, a_x).  k * max_kp <= 4096. */
int tauv_centernet_keypoint_assoc(const int64_t* label, const double* yx, const int32_t* count,
                                  int B, int k, const int64_t* kp_index, const int64_t* kp_label,
                                  const float* kp_score, int kk, const float* affinity,
                                  const int64_t affinity_strides[5], const int32_t* kp_map, int Kp,
                                  int max_kp, int out_h, int out_w,
                                  double keypoint_score_threshold, uint8_t* kp_set, float* kp_yx,
                                  float* kp_score_out, float* kp_aff_out, tauv_stream_t stream);

/* angle_decode(predicted_bin, predicted_offset, theta_range, bin_overlap) — decode.py:291-316
 * Two-bin angle decode over n rows of 4: softmax-select the bin, centre +- pi/2 plus atan2 of the
 * (sin, cos) offset pair, wrapped to [0, 2*pi) and rescaled by theta_range/(2*pi).
 * (bin_overlap only shifts bin limits, which the decode never reads.)  out [n] f32. */
int tauv_angle_decode(const float* predicted_bin, const float* predicted_offset, int64_t n,
                      double theta_range, float* out, tauv_stream_t stream);

/* depth_decode(prediction) = 1/sigmoid(d) - 1, elementwise — decode.py:319-324. */
int tauv_depth_decode(const float* in, int64_t n, float* out, tauv_stream_t stream);

/* generate_heatmap(truth, model_config, train_config, object_config) — centernet/model/loss.py:31-72
 * out[b,c,y,x] = max over valid objects o of frame b with label c of
 *                exp(-((x-cx)^2+(y-cy)^2) / (2*sigma^2)),   0 where no object,
 * cy = floor(center_y*in_h/ratio) (fp32 multiply then divide, unclamped), sigma floored at 0.1.
 * sigma is a double because the reference forms 2*sigma**2 in Python doubles before the fp32 divide.
 *   valid [B,n] u8, label [B,n] i64, center [B,n,2] f32 (y,x), out [B,C,H,W] f32 contiguous. */
int tauv_gaussian_encode(const uint8_t* valid, const int64_t* label, const float* center, int B,
                         int n_objects, int C, int H, int W, int in_h, int in_w,
                         int downsample_ratio, double sigma, float* out, tauv_stream_t stream);

/* The heatmap term of the CenterNet loss fused with its target render:
 *   focal_loss(sigmoid(logits), generate_heatmap(truth), alpha, beta).sum()  — centernet/model/loss.py:182, :233-236,
 *   :302-317.  The [B,C,H,W] target is never written.  Forward leaves, per frame, the sums of the two terms
 *   frame_sums[b] = { sum of (1-p)^alpha log(clamp(p,1e-4)) over the positives (target isclose 1),
 *                     sum of (1-t)^beta p^alpha log(clamp(1-p,1e-4)) over the other cells }   (double, fixed order)
 *   and the number of positives frame_pos[b]; with N = sum_b frame_pos[b] the loss is -(sum_p + sum_n)/N for N > 0 and
 *   -sum_p for N == 0 (loss.py:312-315).  Backward writes d(loss)/d(logits) * grad_out for that N (device scalars, so
 *   neither call synchronises).  Needs W % 4 == 0, 16-byte aligned logits, <= 32 objects per frame and an 8-byte
 *   aligned centre tensor (TAUV_E_UNSUPPORTED otherwise: the Python mirror then composes generate_heatmap with
 *   elementwise torch ops on the GPU).  logits [B,C,H,W] f32 contiguous; truth as for tauv_gaussian_encode. */
size_t tauv_centernet_focal_loss_workspace_bytes(int B, int C, int H, int W);
int tauv_centernet_focal_loss(const float* logits, const uint8_t* valid, const int64_t* label,
                              const float* center, int B, int n_objects, int C, int H, int W, int in_h,
                              int in_w, int downsample_ratio, double sigma, double alpha, double beta,
                              double* frame_sums, int64_t* frame_pos, void* workspace,
                              size_t workspace_bytes, tauv_stream_t stream);
/* The batch's loss from the per-frame sums of tauv_centernet_focal_loss (loss.py:313-317, summed): loss [1] f32 =
 * -(sum_p + sum_n) / N with N = the batch's positive cells, -sum_p when N == 0; n_pos_total [1] i64 = N. */
int tauv_centernet_focal_loss_reduce(const double* frame_sums, const int64_t* frame_pos, int B,
                                     float* loss, int64_t* n_pos_total, tauv_stream_t stream);
int tauv_centernet_focal_loss_backward(const float* logits, const uint8_t* valid, const int64_t* label,
                                       const float* center, int B, int n_objects, int C, int H, int W,
                                       int in_h, int in_w, int downsample_ratio, double sigma,
                                       double alpha, double beta, const int64_t* n_pos_total,
                                       const float* grad_out, float* grad_logits, tauv_stream_t stream);

/* generate_keypoint_heatmap(...) — centernet/model/loss.py:75-135
 *   kp_valid [B,m] u8, kp_label [B,m] i64, kp_center [B,m,2] f32, kp_object_index [B,m] i64,
 *   center [B,n,2] f32 (object centres)
 *   heatmap [B,Kp,H,W], weight [B,Kp,H,W], affinity [B,Kp,2,H,W]  (all f32 contiguous). */
int tauv_keypoint_encode(const uint8_t* kp_valid, const int64_t* kp_label, const float* kp_center,
                         const int64_t* kp_object_index, const float* center, int B, int m,
                         int n_objects, int Kp, int H, int W, int in_h, int in_w,
                         int downsample_ratio, double sigma_heatmap, double sigma_affinity,
                         float* heatmap, float* weight, float* affinity, tauv_stream_t stream);

/* out_index_for_position (loss.py:138-142) and the sub-pixel offset target (loss.py:263-264).
 *   position [n,2] f32 -> index [n,2] i64 (clamped), offset [n,2] f32 (may be NULL). */
int tauv_out_index_offset(const float* position, int64_t n, int in_h, int in_w,
                          int downsample_ratio, int out_h, int out_w, int64_t* index,
                          float* offset, tauv_stream_t stream);

/* gaussian_splat(h, w, cy, cx, sigma) — missing from the reference snapshot; call sites
 * decode.py:328-332, tests/centernet_square_detection.py:108-112.  out [h,w] f32. */
int tauv_gaussian_splat(int h, int w, int cy, int cx, double sigma, float* out,
                        tauv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * YOLACT
 * ---------------------------------------------------------------------------------------- */

/* get_anchor for all FPN levels concatenated — yolact/model/anchors.py:9-41, model.py:47-58.
 * Level l has fpn_h[l] x fpn_w[l] cells; ordering inside a level is aspect-major
 * (idx = a*H*W + i*W + j), exactly as the reference builds it.  out [sum_l A*h*w, 4] (y,x,h,w).
 * hw_host holds the per-(level,aspect) h,w already rounded to fp32 by the caller
 * (2*n_levels*n_aspect floats: h then w) because the reference computes them in Python doubles. */
int tauv_yolact_anchors(const int* fpn_h_host, const int* fpn_w_host, int n_levels, int n_aspect,
                        const float* hw_host, float* out, tauv_stream_t stream);

/* box_decode(box_encoding, anchor, config) — yolact/model/boxes.py:55-61
 *   yx = a_yx + (e_yx*v0)*a_hw ; hw = a_hw*exp(e_hw*v1).  anchor_batch is 1 (broadcast) or B. */
int tauv_yolact_box_decode(const float* encoding, const float* anchor, int B, int N,
                           int anchor_batch, float v0, float v1, float* out, tauv_stream_t stream);

/* box_encode(box, anchor, config) — yolact/model/boxes.py:45-52
 *   g_yx = (b_yx - a_yx)/(v0*a_hw) ; g_hw = log(b_hw/a_hw)/v1. */
int tauv_yolact_box_encode(const float* box, const float* anchor, int B, int N, int anchor_batch,
                           float v0, float v1, float* out, tauv_stream_t stream);

/* iou_matrix(box_a, box_b) — yolact/model/boxes.py:64-85
 *   a [Ba,Na,4], b [Bb,Nb,4] centre-size boxes, Ba/Bb in {1,B} -> out [B,Na,Nb]. */
int tauv_iou_matrix(const float* a, const float* b, int Ba, int Bb, int Na, int Nb, float* out,
                    tauv_stream_t stream);

/* softmax(classification)[..., 1:].max(-1) — yolact/model/nms.py:9-10.
 *   cls [B,N,C1] -> score [B,N] ; argmax_all (optional, i32 [B,N]) = argmax over all C1 classes
 *   (yolact_node.py:129). */
int tauv_yolact_scores(const float* cls, int B, int N, int C1, float* score, int32_t* argmax_all,
                       tauv_stream_t stream);

size_t tauv_yolact_nms_workspace_bytes(int B, int N, int C1, int top_k);

/* nms(classification, box, top_k, iou_threshold, confidence_threshold) — yolact/model/nms.py:7-29
 * Fast NMS, batched over the first n_frames frames (the reference processes frame 0 only).
 *   cls [B,N,C1], box [B,N,4] DECODED boxes ->
 *   keep [n_frames, top_k] i64 (prior indices, descending confidence, ties by prior index),
 *   n_keep [n_frames] i32. */
int tauv_yolact_fast_nms(const float* cls, const float* box, int B, int N, int C1, int n_frames,
                         int top_k, float iou_threshold, float confidence_threshold,
                         int64_t* keep, int32_t* n_keep, void* workspace, size_t workspace_bytes,
                         tauv_stream_t stream);

/* Fused post-process head (yolact_node.py:127-130): scores -> top_k -> box_decode of the top_k
 * priors only -> Fast NMS.  Also returns the decoded boxes / confidence / class of the kept
 * priors so the caller never touches the [B,N,*] tensors again.
 *   enc [B,N,4], anchor [1 or B,N,4] ->
 *   keep [B,top_k] i64, n_keep [B] i32, keep_box [B,top_k,4] f32, keep_score [B,top_k] f32,
 *   keep_class [B,top_k] i32 (argmax over all C1 classes). */
int tauv_yolact_detect(const float* cls, const float* enc, const float* anchor, int B, int N,
                       int C1, int anchor_batch, float v0, float v1, int top_k,
                       float iou_threshold, float confidence_threshold, int64_t* keep,
                       int32_t* n_keep, float* keep_box, float* keep_score, int32_t* keep_class,
                       void* workspace, size_t workspace_bytes, tauv_stream_t stream);

/* assemble_mask(mask_prototype, mask_coeff, box) — yolact/model/masks.py:8-21 with
 * box_to_mask boxes.py:88-103.   mask[i] = sigmoid(sum_p coeff[i,p]*proto[p]) * crop(box[i]).
 * The contraction runs on the tensor cores (tcgen05, bf16 operands, fp32 accumulate in TMEM).
 *   proto [P,H,W] f32, coeff [n,P] f32, box [n,4] f32 or NULL -> out [n,H,W] f32.
 * logits_out (optional, [n,H,W]) receives the pre-sigmoid accumulator for tolerance checks. */
int tauv_yolact_assemble_mask(const float* proto, const float* coeff, const float* box, int n,
                              int P, int H, int W, float* out, float* logits_out,
                              tauv_stream_t stream);

/* Batched mask assembly straight from the detect() outputs: frame b assembles n_keep[b] masks
 * from coeff_all[b, keep[b,i], :] and keep_box[b,i].
 *   proto [B,P,H,W], coeff_all [B,N,P], keep [B,top_k] i64, n_keep [B] i32,
 *   keep_box [B,top_k,4] (NULL = no crop) -> out [B,top_k,H,W] (rows >= n_keep[b] untouched). */
int tauv_yolact_assemble_mask_batched(const float* proto, const float* coeff_all,
                                      const int64_t* keep, const int32_t* n_keep,
                                      const float* keep_box, int B, int N, int P, int H, int W,
                                      int top_k, float* out, tauv_stream_t stream);

/* Mask assembly fused with its consumer — yolact/node/yolact_node.py:102-103,130-131,178-183:
 *     depth = where(depth_mm == 0, nan, depth_mm) / 1000                        (:102-103, float64)
 *     mask  = F.interpolate(assemble_mask(proto, coeff, box)[None], (Hi, Wi))   (:130-131, 'nearest')
 *     mean_depth[i] = nanmean(where(mask[i] > 0.5, depth, nan))                 (:178)
 * computed without materialising any mask: the depth image is pooled onto the prototype grid with the inverse of
 * the nearest-neighbour map (sum and count of the valid readings per prototype pixel, exact integers), and the
 * tensor-core kernel's epilogue adds the pooled values of the pixels that are on (inside the box and logit > 0,
 * i.e. sigmoid > 0.5) per detection.  Traffic per frame: the prototypes once (4*P*H*W) instead of 4*n*H*W of masks.
 *   depth_mm [Hi,Wi] u16 (ROS mono16, 0 = no reading) ->
 *   mean [n] f64 metres (NaN when no selected pixel has a reading), count [n] i64 readings averaged (may be NULL).
 * Sums are integers: the result is deterministic; it differs from the reference only where a logit is within the
 * bf16x2 contraction error (~2e-5) of zero.  workspace: tauv_yolact_mask_depth_workspace_bytes(1, H, W, n). */
size_t tauv_yolact_mask_depth_workspace_bytes(int B, int H, int W, int top_k);
int tauv_yolact_mask_depth(const float* proto, const float* coeff, const float* box, int n, int P,
                           int H, int W, const uint16_t* depth_mm, int Hi, int Wi, double* mean,
                           int64_t* count, void* workspace, size_t workspace_bytes,
                           tauv_stream_t stream);

/* The same for every frame of a batch, straight from the detect() outputs (one depth image per frame).
 *   proto [B,P,H,W], coeff_all [B,N,P], keep [B,top_k], n_keep [B], keep_box [B,top_k,4] (NULL = no crop),
 *   depth_mm [B,Hi,Wi] u16 -> mean [B,top_k] f64 (NaN for rows >= n_keep[b]), count [B,top_k] i64 (may be NULL). */
int tauv_yolact_mask_depth_batched(const float* proto, const float* coeff_all, const int64_t* keep,
                                   const int32_t* n_keep, const float* keep_box, int B, int N,
                                   int P, int H, int W, int top_k, const uint16_t* depth_mm, int Hi,
                                   int Wi, double* mean, int64_t* count, void* workspace,
                                   size_t workspace_bytes, tauv_stream_t stream);

/* Binarised masks at the camera / network-input resolution: what the reference's callers make of
 * assemble_mask before anything else looks at it —
 *     mask = F.interpolate(assemble_mask(proto, coeff, box)[None], (out_h, out_w))[0];  mask_np > 0.5
 *                                   yolact/node/yolact_node.py:135, :178           (mode TAUV_RESIZE_NEAREST)
 *     mask = F.interpolate(assemble_mask(...)[None], (out_h, out_w), mode="bilinear")[0];  mask = mask > 0.5
 *                                   yolact/scripts/evaluate_batch.py:101-102       (mode TAUV_RESIZE_BILINEAR)
 * as one byte per pixel (1 = on): a quarter of the bytes of the fp32 mask the reference upsamples, and the only
 * full-resolution array written.  Index rules are ATen's: nearest takes source min(floor(dst * fp32(in/out)), in-1);
 * bilinear (align_corners=False) takes max(fp32(in/out) * (dst + 0.5) - 0.5, 0), its integer part and the next pixel.
 *   out [n,out_h,out_w] u8.  Differences from the reference are confined to pixels whose (interpolated) mask value is
 *   within the contraction error of 0.5 (|logit| < ~2e-5 for nearest; |value - 0.5| < ~1e-5 for bilinear, which
 *   switches the tensor-core epilogue to its 2-ulp sigmoid).
 *   workspace (256-byte aligned): tauv_yolact_mask_binary_workspace_bytes(1, H, W, n) — the low-resolution masks. */
#define TAUV_RESIZE_NEAREST 0
#define TAUV_RESIZE_BILINEAR 1
size_t tauv_yolact_mask_binary_workspace_bytes(int B, int H, int W, int top_k);
int tauv_yolact_mask_binary(const float* proto, const float* coeff, const float* box, int n, int P,
                            int H, int W, int out_h, int out_w, int mode, uint8_t* out,
                            void* workspace, size_t workspace_bytes, tauv_stream_t stream);

/* The same for every frame of a batch, straight from the detect() outputs.
 *   out [B,top_k,out_h,out_w] u8; rows >= n_keep[b] are not written. */
int tauv_yolact_mask_binary_batched(const float* proto, const float* coeff_all, const int64_t* keep,
                                    const int32_t* n_keep, const float* keep_box, int B, int N,
                                    int P, int H, int W, int top_k, int out_h, int out_w, int mode,
                                    uint8_t* out, void* workspace, size_t workspace_bytes,
                                    tauv_stream_t stream);

/* box_to_mask(box, img_size) — yolact/model/boxes.py:88-103.  box [4] f32 -> out [H,W] {0,1}. */
int tauv_box_to_mask(const float* box, int H, int W, float* out, tauv_stream_t stream);

/* Anchor matching + regression targets — yolact/model/loss.py:16-22, :62-66
 *   anchor [1,N,4], truth_box [B,M,4], truth_valid [B,M] u8 ->
 *   match_index [B,N] i64 (first max on ties), match_iou [B,N] f32,
 *   positive [B,N] u8 (iou >= pos_thr), negative [B,N] u8 (iou <= neg_thr),
 *   target [B,N,4] f32 = box_encode(truth_box[match_index], anchor) where positive, zeros elsewhere
 *   (may be NULL). */
int tauv_yolact_match_anchors(const float* anchor, const float* truth_box,
                              const uint8_t* truth_valid, int B, int N, int M, float pos_thr,
                              float neg_thr, float v0, float v1, int64_t* match_index,
                              float* match_iou, uint8_t* positive, uint8_t* negative,
                              float* target, tauv_stream_t stream);

/* Keypoint-affinity term of the CenterNet loss fused with its target render — centernet/model/loss.py:244-246 before the
 * lambda: (affinity_weight.unsqueeze(2) * F.mse_loss(prediction.keypoint_affinity, keypoint_affinity,
 * reduction="none")).sum() with the weight and the unit-vector field of generate_keypoint_heatmap (loss.py:105-129)
 * computed on the fly, never written.  pred_affinity [B,Kp,2,H,W] f32 (W % 4 == 0, 16-byte aligned), keypoint truth as in
 * tauv_keypoint_encode (1 <= m <= 128) -> partial [tauv_keypoint_affinity_loss_partials(B, Kp, H, W)] f64 whose sum is the
 * term; planes without instances are not read.  Backward: grad_affinity [B,Kp,2,H,W] = 2 * grad_out[0] * weight *
 * (pred - target). */
size_t tauv_keypoint_affinity_loss_partials(int B, int Kp, int H, int W);
int tauv_keypoint_affinity_loss(const float* pred_affinity, const uint8_t* kp_valid,
                                const int64_t* kp_label, const float* kp_center,
                                const int64_t* kp_object_index, const float* center, int B, int m,
                                int n_objects, int Kp, int H, int W, int in_h, int in_w,
                                int downsample_ratio, double sigma_affinity, double* partial,
                                tauv_stream_t stream);
int tauv_keypoint_affinity_loss_backward(const float* pred_affinity, const uint8_t* kp_valid,
                                         const int64_t* kp_label, const float* kp_center,
                                         const int64_t* kp_object_index, const float* center, int B,
                                         int m, int n_objects, int Kp, int H, int W, int in_h, int in_w,
                                         int downsample_ratio, double sigma_affinity,
                                         const float* grad_out, float* grad_affinity,
                                         tauv_stream_t stream);

/* Classification (hard-negative mining) and box terms of the YOLACT loss, forward — yolact/model/loss.py:26-56
 * (per-frame target classes, F.cross_entropy(reduction="none"), torch.topk of -softmax[:, 0] over the negative priors
 * with k = ratio * n_positive, sum over positives + mined negatives) and :58-68 (smooth-L1 over the positives).
 *   cls [B,N,C1] f32 class logits, enc [B,N,4] predicted box encodings, and from tauv_yolact_match_anchors:
 *   target [B,N,4], positive / negative [B,N] u8, match_index [B,N] i64; truth_cls [B,M] i64 ->
 *   selected [B,N] u8 (positives and mined negatives; ties of the background confidence go to the lower prior index),
 *   pos_list [B,N] i32 (each frame's positives in prior order, first n_pos[b] entries; may be NULL),
 *   sums [B,2] f64 (sum of the selected cross entropies, sum of the positives' smooth-L1), n_pos [B] i64.
 * The caller normalises (loss.py:54-57, :70-73): class term = sum(sums[:,0]) / ((1 + ratio) * P), box term =
 * sum(sums[:,1]) / P with P = sum(n_pos), undivided when P == 0.  Deterministic (fixed summation order). */
size_t tauv_yolact_class_box_loss_workspace_bytes(int B, int N);
int tauv_yolact_class_box_loss(const float* cls, const float* enc, const float* target,
                               const uint8_t* positive, const uint8_t* negative,
                               const int64_t* match_index, const int64_t* truth_cls, int B, int N,
                               int C1, int M, int ratio, uint8_t* selected, int32_t* pos_list,
                               double* sums, int64_t* n_pos, void* workspace, size_t workspace_bytes,
                               tauv_stream_t stream);

/* Backward of the two terms (what autograd derives from loss.py:26-73): grad_cls [B,N,C1] = g_cls / ((1 + ratio) P) *
 * selected * (softmax(cls) - onehot(target class)), grad_enc [B,N,4] = g_box / P * positive * smooth_l1'(enc - target);
 * n_pos_total [1] i64 = P (device), grad_cls_loss / grad_box_loss [1] f32 (device).  Either output (with its incoming
 * gradient) may be NULL. */
int tauv_yolact_class_box_loss_backward(const float* cls, const float* enc, const float* target,
                                        const uint8_t* positive, const uint8_t* selected,
                                        const int64_t* match_index, const int64_t* truth_cls, int B,
                                        int N, int C1, int M, int ratio, const int64_t* n_pos_total,
                                        const float* grad_cls_loss, const float* grad_box_loss,
                                        float* grad_cls, float* grad_enc, tauv_stream_t stream);

/* Mask term of the YOLACT loss, forward — yolact/model/loss.py:75-115: for every positive prior (pos_list / n_pos of
 * tauv_yolact_class_box_loss) the BCE between clamp(sigmoid(coeff . proto), 1e-4) and the bilinearly resized mask of
 * its matched truth (seg == match_index), weighted by box_to_mask(truth_box) * nearest-resized img_valid, over the
 * resized truth mask's area.
 *   coeff [B,N,K] f32 (K <= 32), proto [B,K,PH,PW] f32, match_index [B,N] i64, truth_box [B,M,4] f32,
 *   seg [B,SH,SW] (truth index per pixel; seg_bytes = 1: uint8 as the reference's dataset holds it, 4: int32,
 *   8: int64), img_valid [B,SH,SW] u8 ->
 *   tsum [B,M] f64 (area of every truth's resized mask; kept for the backward),
 *   records [tauv_yolact_mask_loss_records_bytes(B, N)] bytes, 16-byte aligned (one 32-byte record per listed positive:
 *   crop box, area, prior, truth — gathered once per call; kept for the backward),
 *   partial [B, tauv_yolact_mask_loss_partials()] f64 whose sum is the sum over positives of (weighted BCE / area),
 *   positives with an empty resized truth mask skipped (loss.py:93-94).  The caller divides by the batch's positives
 *   (loss.py:117-120).  Deterministic. */
int tauv_yolact_mask_loss_partials(void);
size_t tauv_yolact_mask_loss_records_bytes(int B, int N);
int tauv_yolact_mask_loss(const float* coeff, const float* proto, const int32_t* pos_list,
                          const int64_t* n_pos, const int64_t* match_index, const float* truth_box,
                          const void* seg, int seg_bytes, const uint8_t* img_valid, int B, int N, int K,
                          int M,
                          int PH, int PW, int SH, int SW, double* tsum, void* records,
                          double* partial, tauv_stream_t stream);

/* Backward of the mask term: grad_coeff [B,N,K] (zero outside the positives) and grad_proto [B,K,PH,PW]; either may be
 * NULL.  n_pos_total [1] i64 and grad_out [1] f32 on the device.  No atomics: deterministic. */
int tauv_yolact_mask_loss_backward(const float* coeff, const float* proto, const int32_t* pos_list,
                                   const int64_t* n_pos, const int64_t* match_index,
                                   const float* truth_box, const void* seg, int seg_bytes,
                                   const uint8_t* img_valid, int B, int N, int K, int M, int PH, int PW,
                                   int SH, int SW, const double* tsum, const void* records,
                                   const int64_t* n_pos_total,
                                   const float* grad_out, float* grad_coeff, float* grad_proto,
                                   tauv_stream_t stream);

/* Normalisation of the loss terms — yolact/model/loss.py:54-57, :70-73, :117-120 — from the per-frame sums of
 * tauv_yolact_class_box_loss (sums [B,2], may be NULL) and / or the partial sums of tauv_yolact_mask_loss (mask_partial
 * [n_partial], may be NULL): losses[0] = class term, losses[1] = box term, losses[2] = mask term (only the ones whose
 * input is given are written), n_pos_total [1] = the batch's positives P (may be NULL).  One small launch, fixed order. */
int tauv_yolact_loss_reduce(const double* sums, const int64_t* n_pos, int B, int ratio,
                            const double* mask_partial, int n_partial, float* losses,
                            int64_t* n_pos_total, tauv_stream_t stream);

/* Prediction-head outputs packed into the consumers' layout — yolact/model/prediction_head.py:111-113, :122-124,
 * :137-140 (permute(0, 2, 3, 1).reshape(B, -1, C), tanh for the mask coefficients) and model.py:55-58 (torch.cat over
 * the FPN levels): levels[l] [B,CH,H_l,W_l] f32 NCHW (CH = A * C; host array of n_levels <= 8 device pointers),
 * hw[l] = H_l * W_l (host) -> packed [B, sum_l hw[l] * CH] = [B, N, C] with N = A * sum_l hw[l]; tanh_act != 0 applies
 * tanh.  One read and one write of every element. */
int tauv_yolact_pack_heads(const float* const* levels, const int* hw, int n_levels, int B, int CH,
                           int tanh_act, float* packed, tauv_stream_t stream);

/* Its backward: grad_packed [B,N,C] (and the forward output y_packed when tanh_act) -> grad_levels[l] [B,CH,H_l,W_l]. */
int tauv_yolact_pack_heads_backward(const float* grad_packed, const float* y_packed, const int* hw,
                                    int n_levels, int B, int CH, int tanh_act,
                                    float* const* grad_levels, tauv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TAUV_B200_H */
