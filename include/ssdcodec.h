/*
 * ssdcodec.h - C ABI of libssdcodec.so, the B200 (sm_100a) SSD box codec.
 *
 * This is the drop-in boundary for the SSD box codec hot path of
 * Shulk97/JPEG_detection_Resnet_SSD (`localisation_part/`).  The reference has
 * no FFI of its own (it is pure Python/numpy); each entry point below replaces
 * the numpy body of one reference function and is bound through `ctypes` by the
 * Python modules that keep the reference's import paths and signatures
 * (see INTEGRATION.md).  Citations are relative to
 * /root/reference/localisation_part/.
 *
 * Conventions
 *   - plain pointers and sizes, no C++/torch types; all functions return
 *     SSDC_OK (0) or a negative SSDC_ERR_* code; `ssdc_last_error()` returns a
 *     thread-local human readable message for the last failure.
 *   - the caller owns every host buffer; the library never keeps a caller
 *     pointer past the call.  Device scratch, streams and pinned staging are
 *     owned by the `ssdc_ctx`.
 *   - every entry point is thread safe (per-context mutex, `cudaSetDevice` on
 *     entry); ctypes releases the GIL for the duration of a call.
 *   - there is NO CPU fallback: without a usable CUDA device `ssdc_init` fails.
 */
#ifndef SSDCODEC_H_
#define SSDCODEC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSDC_VERSION 100  /* major*10000 + minor*100 + patch */

/* ---- error codes --------------------------------------------------------- */
#define SSDC_OK               0
#define SSDC_ERR_CUDA        -1   /* a CUDA runtime call failed                 */
#define SSDC_ERR_ARG         -2   /* invalid argument                           */
#define SSDC_ERR_CAPACITY    -3   /* caller buffer too small (see total_rows)   */
#define SSDC_ERR_DEGENERATE  -4   /* degenerate ground-truth box (encoder)      */
#define SSDC_ERR_NODEVICE    -5   /* no CUDA device / not an sm_100 device      */
#define SSDC_ERR_STATE       -6   /* call sequence error (e.g. fetch w/o decode)*/

/* ---- enums --------------------------------------------------------------- */
#define SSDC_F32 0
#define SSDC_F64 1

#define SSDC_COORDS_CENTROIDS 0   /* (cx, cy, w, h)           */
#define SSDC_COORDS_MINMAX    1   /* (xmin, xmax, ymin, ymax) */
#define SSDC_COORDS_CORNERS   2   /* (xmin, ymin, xmax, ymax) */

#define SSDC_BORDER_HALF      0   /* d =  0 */
#define SSDC_BORDER_INCLUDE   1   /* d = +1 */
#define SSDC_BORDER_EXCLUDE   2   /* d = -1 */

#define SSDC_MODE_PER_CLASS   0   /* decode_detections        ssd_output_decoder.py:111-226 */
#define SSDC_MODE_FAST        1   /* decode_detections_fast   ssd_output_decoder.py:228-333 */
#define SSDC_MODE_LAYER       2   /* DecodeDetections layer   keras_layers/keras_layer_DecodeDetections.py:109-265 */
#define SSDC_MODE_LAYER_FAST  3   /* DecodeDetectionsFast     keras_layers/keras_layer_DecodeDetectionsFast.py:111-248 */

#define SSDC_CONV_MINMAX2CENTROIDS  0
#define SSDC_CONV_CENTROIDS2MINMAX  1
#define SSDC_CONV_CORNERS2CENTROIDS 2
#define SSDC_CONV_CENTROIDS2CORNERS 3
#define SSDC_CONV_MINMAX2CORNERS    4
#define SSDC_CONV_CORNERS2MINMAX    5

#define SSDC_IOU_OUTER        0   /* mode='outer_product' */
#define SSDC_IOU_ELEMENTWISE  1   /* mode='element-wise'  */

typedef struct ssdc_ctx ssdc_ctx;
typedef struct ssdc_encoder ssdc_encoder;

/* ---- library / context ---------------------------------------------------- */
int         ssdc_version(void);
const char* ssdc_last_error(void);
int         ssdc_device_count(void);

/* Creates a context spanning `n_devices` CUDA devices (batch shards are spread
 * over them in contiguous slices).  `device_ids == NULL`
 * means device 0 only. */
int  ssdc_init(const int* device_ids, int n_devices, ssdc_ctx** out);
void ssdc_destroy(ssdc_ctx* ctx);
int  ssdc_ctx_num_devices(const ssdc_ctx* ctx);
int  ssdc_synchronize(ssdc_ctx* ctx);

/* Per-context options (test / diagnosis switches; none changes a result).  Unknown option: SSDC_ERR_ARG. */
#define SSDC_OPT_NO_SWEEP          0  /* 1: decode_detections runs the general per-class pipeline instead of the image sweep */
#define SSDC_OPT_FLOOR_TARGET      1  /* image sweep: number of best candidates per image that D1 always keeps complete
                                         (speculative score floor, exact fallback below it); 0 = default max(1024, 5 top_k),
                                         < 0 = no floor                                                                 */
#define SSDC_OPT_ENC_GENERAL       2  /* 1: encoder takes the general matching path                                    */
#define SSDC_OPT_ENC_NO_OVERLAP    3  /* 1: encoder writes y_encoded with the fused write kernel (no template stream)  */
#define SSDC_OPT_ENC_DENSE_PATCH   4  /* 1: encoder patches from the dense match array instead of the position list    */
#define SSDC_OPT_LOSS_NO_TMA       5  /* 1: loss uses the plain tile-copy loader                                       */
#define SSDC_OPT_H2D_CHUNK_MB      6  /* host input of ssdc_decode_submit is copied in chunks of this many MiB, each chunk
                                         filtered (D1) while the next one is in flight; 0 = default (64), <0 = one copy */
#define SSDC_OPT_NO_PIPELINE       7  /* 1: the sweep of a device-resident image-sweep decode runs on the main stream (no overlap
                                         with D1 of the next decode)                                                   */
#define SSDC_OPT_ENC_LANES         8  /* device-output encodes (ssdc_encode, on_device != 0) run on this many lanes - stream pair +
                                         scratch each - so consecutive calls overlap; 0 = default (4), 1 = main stream only     */
#define SSDC_OPT_D1_CTAS           9  /* resident D1 CTAs per SM (TMA loader); 0 = default                                     */
#define SSDC_OPT_NO_L2_HINTS       10 /* 1: D1's bulk copies of y_pred carry no L2 eviction hint.  Default: evict_first - the batch
                                         crosses the L2 once and must not displace the keys / histograms D1 leaves for the sweep */
#define SSDC_OPT_D1_WARPS          11 /* consumer warps (= 32-row slices of a tile) per D1 CTA; 0 = default (8)                  */
#define SSDC_OPT_COUNT             12
int     ssdc_set_option(ssdc_ctx* ctx, int option, int64_t value);
int64_t ssdc_get_option(const ssdc_ctx* ctx, int option);

/* Kernel launches issued by this context since creation (bench: gpu_launches). */
int64_t ssdc_launch_count(const ssdc_ctx* ctx);

/* Per-kernel-family device timing (CUDA events on the context's own stream).
 * When enabled every kernel launch is bracketed by events; `ssdc_profile_read`
 * synchronises and returns accumulated milliseconds and launch counts per
 * family, then resets the accumulators.  Family ids: see SSDC_K_*.           */
#define SSDC_K_DECODE_FILTER 0   /* D1  anchor-offset decode + threshold + compaction */
#define SSDC_K_PLAN          1   /*     segment work lists                          */
#define SSDC_K_SORT          2   /* D2  segmented sort by (score desc, anchor asc)  */
#define SSDC_K_NMS           3   /* D3  greedy NMS                                   */
#define SSDC_K_MERGE         4   /* D4  cross-class top-k + output packing           */
#define SSDC_K_ENC_ROWBEST   5   /* E1  GT x anchor IoU, per-GT best anchor          */
#define SSDC_K_ENC_MATCH     6   /* E2  bipartite greedy rounds                      */
#define SSDC_K_ENC_WRITE     7   /* E3  y_encoded write-out (template stream, or the fused write kernel) */
#define SSDC_K_THIN          8   /*     standalone iou / convert / match ops         */
#define SSDC_K_ENC_PATCH     9   /* E3' rows of matched / neutral anchors patched into the streamed template */
#define SSDC_K_COUNT         10
int ssdc_profile_enable(ssdc_ctx* ctx, int on);
int ssdc_profile_read(ssdc_ctx* ctx, double* ms /*SSDC_K_COUNT*/, int64_t* launches /*SSDC_K_COUNT*/);

/* Device timing of a span of work: `ssdc_timer_start` records a CUDA event on every device's
 * stream of the context, `ssdc_timer_stop` records the closing events, waits for them and returns
 * the longest elapsed time over the devices.                                  */
int ssdc_timer_start(ssdc_ctx* ctx);
int ssdc_timer_stop(ssdc_ctx* ctx, double* elapsed_ms /* max over devices */);

/* Device / pinned memory helpers (so a producer can hand device-resident or
 * pinned buffers to the codec; used by bench.py and the tests). */
int ssdc_dev_alloc(ssdc_ctx* ctx, int dev_slot, uint64_t bytes, void** out);
int ssdc_dev_free(ssdc_ctx* ctx, int dev_slot, void* p);
int ssdc_host_alloc(uint64_t bytes, void** out);            /* pinned */
int ssdc_host_free(void* p);
int ssdc_memcpy_h2d(ssdc_ctx* ctx, int dev_slot, void* dst, const void* src, uint64_t bytes);
int ssdc_memcpy_d2h(ssdc_ctx* ctx, int dev_slot, void* dst, const void* src, uint64_t bytes);

/* ---- decoder --------------------------------------------------------------
 * Replaces the numpy bodies of
 *   decode_detections        ssd_output_decoder.py:111-226
 *   decode_detections_fast   ssd_output_decoder.py:228-333
 *   decode_detections_debug  ssd_output_decoder.py:342-467 (out_anchor_idx)
 *   DecodeDetections{,Fast}.call  keras_layers/...:109-265 / :111-248
 */
typedef struct {
    int32_t mode;           /* SSDC_MODE_*                                         */
    int32_t input_coords;   /* SSDC_COORDS_*                                       */
    int32_t normalize;      /* 1: multiply by img_w / img_h                        */
    int32_t border_pixels;  /* SSDC_BORDER_*                                       */
    int32_t top_k;          /* <= 0: 'all'                                         */
    int32_t nms_cap;        /* layer modes: nms_max_output_size; else ignored      */
    int32_t log_wh;         /* 1: w = exp(ow*vw)*wa (default); 0: *_no_log twin    */
    int32_t do_nms;         /* 0: skip NMS (`if iou_threshold:` falsy, :326)       */
    double  conf_thresh;
    double  iou_thresh;
    double  img_h, img_w;
} ssdc_decode_params;

/* Enqueue the whole decode pipeline for `y_pred` (B, A, C+12), dtype SSDC_F32 or
 * SSDC_F64.  `on_device != 0`: `y_pred` is a device pointer on dev_slot 0 and
 * nothing is copied or synchronised (single-device contexts only); otherwise
 * it is a host pointer (pinned or pageable) and the batch is sharded across
 * the context's devices and copied asynchronously.  Results stay on the
 * device(s) until `ssdc_decode_collect`. */
int ssdc_decode_submit(ssdc_ctx* ctx, const void* y_pred, int dtype, int on_device,
                       int64_t B, int64_t A, int C, const ssdc_decode_params* p);

/* Waits for the submitted decode, writes per-image row counts to
 * `out_counts[B]` and the rows of all images back to back to `out_rows`
 * (`row_width` doubles per row: [class, conf, xmin, ymin, xmax, ymax]) and the
 * anchor index of every row to `out_anchor_idx` (may be NULL).  `*total_rows`
 * receives the number of rows.  If `capacity_rows` is too small nothing is
 * written to out_rows, SSDC_ERR_CAPACITY is returned and the call may be
 * repeated with a larger buffer. */
int ssdc_decode_collect(ssdc_ctx* ctx, double* out_rows, int64_t capacity_rows,
                        int32_t* out_counts, int32_t* out_anchor_idx, int64_t* total_rows);

/* Device-resident results of the submitted decode on `dev_slot`, for consumers that stay on the GPU (no
 * counterpart in the reference; SURVEY section 7 hard part 5).  Available when the decode ran with a finite
 * `top_k` on float32 input in per-class or layer mode (the image-sweep path): `*rows` points to
 * (B_slot, top_k, 6) float64 rows [class, conf, xmin, ymin, xmax, ymax] - the layout of the reference's
 * `DecodeDetections` layer output, rows beyond an image's count are unspecified -, `*anchors` to (B_slot, top_k)
 * anchor indices, `*counts` to (B_slot,) row counts; `*b0` / `*n_images` give the image range of the slot's
 * shard.  The pointers are valid until the next decode on this context; work on them has to be ordered after
 * the context's stream (`ssdc_synchronize`).  SSDC_ERR_STATE for other configurations (use `ssdc_decode_collect`). */
int ssdc_decode_results_dev(ssdc_ctx* ctx, int dev_slot, const double** rows, const int32_t** anchors,
                            const int32_t** counts, int64_t* b0, int64_t* n_images, int32_t* top_k);

/* Statistics of the last collected decode that took the image-sweep path (zeros otherwise), summed over the
 * context's devices: out3[0] = candidate keys D1 emitted, out3[1] = images whose speculative score floor engaged
 * (candidates below it were not emitted), out3[2] = images that ran dry inside the trusted set and took the exact
 * fallback (rescan without a floor). */
int ssdc_decode_stats(ssdc_ctx* ctx, int64_t* out3);

/* submit + collect. */
int ssdc_decode(ssdc_ctx* ctx, const void* y_pred, int dtype, int64_t B, int64_t A, int C,
                const ssdc_decode_params* p, double* out_rows, int64_t capacity_rows,
                int32_t* out_counts, int32_t* out_anchor_idx, int64_t* total_rows);

/* greedy_nms / _greedy_nms / _greedy_nms2   ssd_output_decoder.py:27-109.
 * `boxes` (n,4) float64 in `coords` format, `scores` (n,) float64.  Writes the
 * indices of the kept boxes in keep order to `out_keep` (capacity n) and their
 * number to `*n_keep`. */
int ssdc_greedy_nms(ssdc_ctx* ctx, const double* boxes, const double* scores, int64_t n,
                    double iou_thresh, int coords, int border_pixels,
                    int32_t* out_keep, int64_t* n_keep);

/* ---- encoder --------------------------------------------------------------
 * Replaces SSDInputEncoder.__call__            ssd_input_encoder.py:277-418
 *          generate_encoding_template          ssd_input_encoder.py:550-611
 * (anchor generation :420-548 is construction-time host configuration; the
 * anchors are handed over once and cached in device memory.) */
typedef struct {
    int32_t n_classes;        /* including background                            */
    int32_t background_id;
    int32_t coords;           /* SSDC_COORDS_* : layout of `anchors` and targets */
    int32_t border_pixels;
    int32_t matching_multi;   /* 1: matching_type == 'multi'                     */
    int32_t normalize;        /* divide GT by img size                           */
    int32_t log_wh;           /* 0: *_no_log twin                                */
    int32_t reserved;
    double  pos_iou_threshold;
    double  neg_iou_limit;
    double  img_h, img_w;
} ssdc_encode_params;

int  ssdc_encoder_create(ssdc_ctx* ctx, const double* anchors /*A*4, in `coords` format*/,
                         int64_t A, const double* variances /*4*/,
                         const ssdc_encode_params* p, ssdc_encoder** out);
void ssdc_encoder_destroy(ssdc_encoder* enc);

/* Index of the first image whose ground truth is degenerate after the last
 * ssdc_encode* call returned SSDC_ERR_DEGENERATE. */
int64_t ssdc_encoder_bad_image(const ssdc_encoder* enc);

/* `gt` holds the ground-truth rows [class, xmin, ymin, xmax, ymax] (absolute
 * pixels) of all images back to back; image i owns rows gt_offsets[i] ..
 * gt_offsets[i+1]-1.  Writes y_encoded (B, A, n_classes+12) float64, optionally
 * the `diagnostics=True` copy y_matched (:412-416) and match_idx (B, A) int32:
 * matched ground-truth row, -1 background, -2 neutral.  `on_device != 0`: the
 * three output pointers are device pointers on dev_slot 0 and the call only
 * enqueues work: consecutive such calls run beside one another (SSDC_OPT_ENC_LANES;
 * calls whose outputs overlap are ordered), every other entry point of the library
 * that touches device memory (ssdc_decode_submit, ssdc_ssd_loss, ssdc_memcpy_*,
 * ssdc_synchronize, ssdc_timer_*) runs behind them. */
int ssdc_encode(ssdc_encoder* enc, const double* gt, const int64_t* gt_offsets, int64_t B,
                int on_device, double* y_encoded, double* y_matched, int32_t* match_idx);

/* generate_encoding_template(batch_size)  ssd_input_encoder.py:550-611 */
int ssdc_encoding_template(ssdc_encoder* enc, int64_t B, double* out);

/* ---- thin ops --------------------------------------------------------------- */
/* iou()  bounding_box_utils.py:283-383.  outer: out is (m,n); element-wise: m and n
 * must be equal or one of them 1, out is (max(m,n),). */
int ssdc_iou(ssdc_ctx* ctx, const double* boxes1, int64_t m, const double* boxes2, int64_t n,
             int coords, int mode, int border_pixels, double* out);
/* intersection_area() / intersection_area_()  bounding_box_utils.py:119-280 (here the
 * side lengths do include the border term d). */
int ssdc_intersection_area(ssdc_ctx* ctx, const double* boxes1, int64_t m, const double* boxes2, int64_t n,
                           int coords, int mode, int border_pixels, double* out);
/* convert_coordinates()  bounding_box_utils.py:24-87 on `rows` rows of `width`
 * elements (dtype SSDC_F32/F64 input, float64 output), 4 coords at `start`. */
int ssdc_convert_coordinates(ssdc_ctx* ctx, const void* in, int dtype, int64_t rows, int width,
                             int start, int conversion, int border_pixels, double* out);
/* match_bipartite_greedy()  matching_utils.py:22-79: weights (m,n) -> matches (m,) */
int ssdc_match_bipartite_greedy(ssdc_ctx* ctx, const double* weights, int64_t m, int64_t n,
                                int64_t* out_matches);
/* match_multi()  matching_utils.py:81-116: writes the k matches (ascending anchor
 * index) to out_gt / out_anchor (capacity n) and k to *n_matches. */
int ssdc_match_multi(ssdc_ctx* ctx, const double* weights, int64_t m, int64_t n, double threshold,
                     int64_t* out_gt, int64_t* out_anchor, int64_t* n_matches);

/* SSDLoss.compute_loss   keras_loss_function/keras_ssd_loss.py:98-211 (log_loss :78-96, smooth_L1_loss :53-76),
 * forward pass.  `y_true` (B, A, C+12) float32 or float64 (`dtype_true`; converted to float32 like the Keras
 * placeholder), `y_pred` (B, A, C+12) float32; both host pointers, or both device pointers on dev_slot 0 when
 * `on_device != 0` (e.g. the device-resident output of `ssdc_encode`).  Writes the per-image loss (B,) float32 to
 * the host array `out_loss`.  Hard negative mining as in the reference: the `min(max(neg_pos_ratio * n_positive,
 * n_neg_min), #non-zero negative losses)` largest negative classification losses of the whole batch are kept,
 * equal values in flat index order (tf.nn.top_k). */
int ssdc_ssd_loss(ssdc_ctx* ctx, const void* y_true, int dtype_true, const float* y_pred, int on_device,
                  int64_t B, int64_t A, int C, int neg_pos_ratio, int n_neg_min, double alpha, float* out_loss);

/* ---- evaluation --------------------------------------------
 * Evaluator.match_predictions  eval_utils/average_precision_evaluator.py:570-777.
 * Predictions of all classes back to back: class c (1..n_classes) owns
 * [pred_class_offsets[c], pred_class_offsets[c+1]) (pred_class_offsets has n_classes+2 entries,
 * entry 0 unused, entry 1 == 0); `pred_image` is the dense image index of each prediction, confidences
 * and boxes are float32 like the reference's structured array.  Ground truth rows
 * [class, xmin, ymin, xmax, ymax] of image i are [gt_image_offsets[i], gt_image_offsets[i+1]);
 * `gt_neutral` may be NULL.  Per class the predictions are ordered by descending confidence (stable:
 * ties keep their original order); outputs are in that order: `out_order` (index inside the class),
 * true / false positive flags and their running sums.  `only_first != 0` reproduces the reference's
 * `verbose=False` behaviour (only the best prediction of every class is evaluated, :692-696). */
int ssdc_voc_match(ssdc_ctx* ctx, const int32_t* pred_image, const float* pred_conf, const float* pred_box,
                   const int64_t* pred_class_offsets, int n_classes,
                   const double* gt, const uint8_t* gt_neutral, const int64_t* gt_image_offsets,
                   int64_t n_images, double iou_threshold, int border_pixels, int only_first,
                   int32_t* out_order, int32_t* out_tp, int32_t* out_fp, int32_t* out_ctp, int32_t* out_cfp);

/* ---- decoder -> evaluator glue, augmentation box checks -----------
 * apply_inverse_transforms   data_generator/object_detection_2d_misc_utils.py:22-73 for the inverters the
 * reference's own transformations return: Resize (object_detection_2d_geometric_ops.py:75-79: scale by
 * original / resized size, np.round to 0 decimals), the patch samplers (object_detection_2d_patch_sampling_ops.py
 * :316-320 translate; :577, :730 identity).  `rows` (n, width) float64 host array, modified in place; image i owns
 * rows [row_offsets[i], row_offsets[i+1]) and the inverter steps [step_offsets[i], step_offsets[i+1]) of `steps`,
 * three doubles each: kind (0 identity, 1 scale + round, 2 translate), value for the y columns, value for the x
 * columns; applied in order. */
int ssdc_inverse_transform_rows(ssdc_ctx* ctx, double* rows, int64_t n, int width, const int64_t* row_offsets, int64_t B,
                                const double* steps, const int64_t* step_offsets,
                                int xmin_col, int ymin_col, int xmax_col, int ymax_col);

/* The per-class result lists of Evaluator.predict_on_dataset (eval_utils/average_precision_evaluator.py:402-422)
 * straight from the device-resident rows of the last image-sweep decode (see ssdc_decode_results_dev): inverse
 * transforms as above (`steps` may be NULL), confidence rounded to `round_conf_decimals` decimals (< 0: not
 * rounded, the reference's default), coordinates rounded to one decimal, fields narrowed to float32 like the
 * reference's structured array (:668-675).  Records come in (image, row) order: out_image = index of the image in the
 * decoded batch, out_class, out_conf, out_box (xmin, ymin, xmax, ymax).  SSDC_ERR_CAPACITY if `capacity` records do not
 * suffice (B * top_k always does). */
int ssdc_results_for_evaluation(ssdc_ctx* ctx, const double* steps, const int64_t* step_offsets, int round_conf_decimals,
                                int32_t* out_image, int32_t* out_class, float* out_conf, float* out_box,
                                int64_t capacity, int64_t* n_out);

/* BoxFilter.__call__   data_generator/object_detection_2d_image_boxes_validation_utils.py:174-232 for the boxes of
 * one or many images: `boxes` (n, 4) xmin, ymin, xmax, ymax, `image_hw` (n, 2) height and width of the image each box
 * is checked against; criterion 0 'center_point', 1 'iou', 2 'area'.  out_keep[i] = 1 if box i meets every enabled
 * requirement. */
int ssdc_box_filter(ssdc_ctx* ctx, const double* boxes, const double* image_hw, int64_t n,
                    int check_degenerate, int check_min_area, int check_overlap, int criterion,
                    double lower, double upper, double min_area, int border_pixels, uint8_t* out_keep);

#ifdef __cplusplus
}
#endif
#endif  /* SSDCODEC_H_ */
