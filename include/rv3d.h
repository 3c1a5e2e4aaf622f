/*
 * rv3d.h -- C ABI of librv3d.so: the B200 (sm_100a) implementation of torchbox3d's
 * rasterize -> decode -> rotated-IoU / NMS path.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - the caller owns every buffer, including scratch; the library never allocates
 *     device memory behind the caller's back (the one exception is the sort's
 *     temporary storage, which is carved out of the caller's scratch), never changes
 *     the current device, and enqueues all work on the caller's stream;
 *   - functions return RV3D_OK (0) or a negative rv3d_status; no exceptions, no aborts;
 *   - counts that only the device knows stay in device memory: rv3d_nms reads the candidate
 *     count from the compaction counter and leaves the detection count on the device;
 *   - re-entrant, no global mutable state; one host thread per GPU.
 *
 * Each entry point names the reference interface it replaces (paths relative to
 * /root/reference).  The third-party natives the reference binds
 * (torch.ops.detectron2.nms_rotated, weighted_nms_ext.wnms_gpu,
 * mmcv ext_module.box_iou_rotated) are un-vendored; their call sites are cited.
 */
#ifndef RV3D_H_
#define RV3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RV3D_VERSION 100 /* 0.1.0 */

typedef void *rv3d_stream_t; /* cudaStream_t */

typedef enum {
  RV3D_OK = 0,
  RV3D_ERR_ARG = -1,         /* bad shape / null pointer / unsupported enum value */
  RV3D_ERR_ALIGN = -2,       /* pointer not aligned as documented */
  RV3D_ERR_SCRATCH = -3,     /* scratch_bytes smaller than rv3d_*_scratch_bytes() */
  RV3D_ERR_CUDA = -4,        /* a CUDA runtime call / launch failed */
  RV3D_ERR_KEYBITS = -5,     /* (sweep,class,score,candidate) does not fit a 64-bit sort key: split the batch */
  RV3D_ERR_CAPACITY = -6     /* an output buffer's capacity was exceeded (device-side flag) */
} rv3d_status;

int rv3d_version(void);
const char *rv3d_strerror(int status);

/* ------------------------------------------------------------------------------------
 * 1. Rasterization
 * replaces: math/range_view.py:14-44 build_range_view
 *           math/numpy/conversions.py:46-73 cart_to_sph, :9-43 build_range_view_coordinates,
 *           :106-128 z_buffer (numba)            [and the copies in converters/av2/utils.py]
 * ------------------------------------------------------------------------------------ */
#define RV3D_COL_LIBRARY 0   /* col = rint(W - az' - 1)   numpy/conversions.py:35      */
#define RV3D_COL_CONVERTER 1 /* col = W - rint(az')       converters/av2/utils.py:137  */
#define RV3D_COL_CONVERTER_UNIFORM 2 /* converter column + rows uniform in inclination over +-10 deg
                                        (build_uniform_inclination, utils.py:138-145); only
                                        rv3d_range_view_coordinates takes it                 */

typedef struct {
  int32_t batch;        /* B sweeps in one launch                                        */
  int32_t max_points;   /* stride, in points, between consecutive sweeps                 */
  int32_t height;       /* H = n_inclination_bins                                        */
  int32_t width;        /* W the z-buffer ravels with (z_buffer's `width`)               */
  int32_t azimuth_bins; /* n_azimuth_bins of the column formula (the reference's library
                           wrapper leaves this at 1800 whatever `width` is)              */
  int32_t num_lasers;   /* points with laser >= num_lasers are dropped (range_view.py:23);
                           also the length of laser_mapping                              */
  int32_t col_mode;     /* RV3D_COL_*                                                    */
  int32_t reserved;     /* must be 0.  (Arithmetic is always f64, the production configuration:
                           converters/av2/export.py:77-81 casts to Float64 and
                           datasets/argoverse/av2.py:162 passes an f64 offset.)            */
  double lidar_offset[3];
  double min_distance;  /* z_buffer's min_distance (1.0)                                 */
} rv3d_raster_params;

size_t rv3d_rasterize_scratch_bytes(const rv3d_raster_params *p);

/* points: (B, max_points, 4) f32 [x, y, z, intensity], 16-byte aligned.
 * laser: (B, max_points) u8.  n_points: (B,) i32.  laser_mapping: (num_lasers,) i32
 * (row = H - laser_mapping[laser] - 1).
 * image: (B, 7, H, W) f32 [azimuth, inclination, range, x, y, z, intensity]
 * (range_view.py:33).  winner: (B, H, W) i32 index of the point that owns the pixel,
 * -1 if empty; may be NULL. */
int rv3d_rasterize(const rv3d_raster_params *p, const float *points, const uint8_t *laser,
                   const int32_t *n_points, const int32_t *laser_mapping, float *image,
                   int32_t *winner, void *scratch, size_t scratch_bytes, rv3d_stream_t stream);

/* Generic nearest-return scatter == z_buffer(indices, distances, features, H, W, min_distance)
 * (numpy/conversions.py:106-128).  rows/cols (N,) i64; dist (N,) f64 or f32; feat (C,N) f64
 * or f32; image (C,H,W) f32; winner (H,W) i32 or NULL. */
size_t rv3d_zbuffer_scratch_bytes(int32_t height, int32_t width);
int rv3d_zbuffer(const int64_t *rows, const int64_t *cols, const void *dist, int32_t dist_is_f64,
                 const void *feat, int32_t feat_is_f64, int32_t channels, int64_t n, int32_t height,
                 int32_t width, double min_distance, float *image, int32_t *winner, void *scratch,
                 size_t scratch_bytes, rv3d_stream_t stream);

/* cart_to_sph (numpy/conversions.py:46-73): (N,3) f64 -> (N,3) f64 [az, inc, r]. */
int rv3d_cart_to_sph(const double *cart, double *sph, int64_t n, rv3d_stream_t stream);
/* build_range_view_coordinates (numpy/conversions.py:9-43 / converters/av2/utils.py:108-153):
 * sph (N,3) f64 is MUTATED in place exactly like the reference (az' = (az+pi)*W/tau). */
int rv3d_range_view_coordinates(double *sph, const int64_t *laser, const int64_t *laser_mapping,
                                int32_t n_mapping, int64_t n, int32_t n_inclination_bins,
                                int32_t n_azimuth_bins, int32_t col_mode, double *hybrid,
                                rv3d_stream_t stream);

/* Loader-side post-processing of a rasterized range image, fused:
 * feature / cart / mask assembly (prototype/loader.py:623-650: column selection, Waymo tanh(intensity),
 * mask = range > 0) and subsample_range_view (prototype/loader.py:792-815: features *= mask, pad
 * [pad, pad] along W in `circular` or `constant` mode, keep every x_stride-th column).
 * image (B,7,H,W) f32 [az,inc,range,x,y,z,intensity] -> features (B,F,H,Wo), cart (B,3,H,Wo) f32,
 * mask (B,1,H,Wo) u8, Wo = ceil((W + 2 pad) / x_stride). */
#define RV3D_PAD_CIRCULAR 0
#define RV3D_PAD_CONSTANT 1
typedef struct {
  int32_t batch, height, width;
  int32_t x_stride, pad, pad_mode;
  int32_t n_features;
  int32_t feature_channel[8]; /* channel of `image` feeding feature f (0..6) */
  int32_t tanh_channel;       /* image channel passed through tanh (Waymo intensity = 6), or -1 */
} rv3d_inputs_params;
int rv3d_range_view_inputs(const rv3d_inputs_params *p, const float *image, float *features, float *cart,
                           uint8_t *mask, rv3d_stream_t stream);
/* The same assembly fused into the rasterizer (raw sweeps -> network inputs in one scatter + one resolve pass; the
 * 7-plane range image is never materialised).  `p` as in rv3d_rasterize, `ip` as in rv3d_range_view_inputs with
 * ip->batch / height / width equal to p's; results are identical to rv3d_rasterize followed by rv3d_range_view_inputs. */
int rv3d_rasterize_inputs(const rv3d_raster_params *p, const rv3d_inputs_params *ip, const float *points,
                          const uint8_t *laser, const int32_t *n_points, const int32_t *laser_mapping,
                          float *features, float *cart, uint8_t *mask, void *scratch, size_t scratch_bytes,
                          rv3d_stream_t stream);
/* subsample_range_view alone (prototype/loader.py:792-815) on already assembled tensors:
 * range_view (B,C,H,W) f32, mask (B,1,H,W) u8, cart (B,3,H,W) f32 -> the same three, padded and strided. */
int rv3d_subsample_range_view(const float *range_view, const uint8_t *mask, const float *cart, int32_t batch,
                              int32_t channels, int32_t height, int32_t width, int32_t x_stride, int32_t pad,
                              int32_t pad_mode, float *out_range_view, uint8_t *out_mask, float *out_cart,
                              rv3d_stream_t stream);

/* ------------------------------------------------------------------------------------
 * 2. Decoding
 * replaces: math/ops/coding.py:110-144 decode_range_view (+ :79-107 egovehicle_from_azimuth)
 *           nn/decoders/range_decoder.py:127-156 sample_by_range, :49-77 the per-task body
 *           of RangeDecoder.decode, math/linalg/lie/SO3.py:122-134 yaw_to_quat
 * ------------------------------------------------------------------------------------ */
#define RV3D_F32 0
#define RV3D_F16 1
#define RV3D_BF16 2
#define RV3D_MAX_PARTITIONS 8
/* Width of the score field of the 64-bit sort keys [segment | ~order(score) | candidate]:
 * rv3d_decode_compact writes 31 bits (its scores, sigmoid * mask, are never negative, so the top bit of the
 * order-preserving code is constant -- one radix pass less downstream); rv3d_compact_candidates, whose scores are
 * the caller's, writes 32.  rv3d_nms / rv3d_pack_candidates are told which through `score_bits`. */
#define RV3D_SCORE_BITS_DECODE 31

/* decode_range_view: regressands (B,8,H,W) in `dtype`, cart (B,3,H,W) in `cart_dtype` -> out (B,7,H,W) in
 * `dtype` (coding.py:126: the result takes the regressands' dtype); arithmetic in f64 (coding.py:127-128),
 * one cast at the end (:144).  Supported pairs: cart_dtype == dtype, or RV3D_F32 cart with F16 / BF16
 * regressands (the autocast case, nn/arch/detector.py:329-333); anything else is RV3D_ERR_ARG. */
int rv3d_decode_range_view(const void *regressands, const void *cart, void *out, int32_t dtype,
                           int32_t cart_dtype, int32_t batch, int32_t height, int32_t width,
                           int32_t azimuth_invariant, rv3d_stream_t stream);

typedef struct {
  int32_t n_partitions;                 /* 0 = sample_by_range disabled (BCHW_to_BKC) */
  float lower[RV3D_MAX_PARTITIONS];
  float upper[RV3D_MAX_PARTITIONS];
  int32_t rate[RV3D_MAX_PARTITIONS];
} rv3d_partitions;

/* number of candidates per sweep: sum_i H*ceil(W/rate_i), or H*W when disabled */
int64_t rv3d_num_candidates(const rv3d_partitions *parts, int32_t height, int32_t width);

/* sample_by_range: scores (B,1,H,W) f32, categories (B,1,H,W) i64, cuboids (B,7,H,W) f32,
 * cart (B,3,H,W) f32 -> out_scores (B,K) f32, out_categories (B,K) i64, out_cuboids (B,K,7) f32. */
int rv3d_sample_by_range(const float *scores, const int64_t *categories, const float *cuboids,
                         const float *cart, const rv3d_partitions *parts, int32_t batch, int32_t height,
                         int32_t width, float *out_scores, int64_t *out_categories, float *out_cuboids,
                         rv3d_stream_t stream);

typedef struct {
  int32_t batch, n_classes, height, width;
  int32_t dtype;             /* RV3D_F32 / F16 / BF16 of logits and regressands             */
  int32_t cart_dtype;        /* dtype of cart: == dtype, or RV3D_F32 under autocast         */
  int32_t azimuth_invariant; /* enable_azimuth_invariant_targets                            */
  int32_t category_offset;   /* task_offset (range_decoder.py:77)                           */
  int32_t candidate_offset;  /* index of this (stride, task)'s first candidate in the
                                concatenated K axis (range_decoder.py:88-93)                */
  int32_t total_candidates;  /* K of the concatenated axis (key packing)                    */
  int32_t total_classes;     /* classes over all tasks (key packing)                        */
  int32_t capacity;          /* rows available in out_keys / out_boxes                      */
  float min_confidence;      /* compared as float32, like torch does                        */
  rv3d_partitions parts;
} rv3d_decode_params;

/* Fused: sigmoid * mask, max over classes, threshold, decode of the survivors only,
 * range-partition subsampling, warp-aggregated compaction.
 * logits (B,C,H,W), regressands (B,8,H,W) in `dtype`, cart (B,3,H,W) in `cart_dtype`; mask (B,1,H,W) u8/bool.
 * Scores, threshold and boxes are rounded to `dtype`, the range partition (cart.norm) to `cart_dtype`.
 * Survivor r gets: out_keys[r] = sort key (sweep, class | score desc | candidate asc),
 * out_boxes[r] = 8 f32 [x,y,z,l,w,h,yaw,score].  *counter (device i32) is advanced
 * atomically, so several (stride, task) calls append to the same arrays; rows past
 * `capacity` are dropped and counted (the caller checks *counter <= capacity). */
int rv3d_decode_compact(const rv3d_decode_params *p, const void *logits, const void *regressands,
                        const void *cart, const uint8_t *mask, uint64_t *out_keys, float *out_boxes,
                        int32_t *counter, rv3d_stream_t stream);

/* Same compaction for already-dense candidates (the input of batched_multiclass_nms,
 * math/ops/nms.py:181-190): cuboids (B,K,7) f32, scores (B,K) f32, categories (B,K) i64. */
int rv3d_compact_candidates(const float *cuboids, const float *scores, const int64_t *categories,
                            int32_t batch, int32_t k, int32_t total_classes, float min_confidence,
                            int32_t apply_threshold, int32_t capacity, uint64_t *out_keys,
                            float *out_boxes, int32_t *counter, rv3d_stream_t stream);

/* ------------------------------------------------------------------------------------
 * 3. Suppression
 * replaces: math/ops/nms.py:181-266 batched_multiclass_nms, :11-61 hard_multiclass_nms
 *           (detectron2 nms_rotated, :41-45), :64-123 weighted_multiclass_nms,
 *           :126-177 weighted_nms (TorchEx wnms_gpu, :161-170),
 *           math/ops/iou.py:11-47 iou_3d_axis_aligned (mmcv box_iou_rotated, :15)
 * ------------------------------------------------------------------------------------ */
#define RV3D_NMS_HARD 0
#define RV3D_NMS_WEIGHTED 1
#define RV3D_OUT_QUAT 0
#define RV3D_OUT_YAW 1
#define RV3D_MAX_PEERS 8

typedef struct {
  int32_t batch, total_classes, total_candidates;
  int32_t num_pre_nms, num_post_nms;
  int32_t mode;              /* RV3D_NMS_*                                              */
  float iou_threshold;       /* float32(iou_threshold): nms.py:44 passes an f32 tensor  */
  float merge_threshold;     /* weighted only (0.5, nms.py:106)                         */
  int32_t capacity;          /* rows available in keys / boxes; the LIVE count is read from device memory
                                (n_candidates) and clamped to this -- no host copy of it is needed      */
  int32_t out_capacity;      /* rows available in the out_* arrays                      */
  int32_t out_layout;        /* RV3D_OUT_QUAT: out_params (cap,10) [x,y,z,l,w,h,qw,qx,qy,qz]
                                (RangeDecoder.decode); RV3D_OUT_YAW: out_params (cap,7)
                                [x,y,z,l,w,h,yaw] (batched_multiclass_nms)               */
  int32_t score_bits;        /* score field of the keys: 31 (rv3d_decode_compact) or 32
                                (rv3d_compact_candidates); 0 means 32                     */
  float score_lo, score_hi;  /* range the scores are known to lie in (decode: [min_confidence, 1]); only balances the
                                score bins, any scores stay correct.  score_hi <= score_lo: unknown, the library
                                reduces min / max on the device first                     */
  int32_t flags;             /* must be 0 (every pair that survives the pruning runs the bit-exact IoU routine;
                                there is no approximate mode)                              */
  /* Fused detection gather over peer memory (the path's one exchange step, multi-GPU; replaces the per-sweep
   * feather files + dist.barrier() of nn/arch/detector.py:366-380,415-421).  peer_world > 0: the pack kernel
   * ALSO stores every detection as a 16-float row [sweep + sweep_offset, class, score, 0, x,y,z,l, w,h,qw,qx,
   * qy,qz,0,0] into slot `peer_rank` of each rank's (peer_world, peer_capacity + 1, 16) f32 buffer, row 0 of
   * the slot being [rows written, rows kept, seq, 0], with plain 16-byte stores through peer-mapped (NVLink)
   * pointers.  peer_seq == NULL: the caller picked the slot (peer_rows point at it) and orders readers behind the
   * writers itself (a device-side barrier over the ranks).  peer_seq != NULL (device u32, this rank's count of
   * published steps): step k = *peer_seq + 1 is written to slot k & 1 at peer_rows[q] + slot * peer_slot_stride
   * floats; the last pack block to finish publishes the header with k in its third word behind a system-scope
   * fence and sets *peer_seq = k.  Readers wait with rv3d_peer_wait -- no barrier kernel, and since slot and
   * sequence number are read from device memory the step replays unchanged inside a CUDA graph.  The caller waits
   * for step k - 1 (rv3d_peer_wait) before enqueueing step k's rv3d_nms and consumes a step's rows on the same
   * stream before the next one: then no writer overtakes a reader of the slot it reuses.
   * RV3D_OUT_QUAT only.  peer_world == 0: off. */
  int32_t peer_world, peer_rank, peer_capacity, sweep_offset;
  int32_t reserved;
  float *peer_rows[RV3D_MAX_PEERS];
  uint32_t *peer_seq;
  int64_t peer_slot_stride;
  int32_t *host_count;       /* optional HOST pointer (mapped pinned memory, device-accessible): the detection count
                                is also stored there, so a caller waiting on an event can read it without a copy */
} rv3d_nms_params;

size_t rv3d_nms_scratch_bytes(const rv3d_nms_params *p);

/* keys / boxes: the compaction output (left untouched); n_candidates: DEVICE i32, the compaction counter.
 * Outputs, in the reference's order (sweep asc, class asc, score desc):
 *   out_params (out_capacity,10|7) f32, see out_layout          (range_decoder.py:122-123)
 *   out_scores, out_categories, out_batch (out_capacity,) f32      (nms.py:51,113,242)
 *   out_count device i32: rows written.
 * Nothing in this call depends on a host copy of a device count: fixed launch geometry, safe under CUDA-graph
 * capture (replaces the per-sweep host sync of nms.py:210-215).
 * stats (device, 24 x i64, may be NULL): [0] exact rotated-IoU evaluations, [1] kept, [2] frontier rounds,
 * [3] circle tests, [4..9] SM cycles summed over segments per phase (window sort, pull walk, pull IoU, frontier
 * pairs, greedy, publish), [10] slowest segment (cycles), [11] largest segment, [12..17] its phases, [18] pairs
 * above the threshold, [19] upper-bound tests, [20] candidates consumed by the scan, [21..23] sub-phases (frontier walk,
 * frontier evaluation, pull walk + scan; subsets of [7], [7], [5]). */
int rv3d_nms(const rv3d_nms_params *p, const uint64_t *keys, const float *boxes, const int32_t *n_candidates,
             float *out_params, float *out_scores, float *out_categories, float *out_batch, int32_t *out_count,
             int64_t *stats, void *scratch, size_t scratch_bytes, rv3d_stream_t stream);

/* Consumer side of the peer_seq protocol: with k = *seq (device; the steps this rank has published), enqueue a wait
 * until every rank's header in slot k & 1 of THIS rank's buffer `rows` ((2, world, peer_capacity + 1, 16) f32, slots
 * `slot_stride` floats apart) carries a sequence number >= k. */
int rv3d_peer_wait(const float *rows, int64_t slot_stride, int32_t world, int32_t peer_capacity, const uint32_t *seq,
                   rv3d_stream_t stream);

/* detectron2-style entry: boxes (N,5) f32 (xc,yc,w,h,angle_deg), scores (N,) f32 ->
 * keep (N,) i64 original indices in score order, *n_keep device i32.
 * replaces torch.ops.detectron2.nms_rotated (call site nms.py:41-45). */
size_t rv3d_nms_rotated_scratch_bytes(int32_t n);
int rv3d_nms_rotated(const float *boxes, const float *scores, int32_t n, float iou_threshold,
                     int64_t *keep, int32_t *n_keep, void *scratch, size_t scratch_bytes,
                     rv3d_stream_t stream);

/* TorchEx-style entry: boxes (N,5) f32 (x1,y1,x2,y2,ry) and data (N,D) f32 (score LAST),
 * both already sorted by score descending (nms.py:148-154).  output (N,D) f32, keep (N,) i64,
 * count (N,) i64, *n_out device i32.  replaces weighted_nms_ext.wnms_gpu (nms.py:161-170). */
size_t rv3d_wnms_scratch_bytes(int32_t n, int32_t d);
int rv3d_wnms(const float *boxes, const float *data, int32_t n, int32_t d, float nms_threshold,
              float merge_threshold, float *output, int64_t *keep, int64_t *count, int32_t *n_out,
              void *scratch, size_t scratch_bytes, rv3d_stream_t stream);

/* Aligned rotated BEV IoU + axis-aligned 3D IoU of (N,7) f32 cuboids (iou.py:11-47).
 * status (device i32): set to 1 if any 3D IoU is non-finite (the reference raises). */
int rv3d_iou3d_aligned(const float *cuboids_a, const float *cuboids_b, int64_t n, float *iou3d,
                       float *iou_bev, int32_t *status, rv3d_stream_t stream);

/* mmcv.ops.box_iou_rotated(bboxes1, bboxes2, aligned) (call sites: math/ops/assignment.py:24-26,67-69 aligned,
 * prototype/loader.py:785-788 all pairs): boxes (N,5) / (M,5) f32 (xc, yc, w, h, angle in radians) ->
 * out (N,M) f32 row-major, or (N,) when `aligned` (then n == m).  Same arithmetic as the NMS routine. */
int rv3d_box_iou_rotated(const float *boxes_a, int64_t n, const float *boxes_b, int64_t m, int32_t aligned,
                         float *out, rv3d_stream_t stream);

/* Test hook for the pruning the NMS kernels apply in front of the bit-exact IoU routine: aligned pairs of (N,5) f32
 * boxes -- routine 0: (xc, yc, w, h, angle in degrees), the hard mode's detectron2-style routine; routine 1:
 * (x1, y1, x2, y2, ry), the weighted mode's iou_bev -- -> decision (N,) i8 (2: stopped by the IoU upper bound, i.e.
 * treated as "not above the threshold"; 0: sent to the exact routine), bound (N,) f32 (that upper bound), exact (N,)
 * f32 (the bit-exact routine's value). */
int rv3d_pair_decisions(const float *boxes_a, const float *boxes_b, int64_t n, float iou_threshold, int32_t routine,
                        int8_t *decision, float *bound, float *exact, rv3d_stream_t stream);

/* Test hooks for the fast math the rasterizer / decoder use in place of libm (csrc/fastmath.cuh): out[i] =
 * op 0: fast_atan2(a[i], b[i]); 1: fast_exp(a[i]); 2: fast_sqrt(a[i]); 3: (double)atan2f_lite((float)a[i], (float)b[i]).
 * rv3d_debug_column: per point (n, 4) f32, col_fast = the float32 azimuth-bin decision of the scatter kernel, or
 * -1 - c when it declined and the fp64 fallback answered c; col_exact = the reference's arithmetic with libm atan2
 * (numpy/conversions.py:30-35 / converters/av2/utils.py:133-137).  Every decided point must agree. */
int rv3d_debug_fastmath(int32_t op, const double *a, const double *b, double *out, int64_t n, rv3d_stream_t stream);
int rv3d_debug_column(const rv3d_raster_params *p, const float *points, int64_t n, int32_t *col_fast,
                      int32_t *col_exact, rv3d_stream_t stream);

/* yaw (N,) f32 -> quat (N,4) f32 (qw,qx,qy,qz) = (cos(yaw/2),0,0,sin(yaw/2)) (SO3.py:122-134). */
int rv3d_yaw_to_quat(const float *yaw, float *quat, int64_t n, rv3d_stream_t stream);

/* Threshold-only branch of RangeDecoder.decode (use_nms=False, range_decoder.py:110-120):
 * compaction output -> rows ordered by (sweep, candidate). */
size_t rv3d_pack_candidates_scratch_bytes(int32_t n_candidates);
int rv3d_pack_candidates(const uint64_t *keys, const float *boxes, int32_t n_candidates, int32_t batch,
                         int32_t total_classes, int32_t total_candidates, int32_t score_bits, float *out_params,
                         float *out_scores, int64_t *out_categories, int64_t *out_batch, void *scratch,
                         size_t scratch_bytes, rv3d_stream_t stream);

/* ------------------------------------------------------------------------------------
 * 4. Sweep preparation in front of the rasterizer -- fp64 like the numpy code
 * replaces: converters/av2/utils.py:229-295 unmotion_compensate, :43-57 sensor_SE3_egovehicle
 *           (av2 SE3.inverse().transform_point_cloud), :211-226 correct_laser_numbers
 * ------------------------------------------------------------------------------------ */

/* unmotion_compensate: xyz (N,3) f64, offset_ns (N,) i64 (device).  Pose table (device, sorted by time):
 * pose_timestamps_ns (M,) i64, pose_quat_xyzw (M,4) f64 scalar-last as scipy's from_quat takes the
 * reference's (qx,qy,qz,qw) columns, pose_translation (M,3) f64.  target_* (HOST, 4 + 3 doubles): the pose
 * whose timestamp equals `timestamp_ns` (utils.py:258-273; the caller looks it up and fails like the
 * reference when it is absent).  Per point: t = timestamp_ns + offset; rows outside (min, max) pose time
 * are dropped by the reference -> out_valid = 0 and NaN coordinates here (the caller compacts);
 * rotation = scipy Slerp on float64 timestamps, translation = t_low*alpha + (1-alpha)*t_high (:275-276,
 * weights as in the reference); out_xyz = inv(city_SE3_laser) @ city_SE3_roll @ [x y z 1]. */
int rv3d_unmotion_compensate(const double *xyz, const int64_t *offset_ns, int64_t n, int64_t timestamp_ns,
                             const int64_t *pose_timestamps_ns, const double *pose_quat_xyzw,
                             const double *pose_translation, int32_t n_poses, const double *target_quat_xyzw,
                             const double *target_translation, double *out_xyz, uint8_t *out_valid,
                             rv3d_stream_t stream);

/* Table form of the same operator for the production shape (one log = one pose table, many sweeps).
 * rv3d_pose_intervals: once per table -> intervals (M-1, 8) f64 device = per pose pair the normalised lower
 * quaternion (4) and the rotation vector of q0^-1 q1 (3) + 1 pad, i.e. what scipy's Slerp.__init__ precomputes
 * (utils.py:251-256 builds the Slerp once per call).  rv3d_unmotion_compensate_table: same result as
 * rv3d_unmotion_compensate; first_ns / last_ns = pose_timestamps_ns[0] / [M-1] (host copies, the caller owns the
 * table); *n_dropped (device i32, may be NULL, caller zeroes it) counts the rows the reference's filter removes, so
 * the caller can skip the compaction when it is 0. */
int rv3d_pose_intervals(const double *pose_quat_xyzw, int32_t n_poses, double *intervals, rv3d_stream_t stream);
int rv3d_unmotion_compensate_table(const double *xyz, const int64_t *offset_ns, int64_t n, int64_t timestamp_ns,
                                   const int64_t *pose_timestamps_ns, const double *pose_translation,
                                   const double *intervals, int32_t n_poses, int64_t first_ns, int64_t last_ns,
                                   const double *target_quat_xyzw, const double *target_translation,
                                   double *out_xyz, uint8_t *out_valid, int32_t *n_dropped, rv3d_stream_t stream);

/* Rigid transform of (N,3) f64 points: out = xyz @ R^T + t, or with `inverse` the transform of
 * SE3(R, t).inverse() (R^T, R^T.(-t)) as utils.py:54-57 builds sensor_SE3_egovehicle.
 * rotation (9, row-major) and translation (3): HOST doubles. */
int rv3d_transform_points(const double *xyz, int64_t n, const double *rotation, const double *translation,
                          int32_t inverse, double *out, rv3d_stream_t stream);

/* correct_laser_numbers: laser_numbers (N,) i64 -> out (N,) i64 = row_mapping[remap(laser)].
 * laser_mapping (32,) i64 device or NULL (log not in LOG_IDS): l >= 32 -> laser_mapping[l-32]+32,
 * l < 32 -> laser_mapping[l].  row_mapping (n_rows,) i64 = ROW_MAPPING_32 / ROW_MAPPING_64.
 * *out_of_range (device i32, may be NULL) is set to 1 where numpy would raise IndexError. */
int rv3d_correct_laser_numbers(const int64_t *laser_numbers, int64_t n, const int64_t *laser_mapping,
                               const int64_t *row_mapping, int32_t n_rows, int64_t *out,
                               int32_t *out_of_range, rv3d_stream_t stream);

/* ------------------------------------------------------------------------------------
 * 5. Training-time callers of the same operators
 * replaces: the per-instance loop of compute_classification_targets, math/ops/assignment.py:121-139
 *           (per panoptic instance: topk(min(k, n)) of the pixel affinities, everything else zeroed)
 * ------------------------------------------------------------------------------------ */

/* affinity (N,) f32 and segment (N,) i32 in [0, n_segments) per foreground pixel, pixels of one instance in the
 * reference's masked_select (raster) order -> likelihood (N,) f32: the affinity where the pixel is among the k
 * largest of its segment (ties: earlier pixel first; NaN ranks highest like torch.topk), else 0. */
size_t rv3d_instance_topk_scratch_bytes(int64_t n);
int rv3d_instance_topk(const float *affinity, const int32_t *segment, int64_t n, int32_t n_segments, int32_t k,
                       float *likelihood, void *scratch, size_t scratch_bytes, rv3d_stream_t stream);

/* compute_classification_targets as ONE call (math/ops/assignment.py:76-148), float32 tensors:
 * input / target (B,8,H,W), labels (B,H,W) i64 in [0, n_classes] (n_classes = background_index), cart (B,3,H,W),
 * mask (B,1,H,W) bool bytes, panoptics (B,H,W) i64 (0 = background) -> affinities (B,n_classes,H,W) f32,
 * foreground (B,1,H,W) f32, background / reg_weights (B,1,H,W) bool bytes.  affinity_fn: 0 = BEV (aligned rotated
 * IoU of the decoded boxes, clamped, :64-73), 1 = GAUSSIAN without normalisation (exp(-|centre distance| / sigma2),
 * :151-161).  k: slots per instance (<= 64), 0 = nothing is foreground, RV3D_TOPK_ALL = every pixel of an instance is
 * in its top-k (the production setting k = inf, conf/model/range_view.yaml:126).  id_capacity: instance ids per sweep
 * the call can hold; *status (device i32, caller zeroes it) is set to 1 when an id >= id_capacity was met (the results
 * are then incomplete and the caller must retry with a larger capacity); with status = NULL such an id traps the kernel
 * (a sticky launch failure: loud, and no host read on the good path).  Only foreground pixels are decoded; no
 * dense intermediate is written. */
#define RV3D_TOPK_ALL 0x7fffffff
size_t rv3d_classification_targets_scratch_bytes(int32_t batch, int32_t height, int32_t width, int32_t k,
                                                 int32_t id_capacity);
int rv3d_classification_targets(const float *input, const float *target, const int64_t *labels, const float *cart,
                                const uint8_t *mask, const int64_t *panoptics, int32_t batch, int32_t n_classes,
                                int32_t height, int32_t width, int32_t affinity_fn, int32_t az_inv_targets, int32_t k,
                                float sigma2, int32_t id_capacity, float *affinities, float *foreground,
                                uint8_t *background, uint8_t *reg_weights, int32_t *status, void *scratch,
                                size_t scratch_bytes, rv3d_stream_t stream);

/* ------------------------------------------------------------------------------------
 * 6. Detection wire format
 * replaces: math/ops/coding.py:31-58 build_dataframe's per-column .tolist() reads,
 *           nn/arch/detector.py:45-60 SERIALIZED_SCHEMA, :573-584 the evaluation range filter
 * ------------------------------------------------------------------------------------ */
typedef struct {
  float params[10];        /* tx_m ty_m tz_m length_m width_m height_m qw qx qy qz */
  float score;
  int32_t category_index;
  int64_t timestamp_ns;    /* of the detection's sweep (uuids joined on batch_index, coding.py:69) */
  int32_t batch_index;
  float range_m;           /* float32 ||(tx,ty,tz)||, the quantity detector.py:573-576 filters on */
} rv3d_detection_record;   /* 64 bytes */

/* The decoder's outputs (params (N,10), scores / categories / batch_index (N,) f32 as RangeDecoder.decode returns
 * them) -> out (<= N records, order preserved), *out_count (device i32).  sweep_timestamp_ns (batch,) i64 device or
 * NULL.  apply_range_filter: keep range_m <= max_range_m only (detector.py:580). */
size_t rv3d_detection_records_scratch_bytes(int64_t n);
int rv3d_detection_records(const float *params, const float *scores, const float *categories,
                           const float *batch_index, int64_t n, const int64_t *sweep_timestamp_ns, int32_t batch,
                           float max_range_m, int32_t apply_range_filter, rv3d_detection_record *out,
                           int32_t *out_count, void *scratch, size_t scratch_bytes, rv3d_stream_t stream);

/* prepare_for_evaluation's `.sort(col("score"), descending=True).unique()` (detector.py:581-584) on a record stream
 * that is still on the device: records (capacity,) with *count (device i32) live rows -> out: the distinct rows (all
 * 64 bytes compared) in the order (score descending, then a hash of the row, then input order), *out_count (device
 * i32).  polars leaves the order after unique() unspecified; the set of rows is what the evaluation consumes. */
size_t rv3d_records_sort_unique_scratch_bytes(int64_t capacity);
int rv3d_records_sort_unique(const rv3d_detection_record *records, const int32_t *count, int64_t capacity,
                             rv3d_detection_record *out, int32_t *out_count, void *scratch, size_t scratch_bytes,
                             rv3d_stream_t stream);
/* validation_step's `dts.group_by(["log_id", "timestamp_ns"], maintain_order=True)` (detector.py:366-380) on a record
 * stream in the decoder's order (batch_index ascending): offsets (batch + 1,) i32, sweep b's records are
 * [offsets[b], offsets[b + 1]). */
int rv3d_records_group_offsets(const rv3d_detection_record *records, const int32_t *count, int32_t batch,
                               int32_t *offsets, rv3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RV3D_H_ */
