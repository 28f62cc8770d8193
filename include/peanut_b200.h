/*
 * peanut_b200 C-ABI: the drop-in boundary for PEANUT's per-step perception hot path on B200 (sm_100a).
 *
 * The reference (ajzhai/PEANUT) has no FFI layer: the boundary is three Python call sites inside
 * nav/agent.  The Python shims in peanut_b200/ keep those call signatures and bind the
 * entry points below through ctypes (INTEGRATION.md shows the stub).  Conventions:
 *   - every function returns 0 on success, non-zero on failure; pn_last_error() gives the message
 *     (thread-local, valid until the next failing call on that thread);
 *   - pointers named *_dev are CUDA device pointers on the context's device, *_host are host pointers
 *     (pinned host memory makes the copies asynchronous); the caller owns every buffer;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); *_dev entry points
 *     enqueue work and return without synchronising, *_host entry points synchronise before returning;
 *   - tensors are dense, row-major, fp32 unless stated otherwise, in the reference's own layouts (NCHW).
 *   - there is no CPU fallback: every entry point fails if no sm_100 device is present.
 */
#ifndef PEANUT_B200_H_
#define PEANUT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pn_ctx pn_ctx;

/* Arithmetic of the tensor-core path. */
enum pn_precision {
  PN_BF16 = 0, /* bf16 operands, fp32 accumulate (throughput path) */
  PN_TF32 = 1, /* fp32 storage, tf32 operands, fp32 accumulate (fp32-parity path, the drop-in shims' default) */
  PN_FP32 = 2  /* fp32 storage with every stored bit kept, each convolution as three tf32 tensor-core products
                  (hi*hi + hi*lo + lo*hi, fp32 accumulate): ~1e-6 relative, about a third of PN_TF32's speed and three
                  times its activation memory.  The strict-parity mode: what the reference's fp32 CPU / cuDNN-fp32
                  arithmetic is compared with (segmentation.py:45, prediction.py:128-131 run in fp32). */
};

int pn_create(int device, pn_ctx** out);
int pn_destroy(pn_ctx* ctx);
const char* pn_last_error(void);
/* ABI version of this header; bumped on any signature change. */
int pn_abi_version(void);

/*
 * Weights are handed over by checkpoint key, fp32, exactly as stored in the reference checkpoints:
 *   - prediction net: mmcv state_dict keys loaded by init_segmentor (prediction/mmseg/apis/inference.py:33-35),
 *     e.g. "backbone.stem.0.weight", "backbone.layer1.0.bn1.running_var", "decode_head.conv_seg.bias";
 *   - Mask-RCNN: detectron2 checkpoint keys loaded by DefaultPredictor (nav/agent/utils/segmentation.py:38),
 *     e.g. "backbone.bottom_up.res2.0.conv1.weight", "roi_heads.box_predictor.cls_score.bias".
 * BatchNorm folding and re-layout happen in the *_build calls.
 */
int pn_set_weight(pn_ctx* ctx, const char* name, const float* data_host, int ndim, const int64_t* shape);
int pn_clear_weights(pn_ctx* ctx);

/* ---- Stage C: map-completion encoder-decoder.
 * Replaces run_inference()/inference_segmentor() + scipy expit:
 *   nav/agent/prediction.py:112-137, :155-158; prediction/mmseg/apis/inference.py:70-99;
 *   prediction/mmseg/models/segmentors/encoder_decoder.py:260-271.
 * in : partial map  [B, C, H, W]            (reference: C=14, H=W=720)
 * out: class logits [B, num_classes, H, W]  (probabilities when apply_sigmoid != 0)            */
int pn_prednet_build(pn_ctx* ctx, int B, int C, int H, int W, int num_classes, int precision);
int pn_prednet_forward(pn_ctx* ctx, const float* map_dev, int apply_sigmoid, float* out_dev, void* stream);
int pn_prednet_forward_host(pn_ctx* ctx, const float* map_host, int apply_sigmoid, float* out_host);
/* Number of kernel launches one forward pass enqueues (for bench.py's gpu_launches). */
int pn_prednet_num_launches(pn_ctx* ctx);
/* Debug/parity taps: copies an intermediate tensor to fp32 NCHW.  which: 0 = layer4 features
 * [B,2048,H/8,W/8], 1 = low-resolution logits [B,num_classes,H/8,W/8]. */
int pn_prednet_read_tap(pn_ctx* ctx, int which, float* out_dev, void* stream);

/* Algorithmic FLOPs (2*MAC over conv/FC, unpadded) of one forward pass of the built network. */
int pn_prednet_flops(pn_ctx* ctx, double* flops_out);

/* Per-launch device time of one forward pass of a built network (CUDA events around every recorded launch,
 * eager mode, averaged over `iters` passes after one warm-up; run one regular forward first so that the
 * per-call buffers are set).  ms_out[max_ops] receives milliseconds per op, flops_out[max_ops] (may be NULL) the
 * algorithmic FLOPs of each op (0 for non-GEMM ops; mask-head ops are counted at full capacity),
 * names_out a '\n'-separated list of op names.  Measurement aid for bench.py / profiles/. */
enum pn_net { PN_NET_PREDNET = 0, PN_NET_MASKRCNN = 1 };
int pn_net_num_ops(pn_ctx* ctx, int which);
int pn_net_profile(pn_ctx* ctx, int which, int iters, float* ms_out, double* flops_out, int max_ops, char* names_out,
                   int names_bytes);

/* ---- Stage A: Mask-RCNN R101-FPN RGB -> per-category mask stack.
 * Replaces SemanticPredMaskRCNN.get_prediction, nav/agent/utils/segmentation.py:41-62 (call site
 * nav/agent/agent_helper.py:220-225), i.e. detectron2's DefaultPredictor + the per-instance accumulation loop.
 * The configuration mirrors the keys of nav/agent/utils/COCO-InstSeg/mask_rcnn_R_101_cat9.yaml that shape
 * inference (NULL = the reference's values). */
typedef struct pn_maskrcnn_cfg {
  int min_size_test;        /* 800   INPUT.MIN_SIZE_TEST            yaml:30  */
  int max_size_test;        /* 1333  INPUT.MAX_SIZE_TEST            yaml:28  */
  int rpn_pre_nms_topk;     /* 1000  RPN.PRE_NMS_TOPK_TEST          yaml:251 */
  int rpn_post_nms_topk;    /* 1000  RPN.POST_NMS_TOPK_TEST         yaml:249 */
  float rpn_nms_thresh;     /* 0.7   RPN.NMS_THRESH                 yaml:247 */
  int num_classes;          /* 9     ROI_HEADS.NUM_CLASSES          yaml:193 */
  float box_nms_thresh;     /* 0.5   ROI_HEADS.NMS_THRESH_TEST      yaml:192 */
  int detections_per_image; /* 100   TEST.DETECTIONS_PER_IMAGE      yaml:312 */
  float mask_threshold;     /* 0.5   paste_masks_in_image threshold          */
} pn_maskrcnn_cfg;

/* B frames of H x W per forward (reference: 1 x 480 x 640). */
int pn_maskrcnn_build(pn_ctx* ctx, int B, int H, int W, int precision, const pn_maskrcnn_cfg* cfg);
/* rgb [B,H,W,3] uint8 RGB (the flip to BGR of segmentation.py:44 happens inside), goal_cat [B] int32 or NULL,
 * score_thresh = ROI_HEADS.SCORE_THRESH_TEST (segmentation.py:33 sets it to args.sem_pred_prob_thr),
 * sem_pred_prob_thr / goal_thr = the gates of segmentation.py:54-58,
 * sem_out [B,H,W,num_classes+1] fp32: per-category sum of instance masks (last channel always 0). */
int pn_maskrcnn_forward(pn_ctx* ctx, const uint8_t* rgb_dev, const int* goal_cat_dev, float score_thresh,
                        float sem_pred_prob_thr, float goal_thr, float* sem_out_dev, void* stream);
int pn_maskrcnn_forward_host(pn_ctx* ctx, const uint8_t* rgb_host, const int* goal_cat_host, float score_thresh,
                             float sem_pred_prob_thr, float goal_thr, float* sem_out_host);
int pn_maskrcnn_num_launches(pn_ctx* ctx);
/* Network input geometry: resized_hw = ResizeShortestEdge output, padded_hw = padded to a multiple of 32. */
int pn_maskrcnn_input_size(pn_ctx* ctx, int* resized_hw, int* padded_hw);
/* Parity aids.  Stages, in order: preprocess, backbone, fpn, rpn_head, rpn_proposals, box_head, detections,
 * mask_head, paste ("end" = past the last).  pn_maskrcnn_set_call stores the per-call pointers / thresholds,
 * pn_maskrcnn_run_stages runs [first_stage, end_stage) eagerly, pn_maskrcnn_tap copies a named intermediate out
 * (write = 0) or overwrites it (write = 1): NHWC activations ("input", "res2".."res5", "p2".."p6", "box_pooled",
 * "mask_pooled") travel as fp32 NCHW with `channels` channels; raw buffers ("resized_u8", "rpn_head.p2".."p6",
 * "prop_boxes", "prop_scores", "prop_img", "prop_count", "box_out", "det_boxes", "det_scores", "det_classes",
 * "det_count", "mask_logits") as stored. */
int pn_maskrcnn_set_call(pn_ctx* ctx, const uint8_t* rgb_dev, const int* goal_cat_dev, float score_thresh,
                         float sem_pred_prob_thr, float goal_thr, float* sem_out_dev, void* stream);
int pn_maskrcnn_run_stages(pn_ctx* ctx, const char* first_stage, const char* end_stage, void* stream);
int pn_maskrcnn_tap(pn_ctx* ctx, const char* name, int write, int channels, void* buf_dev, int64_t buf_bytes, void* stream);
/* Host-only: the fixed-point coefficient tables of Pillow's bilinear resize (what DefaultPredictor's
 * ResizeShortestEdge applies to the uint8 frame).  bounds_out [out_size][2] = (first input index, tap count),
 * coeffs_out [out_size][*ksize_out] 22-bit fixed point; *ksize_out: in = row width of coeffs_out, out = taps. */
int pn_pil_bilinear_coeffs(int in_size, int out_size, int* bounds_out, int* coeffs_out, int* ksize_out);

/* ---- Glue after stage C: the tail of Agent_State.update_prediction
 * (nav/agent/agent_state.py:357-372).  pred [num_classes, window, window] = stage C output for the prediction window whose
 * origin in the full map is (x1, y1) ((0, 0) when the window is the full map); the goal category's plane is embedded into
 * a zero canvas, cut to the local-map bounds rows [r0, r0+local_w) x columns [c0, c0+local_h), and cells whose explored
 * value (local_map[1], row stride in elements) is >= 0.5 are zeroed.  target_out [local_w, local_h] fp32. */
int pn_target_pred(pn_ctx* ctx, const float* pred_dev, int num_classes, int window, int x1, int y1, int goal_cat, int r0,
                   int c0, int local_w, int local_h, const float* explored_dev, int64_t explored_row_stride,
                   float* target_out_dev, void* stream);

/* ---- Map bookkeeping, batched over E environments with every field in DEVICE memory:
 * Agent_State.get_local_map_boundaries / init_map_and_pose / init_with_obs (stamp) / update_local_map (after the mapper
 * call) / update_full_map, nav/agent/agent_state.py:153-211, 116-122, 276-303, 308-338.  The reference derives every cell
 * index on the host (pose `.cpu().numpy()` at :276, :315, :335); here nothing synchronises. */
typedef struct pn_map_cfg {
  int num_channels;        /* nc = 4 + num_sem_categories           agent_state.py:39 */
  int full_w, full_h;      /* map_size_cm / map_resolution          :41-42 */
  int local_w, local_h;    /* full / global_downscaling             :43-44 */
  int map_resolution;      /* 5 (cm per cell) */
  int map_size_cm;         /* 4800 */
  int global_downscaling;  /* 2 */
  int grid_resolution;     /* 24   arguments.py:95 */
  int col_rad;             /* 4    arguments.py:86 (integer-valued); the explored disk has radius col_rad + 1 */
  float goal_reached_dist; /* 75   arguments.py:102 */
  int f64_cells;           /* 0: int(r * 100.0 / res) in float32 (numpy >= 2 scalar rules), 1: in float64 (numpy < 2) */
} pn_map_cfg;

typedef struct pn_map_arrays {
  float* full_map;             /* [E, nc, full_w, full_h] */
  float* local_map;            /* [E, nc, local_w, local_h]  (the mapper's map_out) */
  float* full_pose;            /* [E, 3] x m, y m, theta deg */
  float* local_pose;           /* [E, 3]  (the mapper's poses_inout) */
  double* origins;             /* [E, 3] */
  int* lmb;                    /* [E, 4] local-map bounds: row0, row1, col0, col1 */
  double* planner_pose_inputs; /* [E, 7] */
  int* loc;                    /* [E, 2] loc_r, loc_c */
  double* dist_to_goal;        /* [E] */
  const int* global_goal;      /* [E, 2] global_goals[0] (read by update_local) */
} pn_map_arrays;

/* init_map_and_pose (:180-211): zero full_map, pose = map centre, 3x3 stamp, first window, local map and pose. */
int pn_map_init(pn_ctx* ctx, const pn_map_cfg* cfg, const pn_map_arrays* arrays, int E, void* stream);
/* init_with_obs (:116-122): 3x3 stamp on local_map[2:4] at the cell of local_pose (after the first mapper call). */
int pn_map_stamp_initial(pn_ctx* ctx, const pn_map_cfg* cfg, const pn_map_arrays* arrays, int E, void* stream);
/* update_local_map after the mapper call (:276-303): planner pose, channel-2 reset, 5x5 trajectory stamp, explored disk
 * under the agent and (once dist_to_goal < goal_reached_dist) at the goal; writes loc and dist_to_goal. */
int pn_map_update_local(pn_ctx* ctx, const pn_map_cfg* cfg, const pn_map_arrays* arrays, int E, void* stream);
/* update_full_map (:308-338): window written back, window recentred on the agent, local map and pose re-cut. */
int pn_map_update_full(pn_ctx* ctx, const pn_map_cfg* cfg, const pn_map_arrays* arrays, int E, void* stream);

/* Head of Agent_State.update_prediction (:350-360), for E environments, without the reference's device->host->device round
 * trip of the window.  pn_map_stamp_local: full_map[e, :, lmb0:lmb1, lmb2:lmb3] = local_map[e] (:350-351; lmb [E,4] device).
 * pn_map_crop_window: window_out[e, c] = full_map[e, c, x1:x1+win_w, y1:y1+win_h] for c < copy_channels (:357-360, or the whole
 * map when the window is the map, :353-354); window_out is [E, out_channels, win_w, win_h] and its planes >= copy_channels are
 * left untouched (a caller whose completion net takes more input planes than the map has keeps its own there). */
int pn_map_stamp_local(pn_ctx* ctx, const float* local_map_dev, float* full_map_dev, const int* lmb_dev, int E, int num_channels,
                       int local_w, int local_h, int full_w, int full_h, void* stream);
int pn_map_crop_window(pn_ctx* ctx, const float* full_map_dev, int E, int num_channels, int full_w, int full_h, int x1, int y1,
                       int win_w, int win_h, int copy_channels, float* window_out_dev, int out_channels, void* stream);

/* N4, the map-sequence file format either side of the path, on the device.
 * pn_map_quantize: out[i] = (uint8)(map[i] * 255), fp32 multiply and truncation - `(full_map.cpu().numpy() * 255).astype(np.uint8)`
 *   of nav/collect_maps.py:79-80 without moving the fp32 map to the host (defined, like numpy's cast, for products in [0, 256)).
 * pn_map_sample: LoadMapFromFile.__call__ (prediction/train_prediction_model.py:66-84) on a sequence seq [T, C, W, H] uint8 that
 *   is resident in HBM: img = seq[t_idx] / 255 as float32, written as img_hwc [W, H, C] (the reference's array) and / or img_chw
 *   [C, W, H] (the layout pn_prednet_forward takes); gt [W, H, num_goals] int64 = seq[T-1, goal_channel0 + g] where the INPUT's
 *   explored channel (channel 1) is still 0, else 0.  Any of the three outputs may be NULL; t_idx may be negative (from the end). */
int pn_map_quantize(pn_ctx* ctx, const float* map_dev, long long count, unsigned char* out_dev, void* stream);
int pn_map_sample(pn_ctx* ctx, const unsigned char* seq_dev, int T, int C, int W, int H, int t_idx, int goal_channel0, int num_goals,
                  float* img_hwc_dev, float* img_chw_dev, long long* gt_dev, void* stream);

/* Agent_State.update_goal_map (:423-452), every step, for E environments: goal_map_out [E, local_w, local_h] fp32 = the
 * cells of category channel goal_cat + 4 (binarised; eroded goal_erode times and dilated once with the 4-neighbour cross
 * unless skip_morph[e] != 0, the reference's "'tv' in goal_name") that carry no OTHER category of channels 4..9;
 * found_goal_out [E] = 1 if any such cell exists, else 0 and goal_map_out = a single 1 at global_goal[e].  All pointers are
 * device pointers.  Assumes map values >= 0 (the mapper clamps to [0, 1]), which makes the reference's two `.sum() != 0`
 * tests equivalent to "any cell set". */
int pn_goal_map(pn_ctx* ctx, const float* local_map_dev, int E, int num_channels, int local_w, int local_h,
                const int* goal_cat_dev, const int* skip_morph_dev, const int* global_goal_dev, int goal_erode,
                float* goal_map_out_dev, int* found_goal_out_dev, void* stream);

/* ---- Agent_State.update_global_goal (nav/agent/agent_state.py:376-416), for E environments with
 * every field in device memory: obstacle dilation by disk(col_rad) -> traversible mask (collision cells blocked, visited
 * cells free, the agent's cell always free), geodesic distance from the agent's cell (the eikonal discretisation of
 * scikit-fmm's second-order distance marcher, solved by a block-parallel fixed-point iteration around an exact sequential
 * replay of the marcher's first cells), exp(-dd / (dist_weight_temperature / map_resolution)) weighting inside the
 * local-map window with the "stuck inside an obstacle: keep the last weights" rule, value = target_pred * weight (or the two
 * special temperatures -1 / 0), first argmax, and the "avoid repeating the last goal" bookkeeping. */
typedef struct pn_goal_cfg {
  int num_channels;               /* channels of full_map (channel 0 = obstacles)                 */
  int full_w, full_h;             /* 960 x 960                                                     */
  int local_w, local_h;           /* 480 x 480                                                     */
  int col_rad;                    /* 4     arguments.py:86 (disk radius of the obstacle dilation)  */
  int map_resolution;             /* 5                                                             */
  double dist_weight_temperature; /* 500   arguments.py:100 (-1: no weighting, 0: frontier mode)   */
} pn_goal_cfg;

typedef struct pn_goal_arrays {
  const float* full_map;        /* [E, nc, full_w, full_h]                                                     */
  const uint8_t* collision_map; /* [E, full_w, full_h] 1 = collision (helper.collision_map == 1), or NULL      */
  const uint8_t* visited_vis;   /* [E, full_w, full_h] 1 = visited (helper.visited_vis == 1), or NULL          */
  const int* lmb;               /* [E, 4]                                                                      */
  const int* loc;               /* [E, 2] loc_r, loc_c (pn_map_update_local)                                   */
  const float* target_pred;     /* [E, local_w, local_h] (pn_target_pred); may be NULL when only_distance      */
  double* dd;                   /* [E, full_w, full_h] out: geodesic distance in cells, inf = blocked/unreached */
  double* dd_wt;                /* [E, local_w, local_h] in/out (self.dd_wt)                                   */
  int* dd_wt_valid;             /* [E] in/out: 0 = self.dd_wt is None                                          */
  double* value;                /* [E, local_w, local_h] out (self.value) or NULL                              */
  int* global_goal;             /* [E, 2] in/out (self.global_goals[0])                                        */
  int* goal_kind;               /* [E] in/out: 1 = list of lists (init / presets), 2 = tuple written here      */
  int* last_global_goal;        /* [E, 2] in/out (self.last_global_goal[0])                                    */
  int* last_kind;               /* [E] in/out: 0 = None, else the kind of the stored goal; a list never equals the tuple
                                   np.unravel_index yields, so only kind 2 can suppress a repeated goal (:413)  */
} pn_goal_arrays;

/* only_distance != 0 stops after dd (the geodesic field alone: what the planner's FMMPlanner.set_goal also computes). */
int pn_global_goal(pn_ctx* ctx, const pn_goal_cfg* cfg, const pn_goal_arrays* arrays, int E, int only_distance, void* stream);

/* ---- Glue between stages A and B: Agent_Helper._preprocess_obs / _preprocess_depth
 * (nav/agent/agent_helper.py:175-217).  depth [E,H,W] fp32 as the simulator emits it (0 = invalid, 1 = max range),
 * rgb [E,H,W,3] uint8 (may be NULL: channels 0-2 are unused by the mapper), sem [E,H,W,num_sem] fp32 (stage A output)
 * -> obs [E, 4+num_sem, frame_height, frame_width] (channel 3 = depth in cm, [ds/2::ds] subsampling). */
int pn_make_obs(pn_ctx* ctx, const float* depth_dev, const uint8_t* rgb_dev, const float* sem_dev, int E, int H, int W,
                int frame_height, int frame_width, int num_sem, float min_depth, float max_depth, float* obs_out_dev,
                void* stream);

/* ---- Stage B: Semantic_Mapping ("Sem_Map_Module"), batched over environments.
 * Replaces Semantic_Mapping.forward, nav/agent/mapping.py:52-179 (call sites nav/agent/agent_state.py:114-115,
 * 273-274).  The configuration mirrors the argparse flags the module reads (nav/arguments.py:44-51, 58-60, 74-84). */
typedef struct pn_semmap_cfg {
  int frame_height;        /* 120 */
  int frame_width;         /* 160 */
  int map_resolution;      /* 5 (cm per cell) */
  int map_size_cm;         /* 4800 */
  int global_downscaling;  /* 2  -> local map of 480 x 480 cells */
  int vision_range;        /* 100 */
  int du_scale;            /* 1 (only value supported) */
  int num_sem_categories;  /* 10 */
  double hfov;             /* 79; double: f = (w/2) / tan(deg2rad(hfov/2)) is evaluated in Python float (depth_utils.py:31, mapping.py:36) */
  double camera_height;    /* 0.88 (m); double: max_z = int((camera_height*100 + 1)/res - min_h) (mapping.py:33, 103) */
  float cat_pred_threshold; /* 5.0 */
  float exp_pred_threshold; /* 1.0 */
  float map_pred_threshold; /* 0.1 */
} pn_semmap_cfg;

int pn_semmap_build(pn_ctx* ctx, int num_envs, const pn_semmap_cfg* cfg);
/* obs [E, 4+S, h, w] (channel 3 = depth in cm, 4.. = semantic masks), pose_delta [E,3] (dx m, dy m, dtheta rad),
 * maps_last [E, 4+S, n, n] - dense when maps_last_strides is NULL, else a strided view with element strides
 * {env, channel, row} (the reference passes a window of full_map, agent_state.py:206-208) -, poses_inout [E,3] (x m, y m, theta deg; UPDATED IN PLACE like the reference mutates
 * poses_last), fp_map_out [E, vr, vr] (may be NULL), map_out [E, 4+S, n, n] (must not alias maps_last). */
int pn_semmap_forward(pn_ctx* ctx, const float* obs_dev, const float* pose_delta_dev, const float* maps_last_dev,
                      const int64_t* maps_last_strides, float* poses_inout_dev, float* fp_map_out_dev,
                      float* map_out_dev, void* stream);
/* Parity tap: ego map [E, 2+S, vr, vr] (obstacle, explored, categories) before resampling, and the per-env
 * stair-mask decision (mapping.py:94).  Either pointer may be NULL. */
int pn_semmap_read_ego(pn_ctx* ctx, float* ego_out_dev, int* stair_flags_out_dev, void* stream);
int pn_semmap_num_launches(pn_ctx* ctx);

/* ---- Result gather across the GPUs of one node (the path shards over independent environments,
 * the only exchange is each rank's results travelling to the rank that hosts the planner).  No reference counterpart: the
 * reference is one process, one environment (nav/collect.py:32-33).  One-sided push over NVLink peer memory: the root
 * (pn_gather_create with rank == root) owns a slab [2 slots][world][bytes_per_rank]; pn_gather_export gives its 64-byte CUDA
 * IPC handle, which the host program hands to every other rank (torch.distributed, MPI, a file ...) for pn_gather_connect.
 * pn_gather_step enqueues on `stream`: the rank's push of local_dev (bytes_per_rank bytes) into its slice + a sequence flag;
 * on the root additionally the wait for every rank's flag, so work enqueued on the root's stream after the step sees all
 * results at pn_gather_result (slices are slice_stride bytes apart; valid until the root's next-but-one step).  Every rank
 * must call pn_gather_step the same number of times.  Waits are bounded (4 s): pn_gather_status reports 0 = ok, 2 = a push
 * timed out waiting for the root, 3 = the root timed out waiting for a rank. */
typedef struct pn_gather pn_gather;
int pn_gather_create(pn_ctx* ctx, int rank, int world, int root, int64_t bytes_per_rank, pn_gather** out);
int pn_gather_export(pn_gather* g, void* handle64_out);
int pn_gather_connect(pn_gather* g, const void* handle64);
int pn_gather_step(pn_gather* g, const void* local_dev, void* stream);
int pn_gather_result(pn_gather* g, void** slab_dev_out, int64_t* slice_stride_out);
int pn_gather_status(pn_gather* g, int* status_out);
int pn_gather_destroy(pn_gather* g);

/* ---- Launch-configuration table of the tensor-core conv kernel.  At *_build time every conv layer picks its launch
 * configuration (N tile, split-K factor, CTA pair, CTAs per SM) from this process-wide table; a layer that is not in it
 * times a shortlist on its own buffers and adds the winner (PN_CONV_AUTOTUNE=table: never time, use the tile model's
 * choice instead; =0: ignore the table too).  Different configurations differ in fp32 summation order, so processes that
 * must produce bit-identical results (the ranks of one job) import the same table before building: text = lines
 * "key bn splits pair opt" as written by pn_conv_tuning_export ('#' starts a comment).  There is no reference counterpart
 * (cuDNN picks its algorithms behind mmcv / detectron2). */
int pn_conv_tuning_import(const char* text, int* entries_out);
/* Writes the table as text into buf (NUL-terminated); *needed_out = bytes required.  buf may be NULL to query the size. */
int pn_conv_tuning_export(char* buf, int64_t buf_bytes, int64_t* needed_out);
int pn_conv_tuning_clear(void);

/* ---- Single fused convolution (conv + per-channel scale/bias + optional residual + ReLU), used by the
 * parity tests of the tensor-core kernel against torch.nn.functional.conv2d.
 * x_dev [B,Cin,H,W], w_host [Cout,Cin,R,S], scale_host/bias_host [Cout] or NULL,
 * residual_dev [B,Cout,Ho,Wo] or NULL, y_dev [B,Cout,Ho,Wo]; force_bn: 0 = auto N tile, else 32/64/128/256;
 * OR-ing 0x1000 forces the direct-store epilogue instead of the TMA-staged one (both are parity-tested). */
int pn_conv2d(pn_ctx* ctx, int precision, const float* x_dev, int B, int Cin, int H, int W, const float* w_host,
              const float* scale_host, const float* bias_host, const float* residual_dev, int Cout, int R, int S,
              int stride, int dil, int pad, int relu, int force_bn, float* y_dev);

/* Tuning / roofline aid: builds one fused conv (+BN scale/bias, +ReLU, optional residual) on constant data and returns
 * the mean device time of `iters` back-to-back launches (CUDA events) and the N tile the heuristic (or force_bn) chose. */
int pn_conv_bench(pn_ctx* ctx, int precision, int B, int Cin, int H, int W, int Cout, int R, int S, int stride, int dil,
                  int pad, int with_residual, int force_bn, int iters, float* ms_out, int* bn_out);

#ifdef __cplusplus
}
#endif
#endif /* PEANUT_B200_H_ */
