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
  PN_TF32 = 1  /* fp32 storage, tf32 operands, fp32 accumulate (fp32-parity path) */
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

/* Per-launch device time of one forward pass (CUDA events around every recorded launch, eager mode,
 * averaged over `iters` passes after one warm-up).  ms_out[max_ops] receives milliseconds per op,
 * names_out a '\n'-separated list of op names.  Measurement aid for bench.py / profiles/. */
int pn_prednet_num_ops(pn_ctx* ctx);
int pn_prednet_profile(pn_ctx* ctx, int iters, float* ms_out, int max_ops, char* names_out, int names_bytes);

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
  float hfov;              /* 79 */
  float camera_height;     /* 0.88 (m) */
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

/* ---- Single fused convolution (conv + per-channel scale/bias + optional residual + ReLU), used by the
 * parity tests of the tensor-core kernel against torch.nn.functional.conv2d.
 * x_dev [B,Cin,H,W], w_host [Cout,Cin,R,S], scale_host/bias_host [Cout] or NULL,
 * residual_dev [B,Cout,Ho,Wo] or NULL, y_dev [B,Cout,Ho,Wo]; force_bn: 0 = auto N tile, else 32/64/128/256;
 * OR-ing 0x1000 forces the direct-store epilogue instead of the TMA-staged one (both are parity-tested). */
int pn_conv2d(pn_ctx* ctx, int precision, const float* x_dev, int B, int Cin, int H, int W, const float* w_host,
              const float* scale_host, const float* bias_host, const float* residual_dev, int Cout, int R, int S,
              int stride, int dil, int pad, int relu, int force_bn, float* y_dev);

#ifdef __cplusplus
}
#endif
#endif /* PEANUT_B200_H_ */
