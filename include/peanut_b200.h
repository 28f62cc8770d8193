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
