// Stage C (map-completion encoder-decoder) network instance.
#pragma once
#include "engine.h"

namespace pn {

// Per-call pointers live in device memory so that the recorded launches (and the CUDA graph made of
// them) stay valid when the caller passes different buffers.
struct PredSlots {
  const float* input;
  float* output;
  int apply_sigmoid;
};

struct PredNet {
  Net net;
  int B = 0, C = 0, H = 0, W = 0, num_classes = 0;
  PredSlots* slots = nullptr;  // device
  Tensor features;             // layer4 output view (NHWC, inside the PSP concat buffer)
  Tensor logits_lowres;        // fp32 NHWC class logits at 1/8 resolution
  float* stage_in = nullptr;   // device staging used by the host-pointer entry point
  float* stage_out = nullptr;
};

void build_prednet(PredNet& net, const WeightStore& w, int B, int C, int H, int W, int num_classes, DType dt);

}  // namespace pn
