// Stage A (Mask-RCNN R101-FPN -> per-category mask stack) network instance and the non-GEMM kernels it uses.
#pragma once
#include "../../include/peanut_b200.h"
#include "engine.h"

namespace pn {

// Architecture / post-processing constants of nav/agent/utils/COCO-InstSeg/mask_rcnn_R_101_cat9.yaml
struct MrcnnCfg {
  int B = 1;                 // frames per forward
  int H = 480, W = 640;      // camera frame (configs/challenge_objectnav2022.local.rgbd.yaml:15-31)
  int min_size = 800;        // INPUT.MIN_SIZE_TEST (yaml:30)
  int max_size = 1333;       // INPUT.MAX_SIZE_TEST (yaml:28)
  int pre_nms_topk = 1000;   // RPN.PRE_NMS_TOPK_TEST (yaml:251)
  int post_nms_topk = 1000;  // RPN.POST_NMS_TOPK_TEST (yaml:249)
  float rpn_nms = 0.7f;      // RPN.NMS_THRESH (yaml:247)
  int num_classes = 9;       // ROI_HEADS.NUM_CLASSES (yaml:193)
  float box_nms = 0.5f;      // ROI_HEADS.NMS_THRESH_TEST (yaml:192)
  int detections = 100;      // TEST.DETECTIONS_PER_IMAGE (yaml:312)
  float mask_thresh = 0.5f;  // paste_masks_in_image threshold
};

// Per-call values read by the recorded launches from device memory (CUDA-graph stable).
struct MrcnnSlots {
  const uint8_t* rgb;     // [B,H,W,3] uint8 RGB
  float* sem_out;         // [B,H,W,num_classes+1] fp32
  const int* goal_cat;    // [B] or null
  float score_thresh;     // ROI_HEADS.SCORE_THRESH_TEST  (segmentation.py:33)
  float sem_thr;          // args.sem_pred_prob_thr gate  (segmentation.py:54)
  float goal_thr;         // args.goal_thr gate           (segmentation.py:56-58)
};

constexpr int kRpnLevels = 5;
constexpr int kRpnCap = 1024;    // per-level candidate capacity (>= pre_nms_topk)
constexpr int kAnchors = 3;
constexpr int kRpnHeadC = 16;    // 3 objectness + 12 deltas, padded

struct RpnLevel {
  const float* head;  // fp32 NHWC [B, H, W, kRpnHeadC]
  int H, W, stride;
  float base[kAnchors][4];  // cell anchors (x1, y1, x2, y2)
};
struct RpnMeta {
  RpnLevel lv[kRpnLevels];
  int pre_topk, post_topk;
  float nms_thr;
  float img_h, img_w;  // resized (unpadded) image size the boxes are clipped to
};

struct PyramidLevel {
  const void* ptr;
  int H, W;
  long long ld;
  float scale;  // 1 / stride
};
struct Pyramid {
  PyramidLevel lv[4];  // p2..p5
  int keep_fp32;       // fp32 parity mode (Net::x3): pooled values are stored unrounded
};

struct MaskRcnn {
  Net net;
  MrcnnCfg cfg;
  int Hn = 0, Wn = 0, Hp = 0, Wp = 0;  // resized and padded network input size
  MrcnnSlots* slots = nullptr;          // device
  // device buffers (capacities fixed; counts live on the device)
  uint8_t* resized_u8 = nullptr;        // [B,Hn,Wn,3] BGR, parity tap of the PIL-exact resize
  float* rpn_head[kRpnLevels] = {};     // fp32 NHWC [B,H,W,16] per level
  int rpn_hw[kRpnLevels][2] = {};
  float* lvl_boxes = nullptr;           // [B][5][kRpnCap][4] kept per level, score order
  float* lvl_scores = nullptr;          // [B][5][kRpnCap]
  int* lvl_count = nullptr;             // [B][5]
  float* prop_boxes = nullptr;          // [B*post_topk][4]
  float* prop_scores = nullptr;         // [B*post_topk]
  int* prop_img = nullptr;              // [B*post_topk] image index or -1
  int* prop_count = nullptr;            // [B]
  float* box_out = nullptr;             // fp32 [B*post_topk][64]: 0..K class logits, K+1.. 4K deltas
  float* det_boxes = nullptr;           // [B][detections][4] (network-input coordinates)
  float* det_scores = nullptr;          // [B][detections]
  int* det_classes = nullptr;           // [B][detections]
  int* det_count = nullptr;             // [B]
  float* mroi_boxes = nullptr;          // compacted over the batch: [B*detections][4]
  int* mroi_img = nullptr;              // [B*detections] image index or -1
  int* mroi_cls = nullptr;              // [B*detections]
  int* mroi_total = nullptr;            // [1] + det_start [B] behind it
  float* mask_logits = nullptr;         // fp32 [B*detections*14*14*4][16]
  Tensor input;                         // stem input [B, Hp, Wp/2, 32]: 7 horizontal taps x BGR packed into channels
  Pyramid pyramid{};
  float* stage_sem = nullptr;           // staging for the host entry point
  uint8_t* stage_rgb = nullptr;
};

void build_maskrcnn(MaskRcnn& m, const WeightStore& w, const MrcnnCfg& cfg, DType dt);
void resized_shape(int h, int w, int min_size, int max_size, int& newh, int& neww);

// preproc.cu
void pil_bilinear_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk, int& ksize);
void add_resize_pack_stem(Net& net, const uint8_t* const* rgb_slot, int B, int H, int W, int Hn, int Wn, const Tensor& out,
                          uint8_t* resized_u8, const float mean_bgr[3], const float std_bgr[3]);
void launch_target_pred(const float* pred, int win, int x1, int y1, int goal_cat, int r0, int c0, int lw, int lh,
                        const float* explored, long long explored_row_stride, float* out, cudaStream_t s);
void launch_make_obs(const float* depth, const uint8_t* rgb, const float* sem, int E, int H, int W, int ds, int h, int w,
                     int nsem, float min_d, float max_d, float* obs, cudaStream_t s);

// mapstate.cu: op 0 init_map_and_pose, 1 init_with_obs stamp, 2 update_local_map tail, 3 update_full_map
void launch_map_stamp_local(const float* local_map, float* full_map, const int* lmb, int E, int nc, int local_w, int local_h,
                            int full_w, int full_h, cudaStream_t s);
void launch_map_crop(const float* full_map, int E, int nc, int full_w, int full_h, int x1, int y1, int win_w, int win_h,
                     int nc_copy, float* out, int out_channels, cudaStream_t s);
// N4 (mapstate.cu): writer-side quantisation and reader-side sample construction of the map-sequence format
void launch_map_quantize(const float* map, long long n, uint8_t* out, int num_sms, cudaStream_t s);
void launch_map_sample(const uint8_t* seq, int T, int C, int W, int H, int t_idx, int goal0, int G, float* img_hwc, float* img_chw,
                       long long* gt, cudaStream_t s);
void launch_global_goal(int device, int num_sms, const pn_goal_cfg& cfg, const pn_goal_arrays& arrays, int E, int only_distance,
                        cudaStream_t s);
void map_bookkeeping(int op, const pn_map_cfg& cfg, const pn_map_arrays& arrays, int E, cudaStream_t s);
void launch_goal_map(const float* local_map, int E, int nc, int w, int h, const int* goal_cat, const int* skip_morph,
                     const int* global_goal, int goal_erode, float* goal_map, int* found, cudaStream_t s);

// detect.cu
void add_upsample2x_add(Net& net, const Tensor& prev, const Tensor& lat);   // lat += nearest_up2(prev)
void add_subsample2(Net& net, const Tensor& in, const Tensor& out);         // max_pool2d(k=1, s=2)
void add_rpn_proposals(Net& net, MaskRcnn& m, const RpnMeta& meta);
void add_roi_align(Net& net, const std::string& name, const Pyramid& pyr, DType dt, const float* boxes, const int* img,
                   int nrois, int S, const Tensor& out);
void add_detections(Net& net, MaskRcnn& m);
void add_paste_accumulate(Net& net, MaskRcnn& m);

}  // namespace pn
