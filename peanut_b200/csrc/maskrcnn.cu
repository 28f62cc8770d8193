// Stage A network: Mask-RCNN R101-FPN as configured by nav/agent/utils/COCO-InstSeg/mask_rcnn_R_101_cat9.yaml,
// recorded as a flat list of tcgen05 convolution launches (backbone, FPN, RPN head, box head FCs, mask head) and
// the control-flow kernels of detect.cu.  Everything runs at fixed capacity with device-side counts (1000
// proposals and 100 detections per frame), so one forward is a constant launch sequence that is replayed as a
// CUDA graph with no host synchronisation (the reference syncs per instance: segmentation.py:48-60).
//
// Weight names are detectron2's checkpoint keys.  detectron2 semantics restated in
// oracle/maskrcnn.py; the yaml lines that fix each constant are cited there and in maskrcnn.h.
#include <cmath>

#include "maskrcnn.h"

namespace pn {

void resized_shape(int h, int w, int min_size, int max_size, int& newh, int& neww) {
  // ResizeShortestEdge.get_output_shape (python float = double arithmetic)
  double scale = min_size * 1.0 / std::min(h, w);
  double nh, nw;
  if (h < w) nh = min_size, nw = scale * w;
  else nh = scale * h, nw = min_size;
  if (std::max(nh, nw) > max_size) {
    scale = max_size * 1.0 / std::max(nh, nw);
    nh = nh * scale, nw = nw * scale;
  }
  newh = static_cast<int>(nh + 0.5), neww = static_cast<int>(nw + 0.5);
}

namespace {

const HostArray& conv_weight(const WeightStore& w, const std::string& name, int cout, int cin, int k) {
  const HostArray& wt = get_weight(w, name);
  const bool ok4 = wt.shape.size() == 4 && wt.shape[0] == cout && wt.shape[1] == cin && wt.shape[2] == k && wt.shape[3] == k;
  const bool ok2 = k == 1 && wt.shape.size() == 2 && wt.shape[0] == cout && wt.shape[1] == cin;
  PN_REQUIRE(ok4 || ok2, "weight shape mismatch at " + name);
  return wt;
}

ConvSpec spec(int cin, int cout, int k, int stride, int pad, bool relu) {
  ConvSpec s;
  s.Cin = cin, s.Cout = cout, s.R = k, s.S = k, s.stride = stride, s.dil = 1, s.pad = pad, s.relu = relu;
  return s;
}

// conv (no bias) + FrozenBN folded (+ residual) (+ ReLU)
Tensor conv_frozen_bn(Net& net, const WeightStore& w, const std::string& name, const Tensor& in, ConvSpec sp,
                      const Tensor* residual = nullptr) {
  const HostArray& wt = conv_weight(w, name + ".weight", sp.Cout, sp.Cin, sp.R);
  std::vector<float> scale, bias;
  fold_bn(w, name + ".norm", sp.Cout, scale, bias, 1e-5f);
  Tensor out = net.arena.tensor(in.B, conv_out(in.H, sp.R, sp.stride, 1, sp.pad), conv_out(in.W, sp.S, sp.stride, 1, sp.pad),
                                pad_channels(sp.Cout, in.dt), in.dt);
  add_conv(net, name, in, out, wt.data.data(), scale.data(), bias.data(), sp, residual);
  return out;
}

// conv with bias, no norm
Tensor conv_bias(Net& net, const WeightStore& w, const std::string& name, const Tensor& in, ConvSpec sp,
                 const Tensor* dst = nullptr) {
  const HostArray& wt = conv_weight(w, name + ".weight", sp.Cout, sp.Cin, sp.R);
  const HostArray& bs = get_weight(w, name + ".bias");
  PN_REQUIRE(bs.numel() == sp.Cout, "bias size mismatch at " + name);
  Tensor out = dst ? *dst
                   : net.arena.tensor(in.B, conv_out(in.H, sp.R, sp.stride, 1, sp.pad), conv_out(in.W, sp.S, sp.stride, 1, sp.pad),
                                      pad_channels(sp.Cout, in.dt), in.dt);
  add_conv(net, name, in, out, wt.data.data(), nullptr, bs.data.data(), sp);
  return out;
}

}  // namespace

void build_maskrcnn(MaskRcnn& m, const WeightStore& w, const MrcnnCfg& cfg, DType dt) {
  Net& net = m.net;
  net.dt = dt;
  m.cfg = cfg;
  const int B = cfg.B, K = cfg.num_classes;
  PN_REQUIRE(B >= 1 && cfg.H >= 32 && cfg.W >= 32, "maskrcnn: bad frame shape");
  PN_REQUIRE(cfg.post_nms_topk <= 4096 && cfg.pre_nms_topk <= kRpnCap, "maskrcnn: top-k out of range");
  resized_shape(cfg.H, cfg.W, cfg.min_size, cfg.max_size, m.Hn, m.Wn);
  m.Hp = round_up(m.Hn, 32), m.Wp = round_up(m.Wn, 32);
  m.slots = static_cast<MrcnnSlots*>(net.arena.alloc(sizeof(MrcnnSlots)));
  Arena& A = net.arena;
  const int R = cfg.post_nms_topk, D = cfg.detections;

  // ---- input: PIL-exact resize, BGR, mean subtraction, zero padding (yaml:26-30, 82-89)
  net.stage("preprocess");
  m.resized_u8 = static_cast<uint8_t*>(A.alloc(static_cast<size_t>(B) * m.Hn * m.Wn * 3));
  PN_REQUIRE(m.Wp % 4 == 0, "maskrcnn: padded width must be a multiple of 4");
  PN_REQUIRE(m.Hp % 2 == 0, "maskrcnn: padded height must be even");
  // one packed vector per PAIR of stem output pixels and per PAIR of image rows (rows 2q - 1, 2q; see k_pack_stem)
  m.input = A.tensor(B, m.Hp / 2 + 1, m.Wp / 4, 64, dt);
  const float mean_bgr[3] = {103.53f, 116.28f, 123.675f}, std_bgr[3] = {1.f, 1.f, 1.f};
  add_resize_pack_stem(net, &m.slots->rgb, B, cfg.H, cfg.W, m.Hn, m.Wn, m.input, m.resized_u8, mean_bgr, std_bgr);
  net.taps["stem_in"] = m.input;

  // ---- ResNet-101 bottom-up, STRIDE_IN_1X1 (yaml:101-112)
  net.stage("backbone");
  const std::string bu = "backbone.bottom_up.";
  Tensor x;
  {
    // 7x7 stride-2 stem as a 4x1 convolution over the tap-packed input (see k_pack_stem), two output pixels per GEMM row and
    // two image rows per input pixel:
    //   W'[p*64 + co][r][h*32 + t*3 + c] = W[co][c][2r + h][t - 2p]  for 2r + h < 7 and 0 <= t - 2p < 7
    // (p = parity of the output pixel, t = 0..8 packed column, r = 0..3 row-pair tap, h = row inside the pair)
    const HostArray& wt = conv_weight(w, bu + "stem.conv1.weight", 64, 3, 7);
    std::vector<float> wp(static_cast<size_t>(128) * 64 * 4, 0.f);
    for (int par = 0; par < 2; ++par)
      for (int co = 0; co < 64; ++co)
        for (int c = 0; c < 3; ++c)
          for (int rr = 0; rr < 7; ++rr)
            for (int s = 0; s < 7; ++s)
              wp[(static_cast<size_t>(par * 64 + co) * 64 + ((rr & 1) * 32 + (s + 2 * par) * 3 + c)) * 4 + (rr >> 1)] =
                  wt.data[((static_cast<size_t>(co) * 3 + c) * 7 + rr) * 7 + s];
    std::vector<float> scale, bias;
    fold_bn(w, bu + "stem.conv1.norm", 64, scale, bias, 1e-5f);
    scale.insert(scale.end(), scale.begin(), scale.begin() + 64);   // both pixel parities
    bias.insert(bias.end(), bias.begin(), bias.begin() + 64);
    ConvSpec s7;
    s7.Cin = 64, s7.Cout = 128, s7.R = 4, s7.S = 1, s7.stride = 1, s7.stride_w = 1, s7.dil = 1, s7.pad = 1, s7.pad_w = 0, s7.relu = true;
    x = A.tensor(B, conv_out(m.Hp, 7, 2, 1, 3), m.Wp / 2, 64, dt);
    Tensor xpair = x;   // the same memory seen as [B, Ho, Wo / 2, 128]: NHWC of a pixel pair = (parity, channel)
    xpair.W = x.W / 2, xpair.C = 128, xpair.ld = 2 * x.ld;
    PN_REQUIRE(x.ld == 64, "maskrcnn: the stem output must be dense");
    add_conv(net, bu + "stem.conv1", m.input, xpair, wp.data(), scale.data(), bias.data(), s7);
    // packed channels 27..31, two of the nine taps per pixel and the eighth row are padding: report the 7x7x3 FLOPs, not the packed K
    net.op_flops.back() = 2.0 * static_cast<double>(x.pixels()) * 64 * 147;
  }
  {
    Tensor p = A.tensor(B, conv_out(x.H, 3, 2, 1, 1), conv_out(x.W, 3, 2, 1, 1), x.C, dt);
    add_maxpool3x3s2(net, x, p);
    x = p;
  }
  const int blocks[4] = {3, 4, 23, 3};
  Tensor res[4];
  int cin = 64;
  for (int si = 0; si < 4; ++si) {
    const int mid = 64 << si, cout = 256 << si;
    for (int bi = 0; bi < blocks[si]; ++bi) {
      const std::string pre = bu + "res" + std::to_string(si + 2) + "." + std::to_string(bi);
      const int stride = (bi == 0 && si > 0) ? 2 : 1;
      Tensor identity = x;
      if (bi == 0) {  // the projection shortcut runs beside conv1 -> conv2 on a side lane
        net.set_lane(1);
        identity = conv_frozen_bn(net, w, pre + ".shortcut", x, spec(cin, cout, 1, stride, 0, false));
        net.set_lane(0);
      }
      Tensor t = conv_frozen_bn(net, w, pre + ".conv1", x, spec(cin, mid, 1, stride, 0, true));
      t = conv_frozen_bn(net, w, pre + ".conv2", t, spec(mid, mid, 3, 1, 1, true));
      if (bi == 0) net.join_lanes();
      x = conv_frozen_bn(net, w, pre + ".conv3", t, spec(mid, cout, 1, 1, 0, true), &identity);
      cin = cout;
    }
    res[si] = x;
    net.taps["res" + std::to_string(si + 2)] = x;
  }

  // ---- FPN (yaml:62-70): lateral 1x1 (+ nearest x2 of the coarser level), output 3x3, p6 = stride-2 subsample of p5
  // ---- RPN head (shared over levels): 3x3 + ReLU, then objectness (3) and anchor deltas (12) as ONE 1x1 -> fp32.
  // The top-down chain (laterals + upsample-adds) and the finest level stay on lane 0; the output conv and the RPN head
  // of each coarser level form a chain of their own on a side lane (P5 and P6 share one) and run beside it.
  net.stage("fpn");
  Tensor P[5];
  const int strides[5] = {4, 8, 16, 32, 64};
  RpnMeta meta{};
  const std::string rp = "proposal_generator.rpn_head.";
  std::vector<float> wm(15 * 256), bm(15);
  {
    const HostArray& wo = conv_weight(w, rp + "objectness_logits.weight", kAnchors, 256, 1);
    const HostArray& bo = get_weight(w, rp + "objectness_logits.bias");
    const HostArray& wd = conv_weight(w, rp + "anchor_deltas.weight", 4 * kAnchors, 256, 1);
    const HostArray& bd = get_weight(w, rp + "anchor_deltas.bias");
    std::copy(wo.data.begin(), wo.data.end(), wm.begin());
    std::copy(wd.data.begin(), wd.data.end(), wm.begin() + 3 * 256);
    std::copy(bo.data.begin(), bo.data.end(), bm.begin());
    std::copy(bd.data.begin(), bd.data.end(), bm.begin() + 3);
  }
  auto add_rpn_level = [&](int l) {
    const double sizes[5] = {32, 64, 128, 256, 512}, ratios[3] = {0.5, 1.0, 2.0};  // yaml:45-58
    Tensor t = conv_bias(net, w, rp + "conv", P[l], spec(256, 256, 3, 1, 1, true));
    Tensor head = A.tensor(B, t.H, t.W, kRpnHeadC, kF32);
    ConvSpec s = spec(256, 15, 1, 1, 0, false);
    s.out_fp32 = true;
    add_conv(net, rp + "predictors.p" + std::to_string(l + 2), t, head, wm.data(), nullptr, bm.data(), s);
    m.rpn_head[l] = static_cast<float*>(head.ptr);
    m.rpn_hw[l][0] = t.H, m.rpn_hw[l][1] = t.W;
    RpnLevel& lv = meta.lv[l];
    lv.head = m.rpn_head[l], lv.H = t.H, lv.W = t.W, lv.stride = strides[l];
    for (int a = 0; a < kAnchors; ++a) {  // DefaultAnchorGenerator.generate_cell_anchors
      const double area = sizes[l] * sizes[l];
      const double ww = std::sqrt(area / ratios[a]), hh = ratios[a] * ww;
      lv.base[a][0] = static_cast<float>(-ww / 2.0), lv.base[a][1] = static_cast<float>(-hh / 2.0);
      lv.base[a][2] = static_cast<float>(ww / 2.0), lv.base[a][3] = static_cast<float>(hh / 2.0);
    }
    net.raw_taps["rpn_head.p" + std::to_string(l + 2)] = {head.ptr, head.bytes()};
  };
  Tensor prev;
  for (int lvl = 5; lvl >= 2; --lvl) {
    const Tensor& c = res[lvl - 2];
    Tensor lat = conv_bias(net, w, "backbone.fpn_lateral" + std::to_string(lvl), c, spec(c.C, 256, 1, 1, 0, false));
    if (lvl < 5) add_upsample2x_add(net, prev, lat);
    prev = lat;
    net.set_lane(lvl == 2 ? 0 : 6 - lvl);  // P5 (+P6) -> lane 1, P4 -> lane 2, P3 -> lane 3, P2 -> lane 0
    P[lvl - 2] = conv_bias(net, w, "backbone.fpn_output" + std::to_string(lvl), lat, spec(256, 256, 3, 1, 1, false));
    net.taps["p" + std::to_string(lvl)] = P[lvl - 2];
    add_rpn_level(lvl - 2);
    if (lvl == 5) {
      P[4] = A.tensor(B, (P[3].H - 1) / 2 + 1, (P[3].W - 1) / 2 + 1, 256, dt);
      add_subsample2(net, P[3], P[4]);
      net.taps["p6"] = P[4];
      add_rpn_level(4);
    }
    net.set_lane(0);
  }
  for (int l = 0; l < 4; ++l) m.pyramid.lv[l] = PyramidLevel{P[l].ptr, P[l].H, P[l].W, P[l].ld, 1.0f / strides[l]};
  net.stage("rpn_head");  // (kept as a stage name; its launches are interleaved with the FPN above)
  net.join_lanes();
  meta.pre_topk = cfg.pre_nms_topk, meta.post_topk = R, meta.nms_thr = cfg.rpn_nms;
  meta.img_h = static_cast<float>(m.Hn), meta.img_w = static_cast<float>(m.Wn);

  // ---- proposals (yaml:224-256)
  net.stage("rpn_proposals");
  m.lvl_boxes = static_cast<float*>(A.alloc(static_cast<size_t>(B) * kRpnLevels * kRpnCap * 4 * sizeof(float)));
  m.lvl_scores = static_cast<float*>(A.alloc(static_cast<size_t>(B) * kRpnLevels * kRpnCap * sizeof(float)));
  m.lvl_count = static_cast<int*>(A.alloc(static_cast<size_t>(B) * kRpnLevels * sizeof(int)));
  m.prop_boxes = static_cast<float*>(A.alloc(static_cast<size_t>(B) * R * 4 * sizeof(float)));
  m.prop_scores = static_cast<float*>(A.alloc(static_cast<size_t>(B) * R * sizeof(float)));
  m.prop_img = static_cast<int*>(A.alloc(static_cast<size_t>(B) * R * sizeof(int)));
  m.prop_count = static_cast<int*>(A.alloc(static_cast<size_t>(B) * sizeof(int)));
  add_rpn_proposals(net, m, meta);
  net.raw_taps["prop_boxes"] = {m.prop_boxes, static_cast<size_t>(B) * R * 4 * sizeof(float)};
  net.raw_taps["prop_scores"] = {m.prop_scores, static_cast<size_t>(B) * R * sizeof(float)};
  net.raw_taps["prop_img"] = {m.prop_img, static_cast<size_t>(B) * R * sizeof(int)};
  net.raw_taps["prop_count"] = {m.prop_count, static_cast<size_t>(B) * sizeof(int)};

  // ---- box head (yaml:165-196): ROIAlign 7x7 -> FC 1024 -> FC 1024 -> {cls (K+1), bbox (4K)} as one GEMM -> fp32
  net.stage("box_head");
  Tensor box_in = A.tensor(B * R, 7, 7, 256, dt);
  add_roi_align(net, "box_roi_align", m.pyramid, dt, m.prop_boxes, m.prop_img, B * R, 7, box_in);
  net.taps["box_pooled"] = box_in;
  {
    const std::string rh = "roi_heads.";
    // fc1 consumes the ROI features flattened as (c, y, x); ours are (y, x, c): permute the weight columns once
    const HostArray& w1 = conv_weight(w, rh + "box_head.fc1.weight", 1024, 256 * 49, 1);
    std::vector<float> w1p(w1.data.size());
    for (int o = 0; o < 1024; ++o)
      for (int c = 0; c < 256; ++c)
        for (int p = 0; p < 49; ++p) w1p[(static_cast<size_t>(o) * 49 + p) * 256 + c] = w1.data[(static_cast<size_t>(o) * 256 + c) * 49 + p];
    Tensor flat = box_in;
    flat.H = 1, flat.W = 1, flat.C = 49 * 256, flat.ld = 49 * 256;
    Tensor f1 = A.tensor(B * R, 1, 1, 1024, dt);
    add_conv(net, rh + "box_head.fc1", flat, f1, w1p.data(), nullptr, get_weight(w, rh + "box_head.fc1.bias").data.data(),
             spec(49 * 256, 1024, 1, 1, 0, true));
    Tensor f2 = A.tensor(B * R, 1, 1, 1024, dt);
    add_conv(net, rh + "box_head.fc2", f1, f2, conv_weight(w, rh + "box_head.fc2.weight", 1024, 1024, 1).data.data(), nullptr,
             get_weight(w, rh + "box_head.fc2.bias").data.data(), spec(1024, 1024, 1, 1, 0, true));
    const HostArray& wc = conv_weight(w, rh + "box_predictor.cls_score.weight", K + 1, 1024, 1);
    const HostArray& wb = conv_weight(w, rh + "box_predictor.bbox_pred.weight", 4 * K, 1024, 1);
    const HostArray& bc = get_weight(w, rh + "box_predictor.cls_score.bias");
    const HostArray& bb = get_weight(w, rh + "box_predictor.bbox_pred.bias");
    const int nout = 5 * K + 1;
    PN_REQUIRE(nout <= 64, "maskrcnn: too many classes for the merged predictor");
    std::vector<float> wm(static_cast<size_t>(nout) * 1024), bm(nout);
    std::copy(wc.data.begin(), wc.data.end(), wm.begin());
    std::copy(wb.data.begin(), wb.data.end(), wm.begin() + static_cast<size_t>(K + 1) * 1024);
    std::copy(bc.data.begin(), bc.data.end(), bm.begin());
    std::copy(bb.data.begin(), bb.data.end(), bm.begin() + K + 1);
    Tensor out = A.tensor(B * R, 1, 1, 64, kF32);
    ConvSpec s = spec(1024, nout, 1, 1, 0, false);
    s.out_fp32 = true;
    add_conv(net, rh + "box_predictor", f2, out, wm.data(), nullptr, bm.data(), s);
    m.box_out = static_cast<float*>(out.ptr);
    net.raw_taps["box_out"] = {out.ptr, out.bytes()};
  }

  // ---- detections (yaml:192, 312; SCORE_THRESH_TEST <- args.sem_pred_prob_thr, segmentation.py:33)
  net.stage("detections");
  m.det_boxes = static_cast<float*>(A.alloc(static_cast<size_t>(B) * D * 4 * sizeof(float)));
  m.det_scores = static_cast<float*>(A.alloc(static_cast<size_t>(B) * D * sizeof(float)));
  m.det_classes = static_cast<int*>(A.alloc(static_cast<size_t>(B) * D * sizeof(int)));
  m.det_count = static_cast<int*>(A.alloc(static_cast<size_t>(B) * sizeof(int)));
  m.mroi_boxes = static_cast<float*>(A.alloc(static_cast<size_t>(B) * D * 4 * sizeof(float)));
  m.mroi_img = static_cast<int*>(A.alloc(static_cast<size_t>(B) * D * sizeof(int)));
  m.mroi_cls = static_cast<int*>(A.alloc(static_cast<size_t>(B) * D * sizeof(int)));
  m.mroi_total = static_cast<int*>(A.alloc(static_cast<size_t>(B + 2) * sizeof(int)));
  add_detections(net, m);
  net.raw_taps["det_boxes"] = {m.det_boxes, static_cast<size_t>(B) * D * 4 * sizeof(float)};
  net.raw_taps["det_scores"] = {m.det_scores, static_cast<size_t>(B) * D * sizeof(float)};
  net.raw_taps["det_classes"] = {m.det_classes, static_cast<size_t>(B) * D * sizeof(int)};
  net.raw_taps["det_count"] = {m.det_count, static_cast<size_t>(B) * sizeof(int)};

  // ---- mask head (yaml:215-223) on the batch-compacted detections; rows beyond the live count are skipped
  net.stage("mask_head");
  Tensor mask_in = A.tensor(B * D, 14, 14, 256, dt);
  add_roi_align(net, "mask_roi_align", m.pyramid, dt, m.mroi_boxes, m.mroi_img, B * D, 14, mask_in);
  net.taps["mask_pooled"] = mask_in;
  {
    const std::string mh = "roi_heads.mask_head.";
    Tensor t = mask_in;
    for (int i = 1; i <= 4; ++i) {
      ConvSpec s = spec(256, 256, 3, 1, 1, true);
      s.m_limit = m.mroi_total, s.m_limit_rows = 14 * 14;
      const std::string name = mh + "mask_fcn" + std::to_string(i);
      Tensor o = A.tensor(B * D, 14, 14, 256, dt);
      add_conv(net, name, t, o, conv_weight(w, name + ".weight", 256, 256, 3).data.data(), nullptr,
               get_weight(w, name + ".bias").data.data(), s);
      t = o;
    }
    // ConvTranspose2d(256, 256, 2, stride 2) = four 1x1 convolutions; output channel (dy*2+dx)*256 + co
    const HostArray& wd = get_weight(w, mh + "deconv.weight");
    PN_REQUIRE(wd.shape.size() == 4 && wd.shape[0] == 256 && wd.shape[1] == 256 && wd.shape[2] == 2 && wd.shape[3] == 2, "deconv shape");
    const HostArray& bd = get_weight(w, mh + "deconv.bias");
    std::vector<float> wp(static_cast<size_t>(1024) * 256), bp(1024);
    for (int q = 0; q < 4; ++q)
      for (int co = 0; co < 256; ++co) {
        bp[q * 256 + co] = bd.data[co];
        for (int ci = 0; ci < 256; ++ci) wp[(static_cast<size_t>(q) * 256 + co) * 256 + ci] = wd.data[(static_cast<size_t>(ci) * 256 + co) * 4 + q];
      }
    Tensor up = A.tensor(B * D, 14, 14, 1024, dt);
    {
      ConvSpec s = spec(256, 1024, 1, 1, 0, true);
      s.m_limit = m.mroi_total, s.m_limit_rows = 14 * 14;
      add_conv(net, mh + "deconv", t, up, wp.data(), nullptr, bp.data(), s);
    }
    Tensor up_rows = up;  // the same memory as [B*D, 14, 14*4, 256]: one row per output pixel of the 28x28 mask
    up_rows.W = 14 * 4, up_rows.C = 256, up_rows.ld = 256;
    Tensor logits = A.tensor(B * D, 14, 14 * 4, 16, kF32);
    {
      ConvSpec s = spec(256, K, 1, 1, 0, false);
      s.out_fp32 = true;
      s.m_limit = m.mroi_total, s.m_limit_rows = 14 * 14 * 4;
      add_conv(net, mh + "predictor", up_rows, logits, conv_weight(w, mh + "predictor.weight", K, 256, 1).data.data(), nullptr,
               get_weight(w, mh + "predictor.bias").data.data(), s);
    }
    m.mask_logits = static_cast<float*>(logits.ptr);
    net.raw_taps["mask_logits"] = {logits.ptr, logits.bytes()};
  }

  // ---- paste + per-category accumulation (segmentation.py:47-62)
  net.stage("paste");
  add_paste_accumulate(net, m);
}

}  // namespace pn
