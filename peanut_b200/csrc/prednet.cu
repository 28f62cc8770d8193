// Map-completion network (stage C): ResNetV1c-50 (output stride 8, dilated) + PSPHead, built as a flat
// list of tcgen05 conv launches and HBM-bound helper kernels.
//
// Architecture follows the reference config nav/pred_model_cfg.py:2-42 and the modules it names:
//   ResNetV1c / ResNet      prediction/mmseg/models/backbones/resnet.py:396-527, 591-674, 688-700
//   Bottleneck (pytorch)    prediction/mmseg/models/backbones/resnet.py:99-307
//   ResLayer                prediction/mmseg/models/utils/res_layer.py:28-96
//   PSPHead / PPM           prediction/mmseg/models/decode_heads/psp_head.py:11-117
//   cls_seg                 prediction/mmseg/models/decode_heads/decode_head.py:225-230
//   final resize, no softmax prediction/mmseg/models/segmentors/encoder_decoder.py:70-80, 244-258
// Weight names are the mmcv checkpoint keys so a real pred_model_wts.pth loads unchanged.
#include "prednet.h"

namespace pn {

namespace {

struct BnConv {
  std::vector<float> scale, bias;
};

Tensor conv_bn(Net& net, const WeightStore& w, const std::string& conv_name, const std::string& bn_name,
               const Tensor& in, ConvSpec sp, const Tensor* residual = nullptr, const Tensor* dst = nullptr) {
  const HostArray& wt = get_weight(w, conv_name + ".weight");
  PN_REQUIRE(wt.shape.size() == 4 && wt.shape[0] == sp.Cout && wt.shape[1] == sp.Cin && wt.shape[2] == sp.R &&
                 wt.shape[3] == sp.S,
             "weight shape mismatch at " + conv_name);
  std::vector<float> scale, bias;
  fold_bn(w, bn_name, sp.Cout, scale, bias);
  const int Ho = conv_out(in.H, sp.R, sp.stride, sp.dil, sp.pad);
  const int Wo = conv_out(in.W, sp.S, sp.stride, sp.dil, sp.pad);
  Tensor out = dst ? *dst : net.arena.tensor(in.B, Ho, Wo, pad_channels(sp.Cout, in.dt), in.dt);
  add_conv(net, conv_name, in, out, wt.data.data(), scale.data(), bias.data(), sp, residual);
  return out;
}

}  // namespace

void build_prednet(PredNet& pn_, const WeightStore& w, int B, int C, int H, int W, int num_classes, DType dt) {
  Net& net = pn_.net;
  net.dt = dt;
  pn_.B = B, pn_.C = C, pn_.H = H, pn_.W = W, pn_.num_classes = num_classes;
  pn_.slots = static_cast<PredSlots*>(net.arena.alloc(sizeof(PredSlots)));

  // ---- input: fp32 NCHW -> NHWC (channels padded for the K blocking)
  Tensor x = net.arena.tensor(B, H, W, pad_channels(C, dt), dt);
  add_nchw_to_nhwc(net, &pn_.slots->input, x, C);

  // ---- deep stem (resnet.py:594-624) + maxpool (:638)
  auto spec = [](int cin, int cout, int k, int stride, int dil, bool relu) {
    ConvSpec s;
    s.Cin = cin, s.Cout = cout, s.R = k, s.S = k, s.stride = stride, s.dil = dil;
    s.pad = (k == 3) ? dil : 0;
    s.relu = relu;
    return s;
  };
  x = conv_bn(net, w, "backbone.stem.0", "backbone.stem.1", x, spec(C, 32, 3, 2, 1, true));
  x = conv_bn(net, w, "backbone.stem.3", "backbone.stem.4", x, spec(32, 32, 3, 1, 1, true));
  x = conv_bn(net, w, "backbone.stem.6", "backbone.stem.7", x, spec(32, 64, 3, 1, 1, true));
  {
    Tensor p = net.arena.tensor(B, conv_out(x.H, 3, 2, 1, 1), conv_out(x.W, 3, 2, 1, 1), x.C, dt);
    add_maxpool3x3s2(net, x, p);
    x = p;
  }

  // ---- residual stages (pred_model_cfg.py:10-15: strides (1,2,1,1), dilations (1,1,2,4), contract_dilation)
  const int blocks[4] = {3, 4, 6, 3};
  const int strides[4] = {1, 2, 1, 1};
  const int dilations[4] = {1, 1, 2, 4};
  int inplanes = 64;
  Tensor concat;  // PSP concat buffer; layer4's last block writes its first 2048 channels
  const int ppm_channels = 512;
  const std::vector<int> scales = {1, 2, 3, 6};
  for (int li = 0; li < 4; ++li) {
    const int planes = 64 << li;
    for (int bi = 0; bi < blocks[li]; ++bi) {
      const std::string pre = "backbone.layer" + std::to_string(li + 1) + "." + std::to_string(bi);
      const int stride = bi == 0 ? strides[li] : 1;
      int dil = dilations[li];
      if (bi == 0 && dil > 1) dil = dil / 2;  // contract_dilation (res_layer.py:69-72)
      Tensor identity = x;
      const bool has_down = bi == 0 && (stride != 1 || inplanes != planes * 4);
      if (has_down) {  // the projection shortcut runs beside conv1 -> conv2 on a side lane
        net.set_lane(1);
        identity = conv_bn(net, w, pre + ".downsample.0", pre + ".downsample.1", x,
                           spec(inplanes, planes * 4, 1, stride, 1, false));
        net.set_lane(0);
      }
      Tensor t = conv_bn(net, w, pre + ".conv1", pre + ".bn1", x, spec(inplanes, planes, 1, 1, 1, true));
      t = conv_bn(net, w, pre + ".conv2", pre + ".bn2", t, spec(planes, planes, 3, stride, dil, true));
      if (has_down) net.join_lanes();
      const bool last = (li == 3 && bi == blocks[li] - 1);
      Tensor dst;
      if (last) {
        concat = net.arena.tensor(B, t.H, t.W, planes * 4 + ppm_channels * static_cast<int>(scales.size()), dt);
        dst = concat.channels(0, planes * 4);
      }
      x = conv_bn(net, w, pre + ".conv3", pre + ".bn3", t, spec(planes, planes * 4, 1, 1, 1, true), &identity,
                  last ? &dst : nullptr);
      inplanes = planes * 4;
    }
  }

  // ---- PSP head (psp_head.py:48-59, 95-117)
  std::vector<Tensor> pooled;
  for (int s : scales) pooled.push_back(net.arena.tensor(B, s, s, inplanes, dt));
  add_ppm_pool(net, x, scales, pooled);
  for (size_t i = 0; i < scales.size(); ++i) {  // the four pyramid branches are independent: one lane each
    const std::string pre = "decode_head.psp_modules." + std::to_string(i) + ".1";
    net.set_lane(static_cast<int>(i) % Net::kMaxLanes);
    Tensor y = conv_bn(net, w, pre + ".conv", pre + ".bn", pooled[i], spec(inplanes, ppm_channels, 1, 1, 1, true));
    add_bilinear_into(net, y, concat.channels(inplanes + ppm_channels * static_cast<int>(i), ppm_channels));
  }
  net.join_lanes();
  Tensor feats = conv_bn(net, w, "decode_head.bottleneck.conv", "decode_head.bottleneck.bn", concat,
                         spec(concat.C, ppm_channels, 3, 1, 1, true));

  // ---- classifier (bias, no norm) in fp32, then x8 bilinear resize (+ optional sigmoid) to NCHW
  {
    const HostArray& wt = get_weight(w, "decode_head.conv_seg.weight");
    const HostArray& bs = get_weight(w, "decode_head.conv_seg.bias");
    PN_REQUIRE(wt.shape.size() == 4 && wt.shape[0] == num_classes && wt.shape[1] == ppm_channels, "conv_seg shape");
    ConvSpec s = spec(ppm_channels, num_classes, 1, 1, 1, false);
    s.out_fp32 = true;
    Tensor logits = net.arena.tensor(B, feats.H, feats.W, round_up(num_classes, 8), kF32);
    add_conv(net, "decode_head.conv_seg", feats, logits, wt.data.data(), nullptr, bs.data.data(), s);
    add_upsample_logits(net, logits, num_classes, H, W, &pn_.slots->output, &pn_.slots->apply_sigmoid);
    pn_.logits_lowres = logits;
  }
  pn_.features = x;
}

}  // namespace pn
