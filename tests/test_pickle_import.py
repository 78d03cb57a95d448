"""CPU test of the Detectron caffe2-pickle weight import (SURVEY.md 8f "next" row 2): a synthetic pickle written with the
reference's own blob names loads identically into the reference detector and into the detectorch_b200 mirror.
Needs the reference tree (its utils.utils.parse_th_to_caffe2 name mapping), so it is skipped on the GPU box."""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import network as net
from oracle import reference_shim as rs

FPN_KW = dict(arch='resnet50', conv_body_layers=['conv1', 'bn1', 'relu', 'maxpool', 'layer1', 'layer2', 'layer3', 'layer4'],
              conv_head_layers='two_layer_mlp', fpn_layers=['layer1', 'layer2', 'layer3', 'layer4'], fpn_extra_lvl=True,
              roi_height=7, roi_width=7, roi_spatial_scale=[0.25, 0.125, 0.0625, 0.03125], roi_sampling_ratio=2,
              use_rpn_head=True, use_mask_head=True, mask_head_type='1up4convs')


@pytest.mark.skipif(not rs.available(), reason="reference tree not present")
def test_detectron_pickle_loads_like_the_reference(tmp_path, built):
    rs.install()
    from model.detector import detector as ref_detector                 # the reference
    from utils.utils import parse_th_to_caffe2
    from detectorch_b200.model.detector import detector as b200_detector
    P = net.synthetic_params("resnet50")
    ref0 = ref_detector(roi_feature_channels=1024, **FPN_KW)
    blobs = {}
    for k in ref0.model.state_dict().keys():                            # trunk: torchvision name -> caffe2 blob name
        if 'running' in k or 'fc' in k:
            continue
        if 'num_batches' in k:
            # torch >= 0.4.1 adds this buffer; the reference's loader (detector.py:300-304) does not skip it, so a blob must exist
            blobs[parse_th_to_caffe2(k.split('.'))] = np.zeros((), np.float32)
            continue
        w = P["model." + k].numpy()
        blobs[parse_th_to_caffe2(k.split('.'))] = w[:, (2, 1, 0), :, :].copy() if k == 'conv1.weight' else w     # pickles hold BGR
    def put(wn, bn, name):
        blobs[wn], blobs[bn] = P[name + ".weight"].numpy(), P[name + ".bias"].numpy()
    put('bbox_pred_w', 'bbox_pred_b', 'bbox_head'); put('cls_score_w', 'cls_score_b', 'classif_head')
    put('conv_rpn_fpn2_w', 'conv_rpn_fpn2_b', 'rpn.conv_rpn')
    put('rpn_cls_logits_fpn2_w', 'rpn_cls_logits_fpn2_b', 'rpn.rpn_cls_prob'); put('rpn_bbox_pred_fpn2_w', 'rpn_bbox_pred_fpn2_b', 'rpn.rpn_bbox_pred')
    put('conv5_mask_w', 'conv5_mask_b', 'mask_head.transposed_conv'); put('mask_fcn_logits_w', 'mask_fcn_logits_b', 'mask_head.classif_logits')
    for i in range(1, 5):
        put('_[mask]_fcn%d_w' % i, '_[mask]_fcn%d_b' % i, 'mask_head.conv_head.fcn%d' % i)
    for i, l in enumerate(FPN_KW['fpn_layers']):
        kc = parse_th_to_caffe2((l + '.' + list(getattr(ref0.model, l).state_dict().keys())[-1]).split('.'))
        kc = kc[:kc.rfind("_")]
        suffix = '_sum_lateral' if i < 3 else '_sum'
        put('fpn_inner_' + kc + suffix + '_w', 'fpn_inner_' + kc + suffix + '_b', 'conv_body.fpn_lateral.%d' % i)
        put('fpn_' + kc + '_sum_w', 'fpn_' + kc + '_sum_b', 'conv_body.fpn_output.%d' % i)
    put('fc6_w', 'fc6_b', 'conv_head.fc6'); put('fc7_w', 'fc7_b', 'conv_head.fc7')
    pkl = os.path.join(str(tmp_path), "model_final.pkl")
    with open(pkl, "wb") as f:
        pickle.dump({'blobs': blobs}, f, protocol=2)

    ref = ref_detector(detector_pkl_file=pkl, roi_feature_channels=1024, **FPN_KW)
    mine = b200_detector(detector_pkl_file=pkl, **FPN_KW)
    sr, sm = ref.state_dict(), mine.state_dict()
    for k, v in P.items():
        assert torch.equal(sr[k], v), "reference loader: " + k           # the synthetic pickle round-trips through the reference
        assert torch.equal(sm[k], v), "mirror loader: " + k
    # and the engine's parameter table accepts exactly these names
    from detectorch_b200 import engine
    t = engine.param_table("resnet50")
    assert set(t) == set(P) and all(t[k] == int(np.prod(P[k].shape)) for k in P)
