"""Mirror of the hot-path part of the reference lib/utils/result_utils.py: postprocess_output (:76-94)
and box_results_with_nms_and_limit (:96-168), computed by the engine's on-device detection stage
(decode with weights (10,10,5,5), clip, score > 0.05, per-class NMS 0.5, top-100) and returned in the
reference's numpy structures."""
import numpy as np
import torch

from ..engine import ST_DETECT

_active = {"engine": None}


def set_active_engine(engine):
    _active["engine"] = engine


def _as_float(x):
    if torch.is_tensor(x):
        return float(x.reshape(-1)[0])
    return float(np.asarray(x).reshape(-1)[0])


def postprocess_output(rois, scaling_factor, im_size, class_scores, bbox_deltas, bbox_reg_weights=(10.0, 10.0, 5.0, 5.0)):
    """-> (scores_final [D], boxes_final [D,4], boxes_per_class list[81] of [n_j,5]) numpy, like result_utils.py:76-94."""
    eng = _active["engine"]
    if eng is None:
        raise RuntimeError("postprocess_output: run detector.forward first (the engine owns the device buffers)")
    if tuple(bbox_reg_weights) != (10.0, 10.0, 5.0, 5.0):
        raise NotImplementedError("only the reference's default bbox_reg_weights are built")
    R = eng.cfg.post_nms_top_n
    n = rois.shape[-2] if rois.dim() == 3 else rois.shape[0]
    if n > R:
        raise RuntimeError("postprocess_output: %d RoIs exceed the engine capacity %d" % (n, R))
    r = rois.reshape(-1, rois.shape[-1]).to(eng.device).float()
    er, ec, eb, en = eng.buffer("rois"), eng.buffer("cls_prob"), eng.buffer("bbox_pred"), eng.buffer("roi_counts")
    # inputs that already are the engine's own views are left in place; anything else is copied in
    if r.data_ptr() != er[0, :, 1:5].data_ptr():
        er[0].zero_(); er[0, :n, 1:5] = r[:, -4:]
    if class_scores.data_ptr() != ec.data_ptr():
        ec[:n] = class_scores.to(eng.device).float()
    if bbox_deltas.data_ptr() != eb.data_ptr():
        eb[:n] = bbox_deltas.to(eng.device).float()
    en[0] = n
    sf = _as_float(scaling_factor)
    im = (im_size.detach().cpu().numpy() if torch.is_tensor(im_size) else np.asarray(im_size)).squeeze()
    eng.set_original_size(float(im[0]), float(im[1]))
    eng.run(None, sf, ST_DETECT, ST_DETECT)
    cnt = int(eng.buffer("det_counts")[0].item())
    boxes = eng.buffer("det_boxes")[0, :cnt].cpu().numpy()
    scores = eng.buffer("det_scores")[0, :cnt].cpu().numpy()
    classes = eng.buffer("det_classes")[0, :cnt].cpu().numpy()
    NC = eng.cfg.num_classes
    cls_boxes = [[] for _ in range(NC)]
    for j in range(1, NC):
        m = classes == j
        cls_boxes[j] = np.hstack((boxes[m], scores[m][:, None])).astype(np.float32, copy=False)
    return scores, boxes, cls_boxes
