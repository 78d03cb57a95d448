"""Host-side mirror of lib/utils/multilevel_rois.py:19-82 (numpy in / numpy out, not on the hot path): distributes a set of RoIs over the
FPN levels for the callers that hand PRE-COMPUTED boxes to the model -- preprocess_sample (Fast R-CNN proposals, eval_fast_FPN.ipynb) and
the notebook's mask step (eval_mask_FPN.ipynb cell 10).  Inside the engine the same mapping runs on the device (collect_kernel,
mask_rois_kernel)."""
import numpy as np


def map_rois_to_fpn_levels(rois, k_min, k_max, roi_canonical_scale=224, roi_canonical_level=4):
    """FPN paper eqn. (1) as multilevel_rois.py:41-53 evaluates it: floor(4 + log2(sqrt(area) / 224 + 1e-6)) clipped to [k_min, k_max],
    area with the +1 pixel convention (boxes.py:75-81)."""
    rois = np.asarray(rois)
    side = np.sqrt((rois[:, 2] - rois[:, 0] + 1) * (rois[:, 3] - rois[:, 1] + 1))
    return np.clip(np.floor(roi_canonical_level + np.log2(side / roi_canonical_scale + 1e-6)), k_min, k_max)


def add_multilevel_rois_for_test(blobs, name, roi_min_level=2, roi_max_level=5):
    """blobs[name] [N,4|5] -> adds blobs[name + '_fpn<l>'] (the rows of level l, ascending original index) for l in [min, max] and
    blobs[name + '_idx_restore_int32'] (argsort of the concatenated original indices: cat(levels)[restore] == blobs[name])."""
    rois = blobs[name]
    lvls = map_rois_to_fpn_levels(rois, roi_min_level, roi_max_level)
    order = []
    for lvl in range(roi_min_level, roi_max_level + 1):
        idx = np.where(lvls == lvl)[0]
        blobs[name + '_fpn' + str(lvl)] = rois[idx, :]
        order.append(idx)
    order = np.concatenate(order) if order else np.empty((0,), np.int64)
    blobs[name + '_idx_restore_int32'] = np.argsort(order, kind='stable').astype(np.int32, copy=False)
    return blobs
