"""Runs the code cells of one of the REFERENCE's eval_*.ipynb notebooks against this repo, the way INTEGRATION.md section 2 says a
maintainer would: `lib/` of this repo (the overlay) first on sys.path, the reference's lib/ behind it.

The cells are the reference's own, executed verbatim (oracle/_ref/notebook_cells.json, extracted by oracle/build_ref.sh), with exactly two
substitutions that stand in for data that cannot exist offline:
  * the paths cell: `pretrained_model_file` -> a synthetic Detectron pickle (same blob names as the published checkpoints),
    `arch` -> resnet50;
  * the dataset cell (CocoDataset over COCO val2014): a two-sample synthetic dataset that yields what CocoDataset.__getitem__ yields
    ({'image': uint8 HWC, 'dbentry': {...}} through the notebook's own `preprocess_sample(...)` transform) behind the notebook's own
    DataLoader(..., collate_fn=collate_custom).
Everything else -- imports, detector(...) kwargs, model.cuda(), empty_results, the whole per-image loop (forward, postprocess_output,
add_multilevel_rois_for_test, mask_head, segm_results, extend_results) -- is the notebook's code.

The reference's pure-Python host modules come from oracle/_ref/reflib.zip (the reference tree itself does not exist on the GPU box).

  python tests/notebook_harness.py eval_mask_FPN.ipynb out.pkl [--cpu]     (--cpu: stop after the model / results cells, no .cuda())
"""
import json
import os
import pickle
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
TARGET_SIZE = 160                   # the notebooks use 800; the synthetic images are small so that the CPU oracle stays fast

# notebook -> oracle parameter-set flags
FLAGS = {
    "eval_fast.ipynb": dict(fpn=False, rpn=False, mask=False), "eval_faster.ipynb": dict(fpn=False, rpn=True, mask=False),
    "eval_mask.ipynb": dict(fpn=False, rpn=True, mask=True), "eval_fast_FPN.ipynb": dict(fpn=True, rpn=False, mask=False),
    "eval_faster_FPN.ipynb": dict(fpn=True, rpn=True, mask=False), "eval_mask_FPN.ipynb": dict(fpn=True, rpn=True, mask=True),
}
IMAGE_SIZES = [(120, 160), (150, 140)]      # two different padded shapes: exercises the per-shape engine cache with shared weights


def synthetic_rgb_image(h, w, seed):
    rng = np.random.RandomState(100 + seed)
    return np.clip(np.array([122.7717, 115.9465, 102.9801]) + 50.0 * rng.randn(h, w, 3), 0, 255).astype(np.uint8)


def synthetic_proposals(h, w, n, seed):
    rng = np.random.RandomState(200 + seed)
    x1, y1 = rng.uniform(0, w - 20, n), rng.uniform(0, h - 20, n)
    return np.stack([x1, y1, np.minimum(x1 + rng.uniform(8, w * 0.8, n), w - 1), np.minimum(y1 + rng.uniform(8, h * 0.8, n), h - 1)], 1).astype(np.float32)


def write_synthetic_pickle(path, P, flags):
    """{'blobs': {caffe2 name: ndarray}} like the published Detectron checkpoints (BGR stem), from the flat parameter dict P."""
    from detectorch_b200.model.detector import caffe2_blob_name
    blobs = {}
    for k, v in P.items():
        if k.startswith("model."):
            w = v.numpy()
            blobs[caffe2_blob_name(k[len("model."):])] = w[:, (2, 1, 0), :, :].copy() if k == "model.conv1.weight" else w

    def put(wn, bn, name):
        blobs[wn], blobs[bn] = P[name + ".weight"].numpy(), P[name + ".bias"].numpy()
    put('bbox_pred_w', 'bbox_pred_b', 'bbox_head'); put('cls_score_w', 'cls_score_b', 'classif_head')
    if flags["rpn"]:
        sfx = '_fpn2' if flags["fpn"] else ''
        put('conv_rpn%s_w' % sfx, 'conv_rpn%s_b' % sfx, 'rpn.conv_rpn')
        put('rpn_cls_logits%s_w' % sfx, 'rpn_cls_logits%s_b' % sfx, 'rpn.rpn_cls_prob')
        put('rpn_bbox_pred%s_w' % sfx, 'rpn_bbox_pred%s_b' % sfx, 'rpn.rpn_bbox_pred')
    if flags["mask"]:
        put('conv5_mask_w', 'conv5_mask_b', 'mask_head.transposed_conv'); put('mask_fcn_logits_w', 'mask_fcn_logits_b', 'mask_head.classif_logits')
        if flags["fpn"]:
            for i in range(1, 5):
                put('_[mask]_fcn%d_w' % i, '_[mask]_fcn%d_b' % i, 'mask_head.conv_head.fcn%d' % i)
    if flags["fpn"]:
        for i, kc in enumerate(("res2_2", "res3_3", "res4_5", "res5_2")):
            suffix = '_sum_lateral' if i < 3 else '_sum'
            put('fpn_inner_' + kc + suffix + '_w', 'fpn_inner_' + kc + suffix + '_b', 'conv_body.fpn_lateral.%d' % i)
            put('fpn_' + kc + '_sum_w', 'fpn_' + kc + '_sum_b', 'conv_body.fpn_output.%d' % i)
        put('fc6_w', 'fc6_b', 'conv_head.fc6'); put('fc7_w', 'fc7_b', 'conv_head.fc7')
    with open(path, "wb") as f:
        pickle.dump({'blobs': blobs}, f, protocol=2)


class SyntheticDataset(torch.utils.data.Dataset):
    """Stands in for data.coco_dataset.CocoDataset (coco_dataset.py:11-70): same attributes the notebooks touch (num_classes, __len__,
    __getitem__ -> sample_transform({'image': uint8 HWC RGB, 'dbentry': roidb entry}))."""

    def __init__(self, sample_transform, with_proposals, num_classes=81):
        self.num_classes, self.sample_transform, self.with_proposals = num_classes, sample_transform, with_proposals

    def __len__(self):
        return len(IMAGE_SIZES)

    def __getitem__(self, idx):
        h, w = IMAGE_SIZES[idx]
        boxes = synthetic_proposals(h, w, 120, idx) if self.with_proposals else np.zeros((0, 4), np.float32)
        sample = {'image': synthetic_rgb_image(h, w, idx), 'dbentry': {'boxes': boxes, 'flipped': False}}
        return self.sample_transform(sample)


def install_paths():
    sys.path.insert(0, ROOT)
    from oracle import reference_shim as rs
    if not rs.staged_available():
        raise RuntimeError("oracle/_ref/reflib.zip + notebook_cells.json are staged by oracle/build_ref.sh where the reference tree exists")
    rs.install_compat(stub_plotting=True)
    sys.path.insert(0, os.path.join(REF_DIR, "reflib.zip"))      # the reference's lib/ (data.*, utils.collate_custom, utils.utils, ...)
    sys.path.insert(0, os.path.join(ROOT, "lib"))                # the overlay, first


def find_cell(cells, needle):
    hits = [k for k in sorted(cells, key=int) if needle in cells[k]]
    if not hits:
        raise KeyError(needle)
    return cells[hits[0]]


def run(notebook, cpu_only=False, workdir=None):
    """Executes the notebook's cells; returns the notebook namespace."""
    install_paths()
    from oracle import network as net
    cells = json.load(open(os.path.join(REF_DIR, "notebook_cells.json")))[notebook]
    flags = FLAGS[notebook]
    workdir = workdir or os.getcwd()
    pkl = os.path.join(workdir, "synthetic_" + notebook.replace(".ipynb", ".pkl"))
    write_synthetic_pickle(pkl, net.synthetic_params("resnet50", **flags), flags)
    ns = {"__name__": "__main__"}
    exec(cells[min(cells, key=int)], ns)                              # imports (cell 1), verbatim
    exec(find_cell(cells, "pretrained_model_file ="), ns)             # paths (cell 3), verbatim ...
    ns["arch"], ns["pretrained_model_file"] = "resnet50", pkl         # ... then pointed at the synthetic checkpoint
    # dataset cell: the notebook's own transform + DataLoader + collate function around the synthetic samples
    tf = ns["preprocess_sample"](target_sizes=[TARGET_SIZE], fpn_on=flags["fpn"])
    ns["dataset"] = SyntheticDataset(tf, with_proposals=not flags["rpn"])
    ns["dataloader"] = ns["DataLoader"](ns["dataset"], batch_size=1, shuffle=False, num_workers=0, collate_fn=ns["collate_custom"])
    model_cell = find_cell(cells, "model = detector(")
    if cpu_only:
        model_cell = model_cell.replace("model = model.cuda()", "")
    exec(model_cell, ns)                                              # cell 7, verbatim
    exec(find_cell(cells, "empty_results"), ns)                       # cell 9, verbatim
    if not cpu_only:
        exec(find_cell(cells, "for i, batch in enumerate(dataloader)"), ns)     # the detection loop (cell 10 / 11), verbatim
    return ns


if __name__ == "__main__":
    nb, out = sys.argv[1], sys.argv[2]
    cpu = "--cpu" in sys.argv
    ns = run(nb, cpu_only=cpu, workdir=os.path.dirname(os.path.abspath(out)))
    res = {"all_boxes": ns["all_boxes"], "all_segms": ns["all_segms"], "detector_module": ns["detector"].__module__,
           "result_utils_file": ns["result_utils"].__file__, "state_dict_keys": sorted(ns["model"].state_dict().keys())}
    with open(out, "wb") as f:
        pickle.dump(res, f)
    print("NOTEBOOK OK", nb)
