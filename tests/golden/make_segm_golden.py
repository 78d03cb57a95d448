"""Generates tests/golden/segm_golden.npz by RUNNING THE REFERENCE's own segm_results
(lib/utils/result_utils.py:170-228, imported unmodified through oracle/reference_shim.py) with the real
cv2.resize of this container (OpenCV 4.13, IPP switched off so that OpenCV's own kernel runs; the IPP kernel differs
by <= 1.6e-6 in value).  pycocotools is not installed: mask_util.encode is replaced by a recorder that captures the
pasted im_mask the reference hands to it; the RLE strings stored next to the masks come from oracle/ref.py.

    python tests/golden/make_segm_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shim as rs  # noqa: E402
from oracle import ref as oref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def cases(rng, im_h, im_w):
    """reference boxes (fp32, image px): ordinary, border-touching, tiny, exactly 2x-downscale (14x14 -> expands to 15x15), full image"""
    b = [[20.3, 30.7, 120.9, 150.2], [0, 0, 60.5, 40.5], [im_w - 50.2, im_h - 70.9, im_w - 1, im_h - 1], [100, 100, 100, 100],
         [150.5, 20.5, 151.0, 90.0], [50, 60, 63, 73], [0, 0, im_w - 1, im_h - 1], [200.2, 5.1, 280.8, 190.3],
         [-5.5, -7.5, 30.2, 25.1], [im_w - 20.0, 50.0, im_w + 15.0, 120.0], [10.0, im_h - 30.0, 90.0, im_h + 12.0], [33.3, 44.4, 47.0, 58.1]]
    for _ in range(12):
        x1, y1 = rng.uniform(0, im_w - 2), rng.uniform(0, im_h - 2)
        b.append([x1, y1, min(x1 + np.exp(rng.uniform(0, np.log(im_w))), im_w - 1), min(y1 + np.exp(rng.uniform(0, np.log(im_h))), im_h - 1)])
    b = np.asarray(b, dtype=np.float32)
    # detections reaching segm_results are clipped to the image (result_utils.py:90 clip_tiled_boxes); the reference itself
    # raises on boxes that lie outside it (negative slice bounds), so only the (M+2)/M expansion may cross the border
    b[:, 0::2] = np.clip(b[:, 0::2], 0, im_w - 1)
    b[:, 1::2] = np.clip(b[:, 1::2], 0, im_h - 1)
    return b


def main():
    rs.install()
    import utils.result_utils as ru
    cv2.ipp.setUseIPP(False)
    captured = []

    def encode(arr):
        captured.append(np.ascontiguousarray(arr[:, :, 0]))
        return [{'size': list(arr.shape[:2]), 'counts': oref.rle_to_string(oref.rle_encode(arr[:, :, 0]))}]
    ru.mask_util.encode = encode
    G = {}
    for (tag, M, K, im_h, im_w) in (("m28", 28, 4, 200, 300), ("m14", 14, 3, 97, 131)):
        rng = np.random.RandomState(M)
        boxes = cases(rng, im_h, im_w)
        D = len(boxes)
        cls = np.sort(rng.randint(1, K, D))
        # smooth-ish blobs so the binary masks have structure, plus noise so thresholds are crossed in many places
        yy, xx = np.mgrid[0:M, 0:M].astype(np.float32) / M
        masks = np.zeros((D, K, M, M), np.float32)
        for d in range(D):
            for k in range(K):
                cx, cy, r = rng.uniform(.3, .7), rng.uniform(.3, .7), rng.uniform(.15, .45)
                blob = 1.0 / (1.0 + np.exp(((xx - cx) ** 2 + (yy - cy) ** 2 - r * r) * 40))
                masks[d, k] = np.clip(blob + 0.25 * rng.randn(M, M), 0, 1)
        cls_boxes = [[] for _ in range(K)]
        for j in range(1, K):
            sel = cls == j
            cls_boxes[j] = np.hstack([boxes[sel], np.ones((sel.sum(), 1), np.float32)])
        del captured[:]
        segms = ru.segm_results(cls_boxes, masks, boxes, im_h, im_w, num_classes=K, M=M)
        pasted = np.stack(captured)
        assert pasted.shape == (D, im_h, im_w)
        # the oracle restatement must reproduce the reference run bit for bit
        mine = oref.segm_results(cls_boxes, masks, boxes, im_h, im_w, num_classes=K, M=M)
        assert mine == segms
        G[tag + "_masks"], G[tag + "_boxes"], G[tag + "_cls"] = masks, boxes, cls.astype(np.int32)
        G[tag + "_size"] = np.array([im_h, im_w], np.int32)
        G[tag + "_pasted_bits"] = np.packbits(pasted.reshape(D, -1), axis=1)
        G[tag + "_rle"] = np.array([s['counts'] for j in range(1, K) for s in segms[j]])
        print(tag, "dets", D, "pixels set", int(pasted.sum()))
    np.savez_compressed(os.path.join(OUT, "segm_golden.npz"), **G)
    print("segm_golden.npz", os.path.getsize(os.path.join(OUT, "segm_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
