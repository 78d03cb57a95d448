"""Generates tests/golden/prep_golden.npz by RUNNING THE REFERENCE's own lib/utils/blob.py (prep_im_for_blob + im_list_to_blob,
imported unmodified through oracle/reference_shim.py) with the real cv2 of this container (OpenCV 4.13, IPP off = OpenCV's own
resize kernel).  Small images keep the file small; the scale logic is exercised with small target sizes.

    python tests/golden/make_prep_golden.py
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
# (H, W, target_size, max_size): up-scale, max_size cap, exact 2x down-scale (area-fast path), down-scale, portrait
CASES = [(30, 40, 50, 84), (22, 75, 50, 84), (100, 120, 50, 84), (60, 45, 32, 50), (48, 30, 50, 1333)]


def main():
    rs.install()
    import utils.blob as rb
    cv2.ipp.setUseIPP(False)
    G = {"cases": np.asarray(CASES, np.int32)}
    for i, (h, w, ts, ms) in enumerate(CASES):
        rng = np.random.RandomState(100 + i)
        im = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        ims, scales = rb.prep_im_for_blob(im.copy(), target_sizes=[ts], max_size=ms)
        blob = rb.im_list_to_blob(ims, fpn_on=True)
        mine_ims, mine_scales = oref.prep_im_for_blob(im, target_sizes=[ts], max_size=ms)
        assert mine_scales == scales and np.array_equal(oref.im_list_to_blob(mine_ims, True), blob), (h, w)
        G["im%d" % i], G["blob%d" % i], G["scale%d" % i] = im, np.ascontiguousarray(blob), np.float64(scales[0])
        print((h, w), "->", blob.shape, "scale", scales[0])
    np.savez_compressed(os.path.join(OUT, "prep_golden.npz"), **G)
    print("prep_golden.npz", os.path.getsize(os.path.join(OUT, "prep_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
