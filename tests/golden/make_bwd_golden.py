"""Golden vectors of the RoIAlign BACKWARD produced by the reference's own CPU loop (lib/cppcuda/roi_align_backward_cpu.cpp:79-186, cut out and
compiled by oracle/build_ref.sh into oracle/_ref/libroialign_bwd_ref.so).  Run where /root/reference exists:
    python tests/golden/make_bwd_golden.py   ->  tests/golden/roialign_bwd_golden.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402


def cases():
    rng = np.random.RandomState(11)
    out = []
    for (PH, sr, cols, scale, B, C, H, W, R) in ((7, 2, 5, 0.25, 2, 6, 24, 30, 48), (14, 0, 4, 0.0625, 1, 5, 20, 28, 32), (7, 2, 4, 0.03125, 1, 4, 13, 19, 24),
                                                (14, 2, 5, 0.125, 3, 3, 16, 22, 40)):
        x1, y1 = rng.uniform(-30, W / scale, R), rng.uniform(-30, H / scale, R)
        r = np.stack([rng.randint(0, B, R).astype(np.float32), x1, y1, x1 + rng.uniform(0, W / scale * 0.9, R), y1 + rng.uniform(0, H / scale * 0.9, R)], 1)
        r[0, 1:] = [5.0, 5.0, 5.2, 5.1]                       # degenerate (forced to 1x1)
        r[1, 1:] = [-50.0, -50.0, W / scale + 40, H / scale + 40]   # larger than the map: samples outside contribute nothing
        r = r.astype(np.float32)
        if cols == 4:
            r = np.ascontiguousarray(r[:, 1:])
        out.append(dict(PH=PH, sr=sr, scale=scale, shape=(B, C, H, W), rois=r, top=rng.randn(R, C, PH, PH).astype(np.float32)))
    return out


if __name__ == "__main__":
    d = {}
    for i, c in enumerate(cases()):
        g = ref.roi_align_backward_ref(c["top"], c["rois"], c["shape"], c["PH"], c["PH"], c["scale"], c["sr"])
        d.update({"rois%d" % i: c["rois"], "top%d" % i: c["top"], "grad%d" % i: g,
                  "meta%d" % i: np.array([c["PH"], c["sr"], c["shape"][0], c["shape"][1], c["shape"][2], c["shape"][3]], np.int64),
                  "scale%d" % i: np.float32(c["scale"])})
    d["n"] = np.int64(len(cases()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "roialign_bwd_golden.npz"), **d)
    print("written", {k: v.shape for k, v in d.items() if k.startswith("grad")})
