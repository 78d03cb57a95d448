"""SURVEY.md 8e contract on hardware, run under torchrun (one process per GPU, NCCL):
16 distinct images are sharded over the ranks (detectorch_b200.sharding.shard_range), every rank runs the fused engine on its
shard, the padded per-image results are gathered with sharding.gather_results over NCCL, and rank 0 asserts that the gathered
tensors are BIT-IDENTICAL to a single-GPU run over all 16 images in image order.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tests/dist_gather_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TOTAL, HH, WW = 16, 128, 160


def run_images(imgs, dev, P):
    from detectorch_b200.engine import Engine
    eng = Engine(arch="resnet50", batch=imgs.size(0), height=HH, width=WW, det_cap=100, device=dev)
    eng.load_state_dict(P)
    eng.run(imgs.to(dev), 1.0)
    torch.cuda.synchronize(dev)
    eng.check_range()
    B = imgs.size(0)
    return {"boxes": eng.buffer("det_boxes").clone(), "scores": eng.buffer("det_scores").clone(), "classes": eng.buffer("det_classes").clone(),
            "counts": eng.buffer("det_counts").clone(), "masks": eng.buffer("masks").view(B, 100, 28, 28).clone()}


def main():
    from detectorch_b200.sharding import gather_results, shard_range
    from oracle import network as net          # synthetic weights / images only
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    P = net.synthetic_params("resnet50")
    imgs = torch.cat([net.synthetic_image(1, HH, WW, seed=40 + i) for i in range(TOTAL)], 0)
    b, e = shard_range(TOTAL, world, rank)
    mine = run_images(imgs[b:e].contiguous(), dev, P)
    full = gather_results(mine)
    ok = True
    if rank == 0:
        want = run_images(imgs, dev, P)
        res = {k: bool(torch.equal(full[k], want[k])) for k in want}
        ok = all(res.values())
        print(json.dumps({"check": "N-GPU gather == single-GPU run, bit-exact", "world": world, "images": TOTAL, "equal": res,
                          "detections_total": int(want["counts"].sum().item()), "nccl": True}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
