#!/usr/bin/env python
"""bench_sharded.py -- BASELINE.json configs[3]: one 2048x2048 image, 32 templates 48x48 sharded over
the ranks (contiguous slices), ONE all-gather of the hit rows over NCCL, replicated global NMS.

    torchrun --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 bench_sharded.py --steps 50

``--mode images --workload C5``: BASELINE.json configs[4] -- a batch of 16 images sharded over the ranks, the whole
template list on every rank, one all-gather of the final per-image hit lists.

Strong scaling (fixed total work).  Every rank checks that the gathered result equals the single-GPU
``MTM.matchTemplates`` answer; rank 0 prints one JSON line.  Host-side timing of the whole sharded call
(the collective makes a pure device-event timing ill-defined), max over ranks.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--mode", default="templates", choices=["templates", "images"],
                    help="templates: one image, template list sharded (configs[3]); images: --images images sharded over the "
                         "ranks, whole template list everywhere, one all-gather of the final hit lists (configs[4], use --workload C5)")
    ap.add_argument("--images", type=int, default=16)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import MTM
    from mtm_b200 import sharded
    import workloads as synth
    if args.mode == "images":
        return bench_images(args, MTM, sharded, synth, torch, dist, rank, world)
    image, temps, params = synth.config(args.workload)
    single = MTM.matchTemplates(temps, image, **params)
    for _ in range(args.warmup):
        got = sharded.matchTemplatesSharded(temps, image, **params)
    same = [(h[0], h[1]) for h in got] == [(h[0], h[1]) for h in single] and all(a[2] == b[2] for a, b in zip(got, single))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        got = sharded.matchTemplatesSharded(temps, image, **params)
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / args.steps, 0.0 if same else 1.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "template-matches/sec", "workload": args.workload + " sharded over ranks (templates)",
                          "n_gpus": world, "value": len(temps) / float(dt[0]), "unit": "matches/s",
                          "ms_per_call": float(dt[0]) * 1e3, "scaling": "strong", "hits": len(got),
                          "identical_to_single_gpu": bool(dt[1].item() == 0.0),
                          "exchange": "all_reduce(MAX) of counts + one all_gather_into_tensor of 6 x int32 hit rows (NCCL)"}))
    if world > 1:
        dist.destroy_process_group()


def bench_images(args, MTM, sharded, synth, torch, dist, rank, world):
    """BASELINE.json configs[4]: a batch of images (seeds 0..n-1 of the workload) sharded over the ranks."""
    images, temps, params = [], None, None
    for k in range(args.images):
        im, temps, params = synth.config(args.workload, seed=0, image_index=k)
        images.append(im)
    got = None
    for _ in range(max(1, args.warmup)):
        got = sharded.matchTemplatesBatchSharded(temps, images, **params)
    # every rank checks its own slice against the per-image call on its GPU
    lo, hi = sharded.shard_bounds(len(images), world, rank)
    same = all([(a[0], a[1], float(a[2])) for a in got[i]] == [(b[0], b[1], float(b[2])) for b in MTM.matchTemplates(temps, images[i], **params)]
               for i in range(lo, hi))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        got = sharded.matchTemplatesBatchSharded(temps, images, **params)
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / args.steps, 0.0 if same else 1.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "template-matches/sec", "workload": "%s: %d images sharded over ranks" % (args.workload, len(images)),
                          "n_gpus": world, "value": len(temps) * len(images) / float(dt[0]), "unit": "matches/s",
                          "ms_per_batch": float(dt[0]) * 1e3, "scaling": "strong", "hits": sum(len(h) for h in got),
                          "identical_to_single_gpu": bool(dt[1].item() == 0.0),
                          "exchange": "all_reduce(MAX) of counts + one all_gather_into_tensor of 7 x int32 rows of the final hit lists (NCCL)"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
