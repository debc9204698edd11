#!/usr/bin/env python
"""Row-sharded association on N GPUs (torchrun): parity of the all-gathered result against the oracle and
device timing (max over ranks).  BASELINE config 5 shape: 1000 keypoints vs 50 000 map points."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppg_slam_b200 import cameras, capi, synth  # noqa: E402
from ppg_slam_b200.sharded import ShardedAssociator  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cam = cameras.UMA
    n, m = 1000, 50000
    rs = np.random.RandomState(11)
    gx, gy = np.meshgrid(np.arange(8, cam.width - 8, 6), np.arange(8, cam.height - 8, 6))
    sel = rs.choice(gx.size, n, replace=False)
    kx = gx.ravel()[sel].astype(np.float32)
    ky = gy.ravel()[sel].astype(np.float32)
    fd = rs.normal(size=(n, 256)).astype(np.float32)
    fd /= np.linalg.norm(fd, axis=1, keepdims=True)
    inp = synth.association_inputs(5, fd, np.stack([kx, ky], 1), m, cam.width, cam.height)
    free = np.ones(n, np.uint8)
    ex = capi.Extractor(cam, device=local, max_batch=1, max_map_points=m)
    sa = ShardedAssociator.from_extractor(ex, inp["map_desc"], inp["proj_uv"], inp["view_cos"], 10.0, 0.8, dev)
    sa.stage_frame(kx, ky, fd, free)
    got = sa.run()
    ok = None
    if rank == 0:
        from oracle import post_ref as O
        ref = O.search_all(cam, kx, ky, fd, free, inp["map_desc"], inp["proj_uv"], inp["view_cos"], 10.0, 0.8)
        ok = all(np.array_equal(got[k], ref[k]) for k in ("best_idx", "second_idx", "accept")) and \
            np.array_equal(got["best_d"].view(np.uint32), ref["best_d"].view(np.uint32))
    for _ in range(3):
        sa.run()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 20
    for _ in range(K):
        sa.run()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / K], device=dev, dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    # device-only part of one rank's shard
    ex.timer_start()
    for _ in range(K):
        ex.assoc_run()
    ms = ex.timer_stop() / K
    if rank == 0:
        print(json.dumps({"sharded_assoc": {"world": world, "rows": m, "keypoints": n, "parity_vs_oracle": bool(ok),
                                            "ms_per_call_incl_allgather_and_d2h": float(dt.item()) * 1e3,
                                            "shard_device_ms": ms, "fallback_rows": ex.assoc_fallback_rows()}}),
              flush=True)
    ex.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
