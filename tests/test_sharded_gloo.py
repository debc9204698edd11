"""world_size-2 gloo test of the row-sharded association host logic (shard split, packing, all-gather,
reassembly).  The per-shard compute is the oracle, so the gathered answer must equal the un-sharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from ppg_slam_b200.sharded import pack_records, shard_rows, unpack_records


def test_shard_rows_partition():
    for n, w in [(50000, 8), (7, 2), (5, 8), (1024, 4), (3, 3)]:
        sh = shard_rows(n, w)
        assert len(sh) == w and sh[0][0] == 0
        assert sum(k for _, k in sh) == n
        for (a, ka), (b, _) in zip(sh, sh[1:]):
            assert a + ka == b
        assert max(k for _, k in sh) - min(k for _, k in sh) <= 1
    assert shard_rows(50000, 8)[0] == (0, 6250)


def test_pack_roundtrip():
    rs = np.random.RandomState(0)
    r = dict(best_idx=rs.randint(-1, 500, 100).astype(np.int32), second_idx=rs.randint(-1, 500, 100).astype(np.int32),
             best_d=rs.rand(100).astype(np.float32), second_d=np.full(100, 1e6, np.float32),
             accept=rs.randint(0, 2, 100).astype(np.uint8))
    u = unpack_records(pack_records(r))
    for k in r:
        np.testing.assert_array_equal(u[k], r[k])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rows, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import post_ref as O
        from ppg_slam_b200 import cameras, synth
        from ppg_slam_b200.sharded import ShardedAssociator
        cam = cameras.EUROC
        rs = np.random.RandomState(4)
        n = 150
        kx = rs.uniform(10, cam.width - 10, n).astype(np.float32)
        ky = rs.uniform(10, cam.height - 10, n).astype(np.float32)
        fd = rs.normal(size=(n, 256)).astype(np.float32)
        fd /= np.linalg.norm(fd, axis=1, keepdims=True)
        inp = synth.association_inputs(6, fd, np.stack([kx, ky], 1), n_rows, cam.width, cam.height)
        free = np.ones(n, np.uint8)

        def compute(row0, rows):
            return O.search_all(cam, kx, ky, fd, free, inp["map_desc"][row0:row0 + rows],
                                inp["proj_uv"][row0:row0 + rows], inp["view_cos"][row0:row0 + rows], 10.0, 0.8)

        sa = ShardedAssociator(n_rows, compute)
        got = sa.run()
        full = O.search_all(cam, kx, ky, fd, free, inp["map_desc"], inp["proj_uv"], inp["view_cos"], 10.0, 0.8)
        ok = all(np.array_equal(got[k], full[k]) for k in ("best_idx", "second_idx", "accept")) and \
            np.array_equal(got["best_d"].view(np.uint32), full["best_d"].view(np.uint32))
        q.put((rank, bool(ok), int(full["accept"].sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [1001, 64])
def test_sharded_association_world2_gloo(n_rows):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rows, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), res
    assert res[0][2] > 0
