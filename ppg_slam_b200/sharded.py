"""Row-sharded association across GPUs: the map-descriptor table is split by rows over the
ranks, every rank scores its rows against the (replicated) frame, and ONE all-gather of the per-row top-2
records rebuilds the full answer.  Rows are independent, so the gather is a concatenation.

The frame path itself (extract + graph + descriptors) is sharded by frame with no communication; this module
is the only place a collective exists.  On the GPU box the exchange is the library's own: ppg_assoc_allgather
(csrc/comm.cu) packs the records and enqueues one ncclAllGather on the ctx stream right behind the scoring kernels,
no host synchronisation in between; `torch.distributed` only carries the NCCL unique id to the ranks.  The CPU tests
of the host logic (gloo) pass a `compute` callable and gather through torch.distributed.
"""
import numpy as np
import torch
import torch.distributed as dist

RECORD_WORDS = 5  # best_idx, second_idx, best_dist (bits), second_dist (bits), accept


def shard_rows(n_rows, world):
    """Contiguous split; the first (n_rows % world) ranks get one extra row.  -> list of (row0, rows)."""
    base, extra = divmod(n_rows, world)
    out, r0 = [], 0
    for r in range(world):
        k = base + (1 if r < extra else 0)
        out.append((r0, k))
        r0 += k
    return out


def pack_records(res):
    """dict of per-row arrays -> (rows, 5) int32 (floats bit-cast)."""
    n = len(res["best_idx"])
    rec = np.empty((n, RECORD_WORDS), np.int32)
    rec[:, 0] = res["best_idx"]
    rec[:, 1] = res["second_idx"]
    rec[:, 2] = np.asarray(res["best_d"], np.float32).view(np.int32)
    rec[:, 3] = np.asarray(res["second_d"], np.float32).view(np.int32)
    rec[:, 4] = res["accept"]
    return rec


def unpack_records(rec):
    rec = np.ascontiguousarray(rec, np.int32)
    return dict(best_idx=rec[:, 0].copy(), second_idx=rec[:, 1].copy(), best_d=rec[:, 2].copy().view(np.float32),
                second_d=rec[:, 3].copy().view(np.float32), accept=rec[:, 4].astype(np.uint8))


class _DevArray:
    """Wraps a raw device pointer so that torch.as_tensor can adopt it without a copy."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class ShardedAssociator:
    """compute(row0, rows) -> dict(best_idx, second_idx, best_d, second_d, accept) for this rank's rows.

    For the GPU path build it with `from_extractor`; the CPU tests pass the oracle as `compute`."""

    def __init__(self, n_rows, compute, device=None, group=None):
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.group = group
        self.n_rows = n_rows
        self.shards = shard_rows(n_rows, self.world)
        self.row0, self.rows = self.shards[self.rank]
        self.max_rows = max(k for _, k in self.shards)
        self.compute = compute
        self.device = device if device is not None else torch.device("cpu")
        self._send = torch.zeros((self.max_rows, RECORD_WORDS), dtype=torch.int32, device=self.device)
        self._recv = torch.zeros((self.world * self.max_rows, RECORD_WORDS), dtype=torch.int32, device=self.device)

    def run(self):
        """Scores this rank's rows, all-gathers, returns the full per-row result on every rank."""
        res = self.compute(self.row0, self.rows)
        if isinstance(res, torch.Tensor):  # already a packed device tensor (rows, 5)
            self._send[:self.rows].copy_(res)
        else:
            self._send[:self.rows].copy_(torch.from_numpy(pack_records(res)))
        dist.all_gather_into_tensor(self._recv, self._send, group=self.group)
        full = self._recv.view(self.world, self.max_rows, RECORD_WORDS)
        parts = [full[r, :k] for r, (_, k) in enumerate(self.shards)]
        return unpack_records(torch.cat(parts, 0).cpu().numpy())

    @classmethod
    def from_extractor(cls, ex, map_desc_full, proj_uv_full, view_cos_full, th, ratio, device, group=None):
        """GPU path: uploads this rank's rows of the table once and creates the library's NCCL communicator; run()
        then scores the frame staged with `stage_frame` and all-gathers the records on the ctx stream (comm.cu)."""
        from . import capi
        n_rows = len(map_desc_full)
        self = cls(n_rows, None, device=device, group=group)
        r0, k = self.row0, self.rows
        uid = [capi.comm_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, group=group)
        ex.comm_init(uid[0], self.rank, self.world)
        if k > 0:  # a rank without rows (fewer rows than ranks) uploads nothing and only takes part in the gather
            ex.upload_map(map_desc_full[r0:r0 + k])
        self._ex, self._proj, self._vcos, self._th, self._ratio = ex, proj_uv_full[r0:r0 + k], view_cos_full[r0:r0 + k], th, ratio
        self.gather_us = None

        def run():
            if self.rows > 0:
                ex.assoc_run()
            ex.assoc_allgather(self.rows, self.max_rows)
            rec, self.gather_us = ex.assoc_allgather_fetch()
            parts = [rec[r, :kk] for r, (_, kk) in enumerate(self.shards)]
            return unpack_records(np.concatenate(parts, 0))

        self.run = run
        return self

    def stage_frame(self, kp_x, kp_y, frame_desc, free_mask):
        if self.rows > 0:
            self._ex.assoc_stage(kp_x, kp_y, frame_desc, free_mask, self._proj, self._vcos, self._th, self._ratio)
