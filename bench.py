#!/usr/bin/env python
"""Benchmark of the PPG-SLAM front-end hot path on B200 (BASELINE.json metric: frames/sec extract+associate
at 752x480 on 1/2/4/8 B200).

  python bench.py --gpus N --steps K --warmup W            # this repo (libppg_b200.so through the C ABI)
  python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle port) on host cores
  torchrun --nproc-per-node N ... bench.py --gpus N ...     # N>1: one rank per GPU, frames sharded, no collective

One step = one batch of 32 synthetic EuRoC-shaped frames per GPU through the whole path:
networks (tcgen05 convolutions) -> keypoints -> point-pair graph -> descriptors -> the whole
Matcher::ExtendMapMatches of every frame against a resident map of M points + its edge graph (window search with
the live frame state, assignment, seed growing; --assoc core: the frozen-state search core only; --frustum:
Frame::CheckInFrustum on the device instead of staged projections).  Every frame has its own local map inside the
resident table (synth.extend_inputs_multi), so every walk accepts and grows a few hundred matches.
`value` times the device work with the frames already in HBM; `e2e` goes through the host-facing pipelined calls with
HOST buffers in pinned memory (ppg_extract_async / ppg_assoc_stage_batch_async / ppg_extend_run_batch /
ppg_extend_fetch_batch_async, then ppg_extract_wait + ppg_extend_collect), ONE host thread per GPU keeping
--e2e-streams contexts in flight, all copies inside the timed region.
With N > 1 ranks a second leg times BASELINE config 5: the map table row-sharded over the ranks, one ncclAllGather of
the per-row top-2 records on the ctx stream (`sharded_assoc` in the JSON line).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 32
MAP_ROWS = 8192
TH, RATIO = 10.0, 0.8
GFLOP_PER_FRAME = 66.633          # conv MACs x 2 at 752x480
CONV1B_GFLOP_PER_FRAME = 2 * 13.307  # 64->64 3x3 at full resolution
CONV1A_GFLOP_PER_FRAME = 2 * 0.208   # 1->64 3x3 at full resolution


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1399.0), d.get("hbm_gbs", 6538.6), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 6:
                continue
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


CAMERA = "EuRoC"  # --camera: EuRoC (the benchmark configuration), TUM-VI, TUM-VI-1024, UMA-VI (other BASELINE shapes)


def _cam():
    from ppg_slam_b200 import cameras
    return cameras.ALL[CAMERA]


def make_workload(batch, seed0=0):
    from ppg_slam_b200 import synth
    cam = _cam()
    frames = [synth.frame(seed0 + s, cam.width, cam.height) for s in range(batch)]
    return cam, frames


def make_assoc_inputs(cam, recs, rows, n_slices=None):
    """The resident map of the step: one slice of `rows / batch` map points planted around EVERY frame's keypoints
    (descriptors, projections, map edges mirroring the frame's point-pair graph), see synth.extend_inputs_multi."""
    from ppg_slam_b200 import synth
    return synth.extend_inputs_multi(17, recs, rows, cam.width, cam.height, th=TH, n_slices=n_slices)


def make_geometry(cam, base, n_frames):
    """--frustum: world points that project (identity pose) where the synthetic projections are, and one slightly
    perturbed pose per frame, so that Frame::CheckInFrustum runs on the device instead of staging projections."""
    from ppg_slam_b200 import synth
    M = len(base["view_cos"])
    rs = np.random.RandomState(5)
    z = rs.uniform(2.0, 9.0, M).astype(np.float32)
    fx, fy, cx, cy = cam.K[0], cam.K[4], cam.K[2], cam.K[5]
    uv = base["proj_uv"]
    P = np.stack([(uv[:, 0] - cx) / fx * z, (uv[:, 1] - cy) / fy * z, z], 1).astype(np.float32)
    nrm = (P / np.linalg.norm(P, axis=1, keepdims=True)).astype(np.float32)
    d = np.linalg.norm(P, axis=1)
    g = synth.frustum_inputs(3, cam, 8, n_frames=n_frames)
    Rcw, tcw = g["Rcw"], (g["tcw"] * 0.05).astype(np.float32)
    Ow = np.stack([-(Rcw[f].T @ tcw[f]) for f in range(n_frames)]).astype(np.float32)
    return dict(world_pos=P, normal=nrm, min_dist=(0.5 * d).astype(np.float32), max_dist=(2.0 * d).astype(np.float32),
                Rcw=Rcw, tcw=tcw, Ow=Ow)


def upload_map(x, base):
    if "geometry" in base:
        g = base["geometry"]
        x.upload_map_geometry(g["world_pos"], g["normal"], g["min_dist"], g["max_dist"])
    x.upload_map(base["map_desc"])
    x.upload_map_graph(base["candidate"], base["observed"], base["bad"], base["edge_off"], base["edge_other"],
                       base["edge_ok"])


# ------------------------------------------------------------------------------------------------
def cpu_reference_pass(n_frames, seed0, threads, rows=MAP_ROWS):
    """The reference path restated on the CPU (oracle): LibTorch-CPU fp32 networks, single-threaded C
    post-processing and Matcher::ExtendMapMatches (window search + seed growing) of every frame against a table of
    `rows` map points in which every frame of the sample has its local map (same construction as the GPU arm).
    Building the synthetic map is not timed.  -> (seconds, frames)."""
    import torch
    from oracle import post_ref as O
    from oracle.net_ref import NetRef
    from ppg_slam_b200 import synth
    torch.set_num_threads(threads)
    cam = _cam()
    net = cpu_reference_pass.net if hasattr(cpu_reference_pass, "net") else NetRef()
    cpu_reference_pass.net = net
    frames = [synth.frame(seed0 + s, cam.width, cam.height) for s in range(n_frames)]
    t = 0.0
    t0 = time.perf_counter()
    recs = []
    for g in frames:
        m = net.forward_u8(g)
        recs.append(O.extract_post(cam, m["prob"], m["heat"], m["desc"]))
    t += time.perf_counter() - t0
    a = make_assoc_inputs(cam, recs, rows, n_slices=BATCH)
    t0 = time.perf_counter()
    for f, rec in enumerate(recs):
        n = rec["n_kp"]
        if n > 0:  # Matcher::ExtendMapMatches, whole function (window search + assignment + seed growing)
            O.extend_map_matches(cam, a["map_desc"], a["candidate"], a["observed"], a["bad"], a["edge_off"],
                                 a["edge_other"], a["edge_ok"], a["proj_all"][f], a["vcos_all"][f], a["tracked"],
                                 rec["kp_x"], rec["kp_y"], rec["desc"], np.full(n, -1, np.int32), rec["edge_start"],
                                 rec["edge_end"], rec["conn_off"], rec["conn_idx"], th=TH, ratio=RATIO)
    t += time.perf_counter() - t0
    return t, n_frames


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 4
    for _ in range(max(args.warmup, 1)):
        cpu_reference_pass(1, 1000, threads)
    t = 0.0
    for k in range(args.steps):
        dt, _ = cpu_reference_pass(sample, 2000 + k * sample, threads)
        t += dt
    fps = args.steps * sample / t
    line = {"metric": "frames/sec extract+associate at %dx%d" % (_cam().width, _cam().height), "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": CAMERA + " synthetic frames, extract + point-pair graph + association vs "
                                   "%d map points (CPU reference path: %d-frame sample per step)" % (MAP_ROWS, sample)},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": "%d frames per step x %d steps; networks torch-CPU fp32 with %d threads, "
                                       "post-processing and association single-threaded C (oracle/)" %
                                       (sample, args.steps, threads)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def _p_bytes(cam):
    return cam.width * cam.height * 4  # one fp32 dense map ("P" in SURVEY 8d)


def sharded_assoc_leg(args, rank, local_rank, world, dist, torch, cam, rec0):
    """BASELINE config 5: the map table (UMA-scale, --shard-rows rows) row-sharded over the ranks, the frame replicated;
    every rank scores its rows (tensor-core filter + exact re-score) and ONE ncclAllGather of the 20-byte per-row
    records, enqueued on the ctx stream behind the kernels (comm.cu), rebuilds the whole answer on every rank.
    Checked once against the un-sharded answer of rank 0."""
    from ppg_slam_b200 import capi, synth
    from ppg_slam_b200.sharded import shard_rows
    M = args.shard_rows
    # the frame everybody scores: rank 0's first record
    n = torch.tensor([rec0["n_kp"] if rank == 0 else 0], device="cuda", dtype=torch.int64)
    dist.broadcast(n, 0)
    N = int(n.item())
    buf = torch.zeros((N, 258), device="cuda", dtype=torch.float32)
    if rank == 0:
        buf.copy_(torch.from_numpy(np.concatenate([rec0["desc"], rec0["kp_x"][:, None], rec0["kp_y"][:, None]], 1)))
    dist.broadcast(buf, 0)
    h = buf.cpu().numpy()
    desc, kx, ky = np.ascontiguousarray(h[:, :256]), np.ascontiguousarray(h[:, 256]), np.ascontiguousarray(h[:, 257])
    inp = synth.association_inputs(23, desc, np.stack([kx, ky], 1), M, cam.width, cam.height, th=TH)
    shards = shard_rows(M, world)
    r0, k = shards[rank]
    per = max(kk for _, kk in shards)
    uid = torch.zeros(128, device="cuda", dtype=torch.uint8)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    x = capi.Extractor(cam, device=local_rank, max_batch=1, max_map_points=max(per, 1024))
    ones = np.ones(N, np.uint8)
    try:
        x.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
        x.upload_map(inp["map_desc"][r0:r0 + k])
        x.assoc_stage(kx, ky, desc, ones, inp["proj_uv"][r0:r0 + k], inp["view_cos"][r0:r0 + k], TH, RATIO)
        for _ in range(3):
            x.assoc_run()
            x.assoc_allgather(k, per)
        rec, _ = x.assoc_allgather_fetch()
        dist.barrier()
        torch.cuda.synchronize()
        iters = 20
        x.timer_start()
        for _ in range(iters):
            x.assoc_run()
            x.assoc_allgather(k, per)
        ms = x.timer_stop() / iters
        _, gather_us = x.assoc_allgather_fetch(records=False)
        t = torch.tensor([ms, gather_us], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, gather_us = float(t[0].item()), float(t[1].item())
        parity = None
        if rank == 0:
            full = capi.Extractor(cam, device=local_rank, max_batch=1, max_map_points=M)
            try:
                full.upload_map(inp["map_desc"])
                want = full.associate(kx, ky, desc, ones, inp["proj_uv"], inp["view_cos"], TH, RATIO)
            finally:
                full.close()
            got = np.concatenate([rec[r, :kk] for r, (_, kk) in enumerate(shards)], 0)
            parity = bool(np.array_equal(got[:, 0], want["best_idx"]) and np.array_equal(got[:, 1], want["second_idx"]) and
                          np.array_equal(got[:, 2], want["best_d"].view(np.int32)) and
                          np.array_equal(got[:, 3], want["second_d"].view(np.int32)) and
                          np.array_equal(got[:, 4].astype(np.uint8), want["accept"]))
        x.comm_destroy()
    finally:
        x.close()
    return {"rows": M, "rows_per_rank": per, "keypoints": N, "ranks": world,
            "ms_per_frame": ms, "gather_us": gather_us, "record_bytes_per_rank": per * 20,
            "what": "assoc_gemm + rescore of the rank's rows, pack, one ncclAllGather on the ctx stream (no host sync); "
                    "max over ranks, CUDA events", "parity_vs_unsharded": parity}


def pin_to_gpu_numa_node(index):
    """One rank per GPU: keep the rank's host thread (and with it the pinned buffers it allocates, first touch) on the
    cores NVML lists as local to that GPU, so that the frame / record DMA does not cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_b200(args, rank, local_rank, world):
    import torch
    from ppg_slam_b200 import capi
    dist = None
    numa_cpus = pin_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # NCCL announces its version on stdout at communicator creation; the contract is ONE JSON line on stdout,
        # so fd 1 points at stderr until the first collective has gone through
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch
    cam, frames = make_workload(B, seed0=rank * B)
    e = capi.Extractor(cam, device=local_rank, max_batch=B, max_map_points=max(args.map_rows, 1024))
    recs = e.run(frames)
    base = make_assoc_inputs(cam, recs, args.map_rows)
    if args.frustum:
        base["geometry"] = make_geometry(cam, base, B)
    geo = base.get("geometry")
    upload_map(e, base)
    core_only = args.assoc == "core"  # search core of ExtendMapMatches only (the round-1 step), for comparison
    empty = (np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros((0, 256), np.float32), np.zeros(0, np.uint8))

    # host buffers of the end-to-end arm live in pinned memory (what a capture ring buffer registered with
    # ppg_host_register looks like to the library): frames, per-frame projections
    pf = capi.pinned_array((B, cam.height, cam.width), np.uint8)
    for i, g in enumerate(frames):
        pf[i] = g
    proj_all = capi.pinned_array(base["proj_all"].shape, np.float32)
    vcos_all = capi.pinned_array(base["vcos_all"].shape, np.float32)
    proj_all[...] = base["proj_all"]
    vcos_all[...] = base["vcos_all"]

    def stage_proj(x, pinned=False):
        if geo is not None:  # Frame::CheckInFrustum on the device: 15 floats per frame cross the bus
            x.assoc_stage_poses(geo["Rcw"], geo["tcw"], geo["Ow"], args.map_rows, 0.5, TH, RATIO)
        elif pinned:
            x.assoc_stage_batch_async(proj_all, vcos_all, TH, RATIO)
        else:
            x.assoc_stage_batch(proj_all, vcos_all, TH, RATIO)

    def device_step(x=None):
        x = x or e
        x.run_device(B)
        if geo is not None:
            stage_proj(x)
        if core_only:
            x.assoc_run_batch(B)
        else:
            x.extend_run_batch(B)

    # ---- contexts: one per stream in flight.  A ctx is single-stream (like the reference's extractor object);
    # throughput callers keep several batches in flight on several ctxs of the same GPU, which also fills the SMs
    # that the latency-bound one-CTA-per-frame kernels (NMS, overlap filter, graph) leave idle.
    ctxs = [e]
    for _ in range(max(1, args.dev_streams, args.e2e_streams) - 1):
        x = capi.Extractor(cam, device=local_rank, max_batch=B, max_map_points=max(args.map_rows, 1024))
        upload_map(x, base)
        ctxs.append(x)
    all_ctxs = list(ctxs)
    dev = ctxs[:max(1, args.dev_streams)]

    # ---- device-timed arm: frames resident in HBM, K steps dealt round-robin to the device contexts
    for x in dev:
        x.upload(frames)
        stage_proj(x)
    clocks = ClockSampler(local_rank)
    t_w = time.perf_counter()
    k = 0
    while k < args.warmup or (time.perf_counter() - t_w < 1.5 and clocks.p is not None):
        for x in dev:
            device_step(x)  # warm-up; keeps the GPU under load until nvidia-smi delivers its first samples
        for x in dev:
            x.sync()
        k += 1
    barrier()
    l0 = sum(x.launch_count() for x in dev)
    for x in dev:
        x.timer_start()
    for i in range(args.steps):
        device_step(dev[i % len(dev)])
    ms = max(x.timer_stop() for x in dev)  # CUDA events on every stream; the longest bracket counts
    launches = sum(x.launch_count() for x in dev) - l0
    # Per-kernel durations for the rooflines: with more than one context in flight an event bracket around a kernel
    # also covers the other stream's kernels it waits for, so the stage events are taken in a single-stream pass of
    # the same step (same buffers, same clocks sampling window), right after the timed steps; median of 5 passes.
    e.set_profiling(True)
    passes = []
    for _ in range(6):
        device_step(e)
        e.sync()
        passes.append(e.stage_times())
    e.set_profiling(False)
    names = [k_ for k_, _ in passes[-1]]
    stage = [(nm, float(np.median([dict(p_)[nm] for p_ in passes[1:] if nm in dict(p_)]))) for nm in names]
    barrier()
    clk = clocks.stop()
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
    else:
        ms_max = ms
    fps = world * B * args.steps / (ms_max * 1e-3)

    # ---- end-to-end arm: host frames in, host records out, every step, through the pipelined calls.  ONE host thread
    # drives a ring of `--e2e-streams` contexts: enqueue batch i on ctx i % R (H2D of frames and projections from pinned
    # memory, networks, post-processing, ExtendMapMatches, D2H of records and results -- nothing waits), and collect
    # the batch that ctx ran R steps earlier.
    keep, fptrs, fstrides, _ = e._frame_ptrs([pf[i] for i in range(B)])
    n_dev_ctx = len(dev)
    ring = ctxs[:max(1, args.e2e_streams)]

    def enqueue(x):
        rc = x.lib.ppg_extract_async(x.h, fptrs, fstrides, B)
        if rc != 0:
            raise capi.PpgError(rc, x.lib.ppg_last_error(x.h).decode())
        stage_proj(x, pinned=True)
        if core_only:
            x.assoc_run_batch(B)
        else:
            x.extend_run_batch(B)
            x.extend_fetch_batch_async(B)

    def collect(x):
        x.extract_wait(B, allow_capacity=True, as_dicts=False)
        if core_only:
            return x.assoc_fetch_batch(B)
        return x.extend_collect(B, as_dicts=False)

    def e2e_run(steps):
        R = len(ring)
        for i in range(steps + R):
            x = ring[i % R]
            if i >= R:
                collect(x)
            if i < steps:
                enqueue(x)

    e2e_run(max(len(ring), args.warmup))
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    t_e2e = time.perf_counter() - t0
    barrier()
    xres = None if core_only else e.extend_collect(B)  # e = ring[0]: the results of its last batch, as dicts
    rec_bytes = e.record_bytes()
    for x in all_ctxs[1:]:
        x.close()
    if dist is not None:
        t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    fps_e2e = world * B * args.steps / t_e2e
    recs = [capi._frame_to_dict(e._outs[i]) for i in range(B)]
    h2d = B * cam.width * cam.height + (B * 64 if geo is not None else B * args.map_rows * 12)
    if core_only:
        d2h_assoc = B * args.map_rows * 17
    else:  # F.mvpMapPoints, F.mvpMapEdges, tracked flags and the counters of every frame
        d2h_assoc = B * (1024 * 4 + e.cfg.max_edges * 4 + args.map_rows + 32)
    d2h = int(B * rec_bytes + d2h_assoc)  # whole records (descriptor area included) in one strided copy

    # ---- batch-1 latency (p50 ms/frame), host in -> host out, synchronous calls
    lat = []
    for k in range(12):
        t0 = time.perf_counter()
        e.lib.ppg_extract(e.h, fptrs, fstrides, 1, e._outs)
        if core_only:
            e.assoc_stage(*empty, base["proj_all"][0], base["vcos_all"][0], TH, RATIO)
            e.assoc_run_frame(0)
            e.assoc_fetch()
        else:
            e.assoc_stage_batch(base["proj_all"][:1], base["vcos_all"][:1], TH, RATIO)
            e.extend_run_batch(1)
            e.extend_fetch_batch(1)
        lat.append((time.perf_counter() - t0) * 1e3)
    p50 = float(np.median(lat[2:]))

    # ---- association GEMM at the UMA-VI scale of BASELINE config 4 (1000 keypoints x 50 000 map rows), N = 1 only
    assoc_roof = None
    if world == 1 and not args.no_assoc_gemm:
        assoc_roof = assoc_gemm_roofline(cam, local_rank)

    sharded = None
    if dist is not None and not args.no_sharded:
        sharded = sharded_assoc_leg(args, rank, local_rank, world, dist, torch, cam, recs[0])

    if rank == 0:
        tf_peak, hbm_peak, how = _peaks()
        sd = dict(stage)
        conv1b_ms = sd.get("conv1b")
        roof = None
        scale = cam.width * cam.height / (752.0 * 480.0)  # conv FLOPs scale with the pixel count
        if conv1b_ms:
            fused = "conv1a" not in sd  # conv1a computed by conv1b's producer warps: no separate stage
            # GFLOP / ms = TFLOP/s; the fused kernel also does conv1a's 0.42 GFLOP per frame (0.21 GMAC)
            ach = (CONV1B_GFLOP_PER_FRAME + (CONV1A_GFLOP_PER_FRAME if fused else 0.0)) * scale * B / conv1b_ms
            traffic = None
            tp = os.path.join(ROOT, "profiles", "conv1b_traffic.json")
            if os.path.exists(tp):
                traffic = json.load(open(tp)).get("fused_dram_bytes_per_launch" if fused else "dram_bytes_per_launch")
            kname = ("conv_t64_fused_kernel[conv1a 1->64 3x3 + ReLU computed in the producer warps, conv1b 64->64 3x3 "
                     "@%dx%d + ReLU + 2x2 pool; algorithmic FLOPs of both layers]" if fused else
                     "conv_t64_kernel[conv1b 64->64 3x3 @%dx%d + ReLU + 2x2 pool]") % (cam.width, cam.height)
            roof = {"bound": "tensor",
                    "kernel": kname,
                    "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": traffic,
                    "peak_source": how + " (sustained bf16 cuBLAS; fp16 runs on the same pipe)",
                    "ms_per_launch": conv1b_ms, "frames_per_launch": B,
                    "timing": "CUDA events around the launch in single-stream passes of the same step (median of 5)"}
        conv_ms = sum(v for k, v in stage if k.startswith("conv") and k != "conv1a" or k.startswith("edge0")
                      or k.startswith("edge1"))  # tensor-core layers
        # streaming post-processing kernels against the measured HBM copy bandwidth; algorithmic bytes per frame as in
        # SURVEY 8d (P = one fp32 H x W map): scan reads P; refine reads + writes P; remap reads P and the 2P map table,
        # writes P; the descriptor sampler reads 4 texels x 1 KB and writes 1 KB per keypoint
        P = _p_bytes(cam)
        nkp = float(np.mean([r["n_kp"] for r in recs]))
        more = {}

        def hbm(name, stage_name, bytes_per_frame):
            t_ms = sd.get(stage_name)
            if t_ms:
                gbs = bytes_per_frame * B / (t_ms * 1e-3) / 1e9
                more[name] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                              "ms_per_launch": t_ms, "algorithmic_bytes_per_frame": int(bytes_per_frame)}
        hbm("scan_kernel", "post.scan", P)
        hbm("refine_kernel", "post.refine", 2 * P)
        hbm("remap_kernel", "post.remap", 4 * P)
        if sd.get("post.refine") and sd.get("post.remap"):
            t_ms = sd["post.refine"] + sd["post.remap"]
            gbs = 6 * P * B / (t_ms * 1e-3) / 1e9
            more["refine+remap"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                    "frac": gbs / hbm_peak, "ms_per_launch": t_ms, "algorithmic_bytes_per_frame": 6 * P}
        hbm("desc_kernel", "post.descriptors", nkp * 5 * 1024)
        hbm("conv1a_tc_kernel", "conv1a", cam.width * cam.height * (1 + 128))
        # convDb (1x1 256 -> 256): reads the fp16 map, writes the fp32 dense descriptors; 0.37 GMAC per frame is nothing
        hbm("conv_t128_kernel[convDb]", "convDb", (cam.width // 8) * (cam.height // 8) * 256 * (2 + 4))
        if conv_ms:
            tfl = (GFLOP_PER_FRAME - 0.84 + (CONV1A_GFLOP_PER_FRAME if "conv1a" not in sd else 0.0)) * scale * B / conv_ms
            more["all_tensor_core_convolutions"] = {"bound": "tensor", "achieved": tfl, "peak": tf_peak,
                                                    "unit": "TFLOP/s", "frac": tfl / tf_peak, "ms_per_step": conv_ms}
        if assoc_roof:
            more["assoc_gemm_kernel"] = assoc_roof
        serial = {k_: round(sd[k_], 4) for k_ in ("post.nms+topk", "post.lines_filter", "post.lines_graph",
                                                    "extend.walk") if k_ in sd}
        cpu_t, cpu_n = 0.0, 0
        cpu_threads = os.cpu_count() or 1
        if world == 1:  # the CPU baseline is taken at N = 1 only (the other ranks would idle behind it)
            cpu_reference_pass(1, 1000, cpu_threads)
            while cpu_t < 8.0 and cpu_n < 40:
                dt, n = cpu_reference_pass(4, 3000 + cpu_n, cpu_threads)
                cpu_t += dt
                cpu_n += n
        assoc_stats = None
        if xres is not None:
            assoc_stats = {"keypoints_per_frame": nkp,
                           "accepted_per_frame": float(np.mean([x_["n_accepted"] for x_ in xres])),
                           "grown_along_edges_per_frame": float(np.mean([x_["n_grown"] for x_ in xres])),
                           "matched_keypoints_per_frame": float(np.mean([(x_["kp_mp"] >= 0).sum() for x_ in xres])),
                           "min_accepted_in_batch": int(min(x_["n_accepted"] for x_ in xres)),
                           "window_rescans_per_frame": float(np.mean([x_["n_rescans"] for x_ in xres]))}
        line = {"metric": "frames/sec extract+associate at %dx%d" % (_cam().width, _cam().height), "value": fps, "unit": "frames/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
                "data": "synthetic",
                "config": {"workload": "%s %dx%d batch-%d synthetic frames per GPU: extract + point-pair graph "
                                       "+ %s of every frame vs %d resident map points (every frame has its local map "
                                       "of %d points in the table)" %
                                       (CAMERA, cam.width, cam.height, B,
                                        "association search core" if core_only else
                                        "Matcher::ExtendMapMatches (window search + assignment + seed growing)",
                                        args.map_rows, args.map_rows // B),
                           "batch_per_gpu": B, "map_rows": args.map_rows, "sharding": "frames (no collective)",
                           "projections": "Frame::CheckInFrustum on the device" if geo is not None else "staged by the host",
                           "device_contexts_in_flight": n_dev_ctx, "e2e_contexts_in_flight": len(ring),
                           "e2e_host_threads": 1, "rank_cpu_affinity": numa_cpus, "host_cores": os.cpu_count(),
                           "l2": "per-step working set ~3.4 GB of activations streams through the 126 MB L2 "
                                 "(inputs larger than L2; no explicit flush)"},
                "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": d2h},
                "latency": {"p50_ms_per_frame_batch1": p50},
                "association": assoc_stats,
                "gpu_launches": int(launches),
                "clocks": clk,
                "roofline": roof,
                "roofline_more": more,
                "serial_kernels_ms_per_step": serial,
                "stages_ms_per_step": {k: round(v, 4) for k, v in stage},
                "sharded_assoc": sharded,
                "cpu_baseline": ({"value": cpu_n / cpu_t, "unit": "frames/s", "cores": cpu_threads, "kind": "port",
                                  "sample": "%d frames (same synthetic workload, every frame with its local map); "
                                            "networks torch-CPU fp32 on %d threads, post-processing + association "
                                            "single-threaded C oracle" % (cpu_n, cpu_threads)} if cpu_n else None)}
        print(json.dumps(line), flush=True)
    e.close()
    capi.drop_pinned()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def assoc_gemm_roofline(cam, device):
    """BASELINE config 4: ~1000 keypoints against 50 000 map descriptors through the brute-force bf16 tensor-core GEMM
    with the fused window test + per-row top-4 (assoc_gemm_kernel) and the exact re-score.  FLOPs = 2 * M * N * 256."""
    from ppg_slam_b200 import capi, synth
    tf_peak, _, how = _peaks()
    M, N = 50000, 1000
    rs = np.random.RandomState(3)
    kx = rs.uniform(8, cam.width - 8, N).astype(np.float32)
    ky = rs.uniform(8, cam.height - 8, N).astype(np.float32)
    fd = rs.normal(size=(N, 256)).astype(np.float32)
    fd /= np.linalg.norm(fd, axis=1, keepdims=True)
    inp = synth.association_inputs(4, fd, np.stack([kx, ky], 1), M, cam.width, cam.height, th=TH)
    x = capi.Extractor(cam, device=device, max_batch=1, max_map_points=M, junction_max_num=1024)
    try:
        x.upload_map(inp["map_desc"])
        x.assoc_stage(kx, ky, fd, np.ones(N, np.uint8), inp["proj_uv"], inp["view_cos"], TH, RATIO)
        for _ in range(3):
            x.assoc_run()
        x.sync()
        x.set_profiling(True)
        ts = []
        for _ in range(5):
            x.assoc_run()
            x.sync()
            ts.append(dict(x.stage_times()))
        x.set_profiling(False)
        g = float(np.median([t["assoc.gemm(top4)"] for t in ts]))
        r = float(np.median([t["assoc.rescore"] for t in ts]))
    finally:
        x.close()
    tfl = 2.0 * M * N * 256 / (g * 1e-3) / 1e12
    return {"bound": "tensor", "achieved": tfl, "peak": tf_peak, "unit": "TFLOP/s", "frac": tfl / tf_peak,
            "ms_per_launch": g, "rescore_ms": r, "rows": M, "keypoints": N,
            "note": "brute-force GEMM + fused window test / top-4; the benchmark step uses the windowed exact-distance "
                    "lists kernel instead (DESIGN.md 3.3)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--map-rows", type=int, default=MAP_ROWS)
    ap.add_argument("--dev-streams", type=int, default=4, help="contexts in flight in the device-timed arm")
    ap.add_argument("--e2e-streams", type=int, default=6,
                    help="contexts in flight in the e2e arm (one host thread drives them all)")
    ap.add_argument("--assoc", default="extend", choices=["extend", "core"],
                    help="extend: the whole Matcher::ExtendMapMatches on the GPU (default); core: its search core only")
    ap.add_argument("--frustum", action="store_true",
                    help="project the map points on the device (Frame::CheckInFrustum) instead of staging projections")
    ap.add_argument("--shard-rows", type=int, default=50000,
                    help="N > 1: rows of the map table of the row-sharded association leg (BASELINE config 5)")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the row-sharded association leg")
    ap.add_argument("--no-assoc-gemm", action="store_true", help="N = 1: skip the 1000 x 50 000 association GEMM timing")
    ap.add_argument("--camera", default="EuRoC", choices=["EuRoC", "TUM-VI", "TUM-VI-1024", "UMA-VI"],
                    help="frame shape / calibration; EuRoC 752x480 is the benchmark configuration")
    args = ap.parse_args()
    global CAMERA
    CAMERA = args.camera
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
