#!/usr/bin/env python
"""Generates the committed golden fixtures from the REAL reference artefacts.  Run in the build
container only (needs /root/reference and cv2); the fixtures travel, this script's inputs do not.

  l0_small.npz      dense outputs of the reference's own TorchScript nets (net/*.pt through torch.jit,
                    fp32 CPU -- the same graphs LibTorch executes at feature/src/PPGExtractor.cpp:152-155)
                    on a 96x64 synthetic frame: junction prob map, heat score map, dense descriptors.
  l0_samples.npz    the same for 752x480 / 512x512 frames, sampled at 4096 fixed positions + moments.
  cv_kat.npz        cv2 4.13 outputs of the four OpenCV calls the extractor makes (undistortPoints,
                    fisheye.undistortPoints, initUndistortRectifyMap, remap) on fixed inputs.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ppg_slam_b200 import cameras, synth  # noqa: E402

REF = os.environ.get("PPG_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def jit_forward(nets, gray):
    x = torch.from_numpy(gray)[None, None].to(torch.float32) / 255.0  # PPGExtractor.cpp:151
    with torch.no_grad():
        f = nets["Backbone"](x)
        j = nets["PointHeatmap"](f)
        h = nets["EdgeHeatmap"](f)
        d = nets["Descriptor"](f)
        prob = F.pixel_shuffle(torch.softmax(j, 1).narrow(1, 0, 64), 8)[0, 0]  # :161-162
        heat = torch.softmax(h, 1).select(1, 1)[0]  # :242
    return prob.contiguous().numpy(), heat.contiguous().numpy(), d[0].contiguous().numpy()


def main():
    torch.set_num_threads(8)
    nets = {n: torch.jit.load(os.path.join(REF, "net", n + ".pt"), map_location="cpu").eval()
            for n in ("Backbone", "PointHeatmap", "EdgeHeatmap", "Descriptor")}
    g = synth.frame(7, 96, 64, n_rect=6, n_line=5)
    prob, heat, desc = jit_forward(nets, g)
    np.savez_compressed(os.path.join(OUT, "l0_small.npz"), gray=g, prob=prob, heat=heat,
                        desc=desc.astype(np.float32))
    samples = {}
    for name, (W, H, seed) in {"euroc": (752, 480, 0), "tumvi": (512, 512, 1)}.items():
        g = synth.frame(seed, W, H)
        prob, heat, desc = jit_forward(nets, g)
        rs = np.random.RandomState(1234)
        pos = rs.randint(0, H * W, 4096)
        dpos = rs.randint(0, desc.size, 4096)
        samples[name + "_pos"] = pos
        samples[name + "_dpos"] = dpos
        samples[name + "_prob"] = prob.ravel()[pos]
        samples[name + "_heat"] = heat.ravel()[pos]
        samples[name + "_desc"] = desc.ravel()[dpos]
        samples[name + "_moments"] = np.array([prob.astype(np.float64).sum(), (prob.astype(np.float64) ** 2).sum(),
                                               heat.astype(np.float64).sum(), (heat.astype(np.float64) ** 2).sum(),
                                               np.abs(desc.astype(np.float64)).sum()])
        samples[name + "_n_ge_thr"] = np.array([(prob >= 1.0 / 128).sum(), (heat > 0.2).sum()])
    np.savez_compressed(os.path.join(OUT, "l0_samples.npz"), **samples)

    import cv2
    kat = {}
    rs = np.random.RandomState(99)
    for cam in (cameras.EUROC, cameras.TUMVI, cameras.UMA):
        K = np.array(cam.K, np.float32).reshape(3, 3)
        D = np.array(cam.D, np.float32).reshape(4, 1)
        pts = np.stack([rs.randint(0, cam.width, 2000), rs.randint(0, cam.height, 2000)], 1).astype(np.float32)
        corners = np.array([[0, 0], [cam.width, 0], [0, cam.height], [cam.width, cam.height]], np.float32)
        pts = np.concatenate([pts, corners])
        if cam.fisheye:
            und = cv2.fisheye.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, None, K).reshape(-1, 2)
        else:
            und = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, None, K).reshape(-1, 2)
        kat[cam.name + "_pts"] = pts
        kat[cam.name + "_und"] = und.astype(np.float32)
    cam = cameras.EUROC
    K = np.array(cam.K, np.float32).reshape(3, 3)
    D = np.array(cam.D, np.float32).reshape(4, 1)
    mx, my = cv2.initUndistortRectifyMap(K, D, np.eye(3), K, (cam.width, cam.height), cv2.CV_32F)
    rows = np.array([0, 1, 17, 239, 240, 400, 478, 479])
    kat["euroc_map_rows"] = rows
    kat["euroc_mx_rows"] = mx[rows]
    kat["euroc_my_rows"] = my[rows]
    kat["euroc_map_sum"] = np.array([mx.astype(np.float64).sum(), my.astype(np.float64).sum()])
    vv, uu = np.mgrid[0:cam.height, 0:cam.width]
    src = (((uu * 7 + vv * 13) % 251).astype(np.float32) / np.float32(250))  # the test rebuilds this formula
    dst = cv2.remap(src, mx, my, cv2.INTER_LINEAR)
    kat["remap_dst_rows"] = dst[rows]
    kat["remap_dst_sum"] = np.array([dst.astype(np.float64).sum()])
    # fisheye map (unused by the shipped configs, kept for completeness)
    camf = cameras.TUMVI
    Kf = np.array(camf.K, np.float32).reshape(3, 3)
    Df = np.array(camf.D, np.float32).reshape(4, 1)
    fx_, fy_ = cv2.fisheye.initUndistortRectifyMap(Kf, Df, np.eye(3), Kf, (camf.width, camf.height), cv2.CV_32F)
    kat["tumvi_mx_rows"] = fx_[rows]
    kat["tumvi_my_rows"] = fy_[rows]
    np.savez_compressed(os.path.join(OUT, "cv_kat.npz"), **kat)
    for f in ("l0_small.npz", "l0_samples.npz", "cv_kat.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
