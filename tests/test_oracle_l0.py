"""Pins the L0 oracle (oracle/net_ref.py: the four networks restated in plain torch fp32 from the exported
weight blob) against golden outputs of the reference's OWN TorchScript files (tests/golden/make_golden.py
ran net/*.pt through torch.jit; PPGExtractor.cpp:77-92,152-155)."""
import os

import numpy as np
import pytest

from oracle.net_ref import NetRef
from ppg_slam_b200 import synth


@pytest.fixture(scope="module")
def net():
    return NetRef()


def test_small_frame_matches_torchscript(net, golden_dir):
    g = np.load(os.path.join(golden_dir, "l0_small.npz"))
    np.testing.assert_array_equal(synth.frame(7, 96, 64, n_rect=6, n_line=5), g["gray"])  # synth is reproducible
    r = net.forward_u8(g["gray"])
    # same weights, same ops, same library: only thread-partitioning noise is allowed
    assert np.abs(r["prob"] - g["prob"]).max() < 2e-6
    assert np.abs(r["heat"] - g["heat"]).max() < 2e-5
    assert np.abs(r["desc"] - g["desc"]).max() < 2e-3 * max(1.0, np.abs(g["desc"]).max() / 100)


@pytest.mark.parametrize("name,W,H,seed", [("euroc", 752, 480, 0), ("tumvi", 512, 512, 1)])
def test_full_frame_samples(net, golden_dir, name, W, H, seed):
    g = np.load(os.path.join(golden_dir, "l0_samples.npz"))
    r = net.forward_u8(synth.frame(seed, W, H))
    pos, dpos = g[name + "_pos"], g[name + "_dpos"]
    assert np.abs(r["prob"].ravel()[pos] - g[name + "_prob"]).max() < 5e-6
    assert np.abs(r["heat"].ravel()[pos] - g[name + "_heat"]).max() < 5e-5
    d = r["desc"].ravel()[dpos]
    assert np.abs(d - g[name + "_desc"]).max() < 5e-3
    m = g[name + "_moments"]
    assert abs(r["prob"].astype(np.float64).sum() - m[0]) < 1e-3 * max(1.0, m[0])
    assert abs(r["heat"].astype(np.float64).sum() - m[2]) < 1e-3 * max(1.0, m[2])
    n = g[name + "_n_ge_thr"]
    assert abs(int((r["prob"] >= 1.0 / 128).sum()) - int(n[0])) <= 3
    assert abs(int((r["heat"] > 0.2).sum()) - int(n[1])) <= 30


def test_pixel_shuffle_mapping(net):
    """P[8h+i][8w+j] = softmax[c=8i+j][h][w]  (SURVEY 8a a4)."""
    import torch
    import torch.nn.functional as F
    t = torch.arange(65 * 2 * 3, dtype=torch.float32).reshape(1, 65, 2, 3)
    p = F.pixel_shuffle(t.narrow(1, 0, 64), 8)[0, 0]
    for (h, w, i, j) in [(0, 0, 0, 0), (1, 2, 3, 5), (0, 1, 7, 7)]:
        assert p[8 * h + i, 8 * w + j] == t[0, 8 * i + j, h, w]
