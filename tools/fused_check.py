"""Where does the fused conv1a + conv1b kernel differ from the two separate kernels?  (GPU box; validation aid.)

Runs the same frames with PPG_CONV_KERNEL=5 (fused, default) and =3 (separate kernels), fetches conv1b's raw fp16 output
and prints how many values differ, by how many fp16 ulps, and where (tile coordinates of the 38 x 4 output tiles)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppg_slam_b200 import cameras, capi, synth  # noqa: E402


def run(mode, frames):
    os.environ["PPG_CONV_KERNEL"] = mode
    e = capi.Extractor(cameras.EUROC, max_batch=len(frames))
    try:
        e.run(frames, allow_capacity=True)
        return [e.layer_output("conv1b", i).view(np.float16).reshape(240, 376, 64).copy() for i in range(len(frames))]
    finally:
        e.close()


def main():
    frames = [synth.frame(s, 752, 480) for s in (4, 5, 6)]
    a, b = run("5", frames), run("3", frames)
    for i in range(len(frames)):
        x, y = a[i], b[i]
        d = x.view(np.int16).astype(np.int32) - y.view(np.int16).astype(np.int32)
        nz = np.argwhere(d != 0)
        print("frame %d: %d of %d values differ (%.4f %%), max ulp %d, max abs %.4g, ref max %.4g" % (
            i, len(nz), d.size, 100.0 * len(nz) / d.size, np.abs(d).max() if len(nz) else 0,
            np.abs(x.astype(np.float32) - y.astype(np.float32)).max(), np.abs(y.astype(np.float32)).max()))
        if len(nz):
            yy, xx, cc = nz[:, 0], nz[:, 1], nz[:, 2]
            print("   pooled rows: hist of (y %% 2) %s, of (x %% 19) %s" % (np.bincount(yy % 2, minlength=2),
                                                                          np.bincount(xx % 19, minlength=19)))
            print("   channels: %s" % np.bincount(cc, minlength=64))
            print("   first: %s" % nz[:8].tolist())
            print("   values fused %s separate %s" % (x[tuple(nz[0])], y[tuple(nz[0])]))


if __name__ == "__main__":
    main()
