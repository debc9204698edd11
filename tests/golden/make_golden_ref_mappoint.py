"""Golden vectors of MapPoint::ComputeDistinctiveDescriptors (feature/src/MapPoint.cpp:234-302) made by the REFERENCE's
own C++ (oracle/ref_build.py compiles feature/src/MapPoint.cpp from /root/reference): a real MapPoint observed by raw key
frames, the real function, the mDescriptor it chose.  Run in the build container:
python tests/golden/make_golden_ref_mappoint.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as R  # noqa: E402


def cases():
    rs = np.random.RandomState(7)
    for k, n in enumerate([1, 2, 3, 4, 6, 9, 14, 27, 40]):
        d = rs.randn(256) + rs.randn(n, 256) * [0.05, 0.3, 1.0][k % 3]
        if k % 4 == 3:
            d[1] = d[0]  # duplicate observations: equal medians, the first one wins (strict <)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        yield "mp%d" % k, d.astype(np.float32)


def main():
    out = {}
    for name, d in cases():
        ref = R.distinctive_descriptor(d)
        out[name + "/obs_desc"] = d
        out[name + "/ref_descriptor"] = ref
        row = [i for i in range(len(d)) if np.array_equal(ref.view(np.uint32), d[i].view(np.uint32))]
        print(name, "observations", len(d), "chosen row(s)", row)
    path = os.path.join(ROOT, "tests", "golden", "ref_l2_mappoint.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
