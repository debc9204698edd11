"""Converts the reference's DBoW3 vocabulary (Vocabulary/voc_*_9x3.gz, binary + QuickLZ despite the suffix) into the
flat POD blob the library loads (ppg_slam_b200/weights/voc_<name>.bin), like tools/export_weights.py does for net/*.pt.
Layout: int32 magic 'PVOC', k, L, scoring, weighting, n_nodes, dim | children int32 [n_nodes][k] | word_id int32
[n_nodes] | weight float64 [n_nodes] | desc float32 [n_nodes][dim].
python tools/export_vocabulary.py [/root/reference/Vocabulary/voc_euroc_9x3.gz ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppg_slam_b200 import vocabulary  # noqa: E402


def main():
    srcs = sys.argv[1:] or ["/root/reference/Vocabulary/voc_euroc_9x3.gz", "/root/reference/Vocabulary/voc_tum_9x3.gz"]
    for src in srcs:
        v = vocabulary.Vocabulary(src)
        name = os.path.basename(src).split(".")[0]
        out = os.path.join(ROOT, "ppg_slam_b200", "weights", name + ".bin")
        vocabulary.save_blob(v, out)
        print(out, v.k, v.L, v.scoring, v.weighting, v.n_nodes, v.n_words)


if __name__ == "__main__":
    main()
