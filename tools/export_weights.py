#!/usr/bin/env python
"""Export the reference's four TorchScript nets to one flat fp32 blob.

Source of the weights: /root/reference/net/{Backbone,PointHeatmap,EdgeHeatmap,Descriptor}.pt
(loaded by the reference at feature/src/PPGExtractor.cpp:77-80).  The blob holds the RAW
parameters (BatchNorm statistics are NOT folded here; the product library folds them at load
time, the oracle applies them the way torch.batch_norm does).

Blob layout (little endian):
    char[8]  magic   "PPGW0001"
    u32      n_tensors
    n_tensors x { char[48] name; u32 ndim; u32 dims[4]; u64 offset_bytes (from file start) }
    raw fp32 data, each tensor 64-byte aligned, C-contiguous (OIHW for conv weights)

Run in the build container (needs /root/reference):  python tools/export_weights.py
"""
import os, struct, sys
import numpy as np
import torch

REF = os.environ.get("PPG_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ppg_slam_b200", "weights", "ppg_weights.bin")

NETS = [("backbone", "Backbone.pt"), ("junction", "PointHeatmap.pt"),
        ("edge", "EdgeHeatmap.pt"), ("descriptor", "Descriptor.pt")]


def main():
    tensors = []
    for prefix, fn in NETS:
        m = torch.jit.load(os.path.join(REF, "net", fn), map_location="cpu")
        for k, v in m.state_dict().items():
            if k.endswith("num_batches_tracked"):
                continue
            tensors.append((prefix + "." + k, v.detach().to(torch.float32).contiguous().numpy()))
    n = len(tensors)
    header_bytes = 8 + 4 + n * (48 + 4 + 16 + 8)
    off = (header_bytes + 63) // 64 * 64
    table = []
    for name, a in tensors:
        table.append((name, a, off))
        off = (off + a.nbytes + 63) // 64 * 64
    buf = bytearray(off)
    buf[0:8] = b"PPGW0001"
    struct.pack_into("<I", buf, 8, n)
    p = 12
    for name, a, o in table:
        nb = name.encode()
        assert len(nb) < 48
        buf[p:p + len(nb)] = nb
        dims = list(a.shape) + [1] * (4 - a.ndim)
        struct.pack_into("<I4IQ", buf, p + 48, a.ndim, *dims, o)
        p += 48 + 4 + 16 + 8
        buf[o:o + a.nbytes] = a.tobytes()
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "wb") as f:
        f.write(buf)
    print("wrote", os.path.normpath(OUT), len(buf), "bytes,", n, "tensors,",
          sum(a.size for _, a, _ in table), "params")


if __name__ == "__main__":
    sys.exit(main())
