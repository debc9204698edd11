"""Reader for the flat weight blob written by tools/export_weights.py."""
import os, struct
import numpy as np

DEFAULT_BLOB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "weights", "ppg_weights.bin")


def load_blob(path=DEFAULT_BLOB):
    """-> dict name -> np.float32 array (views into one buffer)."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != b"PPGW0001":
        raise ValueError("bad weight blob magic in %s" % path)
    (n,) = struct.unpack_from("<I", buf, 8)
    out = {}
    p = 12
    for _ in range(n):
        name = buf[p:p + 48].split(b"\0", 1)[0].decode()
        ndim, d0, d1, d2, d3, off = struct.unpack_from("<I4IQ", buf, p + 48)
        p += 48 + 4 + 16 + 8
        shape = (d0, d1, d2, d3)[:ndim]
        cnt = int(np.prod(shape)) if ndim else 1
        out[name] = np.frombuffer(buf, dtype="<f4", count=cnt, offset=off).reshape(shape)
    return out
