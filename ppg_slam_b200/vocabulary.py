"""Reader of the reference's bag-of-words vocabularies (Vocabulary/voc_*_9x3.gz) and the POD form the GPU transform
uses.  Despite the suffix the files are DBoW3's own binary format (DBoW3 is a third-party dependency of the reference,
find_package(DBoW3) in CMakeLists.txt:18, unpinned and not vendored; the layout below restates
DBoW3::Vocabulary::toStream / fromStream):

    uint64 signature 88877711233 | bool compressed | uint32 n_nodes | [compressed: uint32 n_chunks, then QuickLZ 1.5
    level-1 packets of 10000 bytes each] | int32 k, L, scoring, weighting | (n_nodes - 1) x {uint32 id, uint32 parent,
    double weight, descriptor} in the order a depth-first walk writes them | uint32 n_words | n_words x {uint32 word id,
    uint32 node id}.   descriptor (DescManip::toStream) = int32 cols, rows, type + raw rows (CV_32F, 1 x 256 here).
"""
import struct

import numpy as np

SIGNATURE = 88877711233


def _qlz_decompress(src):
    """One QuickLZ 1.5.0 packet, compression level 1, no streaming buffer -> bytes."""
    flags = src[0]
    if flags & 2:
        csize, dsize = struct.unpack_from("<II", src, 1)
        hl = 9
    else:
        csize, dsize = src[1], src[2]
        hl = 3
    if not flags & 1:
        return bytes(src[hl:hl + dsize]), csize
    if (flags >> 2) & 3 != 1:
        raise ValueError("QuickLZ packet of compression level %d (only level 1 is handled)" % ((flags >> 2) & 3))
    dst = bytearray(dsize)
    table = [0] * 4096
    bitlut = (4, 0, 1, 0, 2, 0, 1, 0, 3, 0, 1, 0, 2, 0, 1, 0)
    s, d = hl, 0
    last = dsize - 1
    last_matchstart = last - 6 - 4
    last_hashed = -1
    cword = 1
    pad = bytes(src) + b"\0" * 8

    def rd(buf, i, n):
        return int.from_bytes(buf[i:i + n], "little")

    def hash_upto(lh, mx):
        while lh < mx:
            lh += 1
            v = dst[lh] | (dst[lh + 1] << 8) | (dst[lh + 2] << 16)
            table[((v >> 12) ^ v) & 0xfff] = lh
        return lh

    while True:
        if cword == 1:
            cword = rd(pad, s, 4)
            s += 4
        fetch = rd(pad, s, 4)
        if cword & 1:
            cword >>= 1
            off = table[(fetch >> 4) & 0xfff]
            if fetch & 0xf:
                mlen = (fetch & 0xf) + 2
                s += 2
            else:
                mlen = pad[s + 2]
                s += 3
            for i in range(mlen):  # byte by byte: the match may overlap its own output
                dst[d + i] = dst[off + i]
            d += mlen
            hash_upto(last_hashed, d - mlen)
            last_hashed = d - 1
        elif d < last_matchstart:
            n = bitlut[cword & 0xf]
            dst[d:d + 4] = pad[s:s + 4]
            cword >>= n
            d += n
            s += n
            last_hashed = hash_upto(last_hashed, d - 3)
        else:
            while d <= last:
                if cword == 1:
                    s += 4
                    cword = 1 << 31
                dst[d] = pad[s]
                d += 1
                s += 1
                cword >>= 1
            return bytes(dst), csize


class Vocabulary:
    """k-ary tree of DBoW3::Vocabulary: nodes[i] = (parent, weight, descriptor); children in file order (the order
    DBoW3's loader appends them, which is the order transform() visits them)."""

    def __init__(self, path):
        raw = open(path, "rb").read()
        sig, = struct.unpack_from("<Q", raw, 0)
        if sig != SIGNATURE:
            raise ValueError("%s: not a DBoW3 binary vocabulary" % path)
        compressed = raw[8] != 0
        n_nodes, = struct.unpack_from("<I", raw, 9)
        if compressed:
            n_chunks, = struct.unpack_from("<I", raw, 13)
            off, parts = 17, []
            for _ in range(n_chunks):
                data, used = _qlz_decompress(memoryview(raw)[off:])
                parts.append(data)
                off += used
            buf = b"".join(parts)
        else:
            buf = raw[13:]
        self.k, self.L, self.scoring, self.weighting = struct.unpack_from("<iiii", buf, 0)
        o = 16
        self.n_nodes = n_nodes
        self.parent = np.zeros(n_nodes, np.int32)
        self.weight = np.zeros(n_nodes, np.float64)
        self.word_id = np.full(n_nodes, -1, np.int32)
        self.children = [[] for _ in range(n_nodes)]
        desc = None
        for _ in range(n_nodes - 1):
            nid, pid, w = struct.unpack_from("<IId", buf, o)
            o += 16
            cols, rows, typ = struct.unpack_from("<iii", buf, o)
            o += 12
            if typ != 5 or rows != 1:  # CV_32F, one row
                raise ValueError("unexpected descriptor type %d (%d x %d)" % (typ, rows, cols))
            if desc is None:
                desc = np.zeros((n_nodes, cols), np.float32)
            desc[nid] = np.frombuffer(buf, np.float32, cols, o)
            o += 4 * cols
            self.parent[nid] = pid
            self.weight[nid] = w
            self.children[pid].append(nid)
        self.desc = desc
        n_words, = struct.unpack_from("<I", buf, o)
        o += 4
        self.n_words = n_words
        self.word_node = np.zeros(n_words, np.int32)
        for _ in range(n_words):
            wid, nid = struct.unpack_from("<II", buf, o)
            o += 8
            self.word_id[nid] = wid
            self.word_node[wid] = nid
        if o != len(buf):
            raise ValueError("%d trailing bytes in the vocabulary" % (len(buf) - o))

    def child_table(self):
        """-> (n_nodes, k) int32 children in visiting order, -1 padded (leaves: all -1)."""
        t = np.full((self.n_nodes, self.k), -1, np.int32)
        for i, ch in enumerate(self.children):
            if len(ch) > self.k:
                raise ValueError("node %d has %d children (k = %d)" % (i, len(ch), self.k))
            t[i, :len(ch)] = ch
        return t


MAGIC = 0x434F5650  # 'PVOC'


class PodVocabulary:
    """The flat form (tools/export_vocabulary.py) -- what ppg_upload_vocabulary takes."""

    def __init__(self, k, L, scoring, weighting, children, word_id, weight, desc):
        self.k, self.L, self.scoring, self.weighting = int(k), int(L), int(scoring), int(weighting)
        self.children = np.ascontiguousarray(children, np.int32)
        self.word_id = np.ascontiguousarray(word_id, np.int32)
        self.weight = np.ascontiguousarray(weight, np.float64)
        self.desc = np.ascontiguousarray(desc, np.float32)
        self.n_nodes = self.children.shape[0]
        self.n_words = int((self.word_id >= 0).sum())

    def child_table(self):
        return self.children


def save_blob(v, path):
    ch = np.ascontiguousarray(v.child_table(), np.int32)
    hdr = np.array([MAGIC, v.k, v.L, v.scoring, v.weighting, v.n_nodes, v.desc.shape[1]], np.int32)
    with open(path, "wb") as f:
        f.write(hdr.tobytes())
        f.write(ch.tobytes())
        f.write(np.ascontiguousarray(v.word_id, np.int32).tobytes())
        f.write(np.ascontiguousarray(v.weight, np.float64).tobytes())
        f.write(np.ascontiguousarray(v.desc, np.float32).tobytes())


def load_blob(path):
    raw = open(path, "rb").read()
    magic, k, L, sc, wt, n, dim = np.frombuffer(raw, np.int32, 7)
    if magic != MAGIC:
        raise ValueError("%s: not a vocabulary blob" % path)
    o = 28
    ch = np.frombuffer(raw, np.int32, n * k, o).reshape(n, k)
    o += 4 * n * k
    wid = np.frombuffer(raw, np.int32, n, o)
    o += 4 * n
    w = np.frombuffer(raw, np.float64, n, o)
    o += 8 * n
    d = np.frombuffer(raw, np.float32, n * dim, o).reshape(n, dim)
    return PodVocabulary(k, L, sc, wt, ch, wid, w, d)


def random_vocabulary(seed, k, L, dim=256, scoring=1, zero_weight_frac=0.1):
    """A synthetic k-ary tree of depth L with unit-ish random node descriptors (tests)."""
    rs = np.random.RandomState(seed)
    n = sum(k ** l for l in range(L + 1))
    children = np.full((n, k), -1, np.int32)
    word_id = np.full(n, -1, np.int32)
    weight = np.zeros(n, np.float64)
    desc = rs.normal(size=(n, dim)).astype(np.float32)
    desc /= np.linalg.norm(desc, axis=1, keepdims=True) * rs.uniform(1.0, 3.0, (n, 1)).astype(np.float32)
    nxt, level, words = 1, [0], 0
    for l in range(L):
        new = []
        for p in level:
            # children ids are NOT contiguous with the parent order in DBoW3 files either: shuffle within the level
            ids = list(range(nxt, nxt + k))
            nxt += k
            children[p, :] = ids
            new += ids
        level = new
    for nid in level:
        word_id[nid] = words
        words += 1
        weight[nid] = 0.0 if rs.rand() < zero_weight_frac else rs.uniform(0.01, 3.0)
    return PodVocabulary(k, L, scoring, 0, children, word_id, weight, desc)
