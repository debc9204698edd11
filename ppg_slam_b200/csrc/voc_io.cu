// Host-side reader of the reference's bag-of-words vocabularies.  `System` loads Vocabulary/voc_*_9x3.gz through
// DBoW3::Vocabulary::load (system/src/System.cpp); despite the suffix the files are DBoW3's binary stream
// (Vocabulary::toStream / fromStream of DBoW3, a third-party dependency that is not vendored in the reference):
//
//   uint64 signature 88877711233 | bool compressed | uint32 n_nodes |
//   [compressed: uint32 n_chunks, then one QuickLZ 1.5.0 level-1 packet per 10000 bytes of the stream below] |
//   int32 k, L, scoring, weighting | (n_nodes - 1) x { uint32 id, uint32 parent, double weight,
//   descriptor = int32 cols, rows, type + raw rows } | uint32 n_words | n_words x { uint32 word id, uint32 node id }
//
// Children are appended to their parent in file order, which is the order DBoW3::Vocabulary::transform visits them
// (ties between children keep the first).  Also reads the flat blob tools/export_vocabulary.py writes ('PVOC').
// No device code here; ppg_load_vocabulary hands the arrays to ppg_upload_vocabulary.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "ctx.cuh"

struct ppg_voc_file {
    int k = 0, L = 0, scoring = 0, weighting = 0, n_nodes = 0, dim = 0, n_words = 0;
    std::vector<int32_t> children, word_id;
    std::vector<double> weight;
    std::vector<float> desc;
    std::string err;
};

namespace {

constexpr uint64_t kSignature = 88877711233ull;
constexpr int32_t kBlobMagic = 0x434F5650;  // 'PVOC'

template <typename T>
bool rd(const std::vector<uint8_t>& b, size_t& o, T* v) {
    if (o + sizeof(T) > b.size()) return false;
    memcpy(v, b.data() + o, sizeof(T));
    o += sizeof(T);
    return true;
}

// One QuickLZ 1.5.0 packet (compression level 1, streaming buffer 0) appended to `out`; returns the packet's
// compressed size, 0 on a malformed packet.
size_t qlz_packet(const uint8_t* src, size_t avail, std::vector<uint8_t>& out) {
    if (avail < 3) return 0;
    const uint8_t flags = src[0];
    size_t csize, dsize, hl;
    if (flags & 2) {
        if (avail < 9) return 0;
        uint32_t c, d;
        memcpy(&c, src + 1, 4);
        memcpy(&d, src + 5, 4);
        csize = c;
        dsize = d;
        hl = 9;
    } else {
        csize = src[1];
        dsize = src[2];
        hl = 3;
    }
    if (csize > avail || csize < hl) return 0;
    const size_t base = out.size();
    if (!(flags & 1)) {  // stored
        if (hl + dsize > csize) return 0;
        out.insert(out.end(), src + hl, src + hl + dsize);
        return csize;
    }
    if (((flags >> 2) & 3) != 1) return 0;  // only level 1 (what DBoW3 bundles) is handled
    out.resize(base + dsize);
    uint8_t* dst = out.data() + base;
    std::vector<uint8_t> pad(src, src + csize);
    pad.resize(csize + 8, 0);
    const uint8_t* s = pad.data();
    std::vector<long> table(4096, 0);
    static const uint32_t bitlut[16] = {4, 0, 1, 0, 2, 0, 1, 0, 3, 0, 1, 0, 2, 0, 1, 0};
    auto rd32 = [&](size_t i) {
        uint32_t v;
        memcpy(&v, s + i, 4);
        return v;
    };
    long d = 0, last = (long)dsize - 1, last_matchstart = last - 6 - 4, last_hashed = -1;
    size_t sp = hl;
    uint32_t cword = 1;
    auto hash_upto = [&](long lh, long mx) {
        while (lh < mx) {
            lh++;
            const uint32_t v = dst[lh] | (dst[lh + 1] << 8) | (dst[lh + 2] << 16);
            table[((v >> 12) ^ v) & 0xfff] = lh;
        }
        return lh;
    };
    for (;;) {
        if (sp + 4 > pad.size()) return 0;
        if (cword == 1) {
            cword = rd32(sp);
            sp += 4;
        }
        if (sp + 4 > pad.size()) return 0;
        const uint32_t fetch = rd32(sp);
        if (cword & 1) {
            cword >>= 1;
            const long off = table[(fetch >> 4) & 0xfff];
            long mlen;
            if (fetch & 0xf) {
                mlen = (fetch & 0xf) + 2;
                sp += 2;
            } else {
                mlen = s[sp + 2];
                sp += 3;
            }
            if (d + mlen > (long)dsize || off + mlen > (long)dsize) return 0;
            for (long i = 0; i < mlen; i++) dst[d + i] = dst[off + i];  // may overlap its own output
            d += mlen;
            hash_upto(last_hashed, d - mlen);
            last_hashed = d - 1;
        } else if (d < last_matchstart) {
            const uint32_t n = bitlut[cword & 0xf];
            if (d + 4 > (long)dsize) return 0;
            memcpy(dst + d, s + sp, 4);
            cword >>= n;
            d += n;
            sp += n;
            last_hashed = hash_upto(last_hashed, d - 3);
        } else {
            while (d <= last) {
                if (cword == 1) {
                    sp += 4;
                    cword = 1u << 31;
                }
                if (sp >= pad.size()) return 0;
                dst[d++] = s[sp++];
                cword >>= 1;
            }
            return csize;
        }
    }
}

bool parse_stream(const std::vector<uint8_t>& buf, int n_nodes, ppg_voc_file* v) {
    size_t o = 0;
    int32_t k, L, sc, wt;
    if (!rd(buf, o, &k) || !rd(buf, o, &L) || !rd(buf, o, &sc) || !rd(buf, o, &wt) || k < 1 || k > 1024 || L < 1) {
        v->err = "bad vocabulary header";
        return false;
    }
    v->k = k;
    v->L = L;
    v->scoring = sc;
    v->weighting = wt;
    v->n_nodes = n_nodes;
    v->children.assign((size_t)n_nodes * k, -1);
    v->word_id.assign(n_nodes, -1);
    v->weight.assign(n_nodes, 0.0);
    std::vector<int> nch(n_nodes, 0);
    for (int i = 1; i < n_nodes; i++) {
        uint32_t nid, pid;
        double w;
        int32_t cols, rows, type;
        if (!rd(buf, o, &nid) || !rd(buf, o, &pid) || !rd(buf, o, &w) || !rd(buf, o, &cols) || !rd(buf, o, &rows) ||
            !rd(buf, o, &type)) {
            v->err = "truncated node table";
            return false;
        }
        if (nid >= (uint32_t)n_nodes || pid >= (uint32_t)n_nodes || type != 5 /* CV_32F */ || rows != 1 || cols < 1 ||
            (v->dim && cols != v->dim) || nch[pid] >= k) {
            v->err = "unexpected node record (id / parent / descriptor type)";
            return false;
        }
        if (!v->dim) {
            v->dim = cols;
            v->desc.assign((size_t)n_nodes * cols, 0.f);
        }
        if (o + (size_t)cols * 4 > buf.size()) {
            v->err = "truncated descriptor";
            return false;
        }
        memcpy(&v->desc[(size_t)nid * cols], buf.data() + o, (size_t)cols * 4);
        o += (size_t)cols * 4;
        v->weight[nid] = w;
        v->children[(size_t)pid * k + nch[pid]++] = (int32_t)nid;
    }
    uint32_t n_words;
    if (!rd(buf, o, &n_words)) {
        v->err = "missing word table";
        return false;
    }
    v->n_words = (int)n_words;
    for (uint32_t i = 0; i < n_words; i++) {
        uint32_t wid, nid;
        if (!rd(buf, o, &wid) || !rd(buf, o, &nid) || nid >= (uint32_t)n_nodes) {
            v->err = "truncated word table";
            return false;
        }
        v->word_id[nid] = (int32_t)wid;
    }
    if (o != buf.size()) {
        v->err = "trailing bytes after the word table";
        return false;
    }
    return true;
}

bool parse_file(const char* path, ppg_voc_file* v) {
    FILE* f = fopen(path, "rb");
    if (!f) {
        v->err = std::string("cannot open ") + path;
        return false;
    }
    std::vector<uint8_t> raw;
    uint8_t tmp[65536];
    size_t n;
    while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) raw.insert(raw.end(), tmp, tmp + n);
    fclose(f);
    if (raw.size() >= 28) {
        int32_t hdr[7];
        memcpy(hdr, raw.data(), 28);
        if (hdr[0] == kBlobMagic) {  // flat blob: magic, k, L, scoring, weighting, n_nodes, dim, then the four arrays
            const size_t nn = (size_t)hdr[5], k = (size_t)hdr[1], dim = (size_t)hdr[6];
            if (raw.size() != 28 + nn * k * 4 + nn * 4 + nn * 8 + nn * dim * 4) {
                v->err = "vocabulary blob has the wrong size";
                return false;
            }
            v->k = hdr[1];
            v->L = hdr[2];
            v->scoring = hdr[3];
            v->weighting = hdr[4];
            v->n_nodes = hdr[5];
            v->dim = hdr[6];
            size_t o = 28;
            v->children.resize(nn * k);
            memcpy(v->children.data(), raw.data() + o, nn * k * 4);
            o += nn * k * 4;
            v->word_id.resize(nn);
            memcpy(v->word_id.data(), raw.data() + o, nn * 4);
            o += nn * 4;
            v->weight.resize(nn);
            memcpy(v->weight.data(), raw.data() + o, nn * 8);
            o += nn * 8;
            v->desc.resize(nn * dim);
            memcpy(v->desc.data(), raw.data() + o, nn * dim * 4);
            v->n_words = 0;
            for (int32_t w : v->word_id) v->n_words += w >= 0;
            return true;
        }
    }
    size_t o = 0;
    uint64_t sig;
    uint8_t compressed;
    uint32_t n_nodes;
    if (!rd(raw, o, &sig) || sig != kSignature || !rd(raw, o, &compressed) || !rd(raw, o, &n_nodes) || n_nodes < 2) {
        v->err = "not a DBoW3 binary vocabulary (nor a PVOC blob)";
        return false;
    }
    std::vector<uint8_t> buf;
    if (compressed) {
        uint32_t n_chunks;
        if (!rd(raw, o, &n_chunks)) {
            v->err = "truncated chunk count";
            return false;
        }
        for (uint32_t c = 0; c < n_chunks; c++) {
            const size_t used = o < raw.size() ? qlz_packet(raw.data() + o, raw.size() - o, buf) : 0;
            if (!used) {
                v->err = "malformed QuickLZ packet " + std::to_string(c);
                return false;
            }
            o += used;
        }
    } else {
        buf.assign(raw.begin() + o, raw.end());
    }
    return parse_stream(buf, (int)n_nodes, v);
}

thread_local std::string g_voc_err;

}  // namespace

extern "C" {

int ppg_vocabulary_open(const char* path, ppg_voc_file** out, ppg_vocabulary* view) {
    if (!path || !out || !view) return PPG_ERR_ARG;
    ppg_voc_file* v = new ppg_voc_file();
    if (!parse_file(path, v)) {
        g_voc_err = v->err;
        delete v;
        *out = nullptr;
        return PPG_ERR_WEIGHTS;
    }
    view->k = v->k;
    view->L = v->L;
    view->scoring = v->scoring;
    view->weighting = v->weighting;
    view->n_nodes = v->n_nodes;
    view->dim = v->dim;
    view->children = v->children.data();
    view->word_id = v->word_id.data();
    view->weight = v->weight.data();
    view->desc = v->desc.data();
    *out = v;
    return PPG_OK;
}

void ppg_vocabulary_close(ppg_voc_file* v) { delete v; }

const char* ppg_vocabulary_error(void) { return g_voc_err.c_str(); }

int ppg_load_vocabulary(ppg_ctx* c, const char* path) {
    if (!c || !path) return ppg::set_err(c, PPG_ERR_ARG, "ppg_load_vocabulary: null argument");
    ppg_voc_file* v = nullptr;
    ppg_vocabulary view;
    int rc = ppg_vocabulary_open(path, &v, &view);
    if (rc != PPG_OK) return ppg::set_err(c, rc, std::string("ppg_load_vocabulary: ") + g_voc_err);
    rc = ppg_upload_vocabulary(c, &view);
    ppg_vocabulary_close(v);
    return rc;
}

}  // extern "C"
