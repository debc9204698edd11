// C ABI of libppg_b200.so: context creation (weights, tensor maps, LUTs, buffers) and the extraction path.
// Replaces PPGExtractor::PPGExtractor / run / inference (feature/src/PPGExtractor.cpp:55-156).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <map>

#include "ctx.cuh"
#include "net_direct.cuh"

using namespace ppg;

namespace ppg {

static thread_local std::string g_create_err;

int set_err(const ppg_ctx* c, int code, const std::string& msg) {
    if (c)
        c->err = msg;
    else
        g_create_err = msg;
    return code;
}
int cuda_fail(const ppg_ctx* c, cudaError_t e, const char* what) {
    return set_err(c, PPG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// ------------------------------------------------------------------ weight blob (tools/export_weights.py)
struct Blob {
    std::vector<uint8_t> buf;
    struct T {
        int ndim;
        uint32_t d[4];
        const float* p;
        size_t count() const {
            size_t n = 1;
            for (int i = 0; i < ndim; i++) n *= d[i];
            return n;
        }
    };
    std::map<std::string, T> t;
    bool load(const char* path, std::string& why) {
        FILE* f = fopen(path, "rb");
        if (!f) {
            why = std::string("cannot open weight blob ") + path;
            return false;
        }
        fseek(f, 0, SEEK_END);
        long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        buf.resize(sz);
        size_t rd = fread(buf.data(), 1, sz, f);
        fclose(f);
        if ((long)rd != sz || sz < 12 || memcmp(buf.data(), "PPGW0001", 8) != 0) {
            why = "bad weight blob (magic/size)";
            return false;
        }
        uint32_t n;
        memcpy(&n, buf.data() + 8, 4);
        size_t p = 12;
        for (uint32_t i = 0; i < n; i++) {
            if (p + 76 > buf.size()) {
                why = "truncated weight blob header";
                return false;
            }
            char name[49];
            memcpy(name, buf.data() + p, 48);
            name[48] = 0;
            T e;
            uint32_t nd;
            memcpy(&nd, buf.data() + p + 48, 4);
            memcpy(e.d, buf.data() + p + 52, 16);
            uint64_t off;
            memcpy(&off, buf.data() + p + 68, 8);
            e.ndim = (int)nd;
            if (nd > 4 || off + e.count() * 4 > buf.size()) {
                why = "weight blob entry out of range";
                return false;
            }
            e.p = reinterpret_cast<const float*>(buf.data() + off);
            t[name] = e;
            p += 76;
        }
        return true;
    }
    const T* get(const std::string& k, std::string& why) const {
        auto it = t.find(k);
        if (it == t.end()) {
            why = "weight blob lacks tensor " + k;
            return nullptr;
        }
        return &it->second;
    }
};

// ------------------------------------------------------------------ camera tables (host, double precision)
// cv::undistortPoints(pts, K, D, noArray(), K) for D = (k1,k2,p1,p2): 5 fixed-point iterations
// (PPGExtractor.cpp:223; GeometricCamera.cpp:40).  Evaluated for every integer pixel once.
static void undistort_pinhole(const float* K, const float* D, double u, double v, float* ox, float* oy) {
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const double ifx = 1. / fx, ify = 1. / fy;
    const double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3];
    double x = (u - cx) * ifx, y = (v - cy) * ify;
    const double x0 = x, y0 = y;
    for (int it = 0; it < 5; it++) {
        const double r2 = x * x + y * y;
        const double icdist = 1. / (1 + (k2 * r2 + k1) * r2);
        if (icdist < 0) {
            x = x0;
            y = y0;
            break;
        }
        const double dX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
        const double dY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
        x = (x0 - dX) * icdist;
        y = (y0 - dY) * icdist;
    }
    *ox = (float)(fx * x + cx);
    *oy = (float)(fy * y + cy);
}

// cv::fisheye::undistortPoints(pts, pts, K, D, Mat(), K): Newton on theta, <= 10 iterations, eps 1e-8
// (PPGExtractor.cpp:221).  Non-converged / sign-flipped points get the (-1e6,-1e6) sentinel.
static void undistort_fisheye(const float* K, const float* D, double px, double py, float* ox, float* oy) {
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const double k0 = D[0], k1 = D[1], k2 = D[2], k3 = D[3];
    const double pwx = (px - cx) / fx, pwy = (py - cy) / fy;
    double theta_d = sqrt(pwx * pwx + pwy * pwy);
    const double half_pi = 3.1415926535897932384626433832795 / 2.;
    theta_d = fmin(fmax(-half_pi, theta_d), half_pi);
    bool converged = false;
    double theta = theta_d, scale = 0.0;
    if (fabs(theta_d) > 1e-8) {
        for (int j = 0; j < 10; j++) {
            const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
            const double a = k0 * t2, b = k1 * t4, c = k2 * t6, d = k3 * t8;
            const double fix = (theta * (1 + a + b + c + d) - theta_d) / (1 + 3 * a + 5 * b + 7 * c + 9 * d);
            theta = theta - fix;
            if (fabs(fix) < 1e-8) {
                converged = true;
                break;
            }
        }
        scale = tan(theta) / theta_d;
    } else {
        converged = true;
    }
    const bool flipped = (theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0);
    if (converged && !flipped) {
        *ox = (float)(fx * (pwx * scale) + cx);
        *oy = (float)(fy * (pwy * scale) + cy);
    } else {
        *ox = -1000000.0f;
        *oy = -1000000.0f;
    }
}

// cv::[fisheye::]initUndistortRectifyMap(K, D, I, K, size, CV_32F) (PPGExtractor.cpp:65-71), one pixel.
static void undistort_map_px(const float* K, const float* D, int fisheye, int u, int v, float* mx, float* my) {
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const double x = ((double)u - cx) / fx, y = ((double)v - cy) / fy;
    if (!fisheye) {
        const double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3];
        const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
        const double kr = 1 + (k2 * r2 + k1) * r2;
        const double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2);
        const double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy;
        *mx = (float)(fx * xd + cx);
        *my = (float)(fy * yd + cy);
    } else {
        const double r = sqrt(x * x + y * y), theta = atan(r);
        const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
        const double theta_d = theta * (1 + D[0] * t2 + D[1] * t4 + D[2] * t6 + D[3] * t8);
        const double scale = (r == 0) ? 1.0 : theta_d / r;
        *mx = (float)(fx * x * scale + cx);
        *my = (float)(fy * y * scale + cy);
    }
}

// ------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// NHWC fp16 activation [B][H][W][C] viewed as a 4-D tensor (C, W, H, B); box = 64 channels x 16 x 8 pixels.
static bool make_act_map(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// K-major 2-D operand [rows][cols] (16-bit elements); box = 64 columns x box_rows rows, 128-byte swizzle.
bool make_kmajor_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, bool bf16) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t es[2] = {1, 1};
    return enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
               const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T>
static cudaError_t dalloc(T** p, size_t count) {
    return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
}

// Adds one tensor-core conv layer: folds BatchNorm (if bn prefix given), reorders OIHW fp32 -> [tap][N][Cin] fp16.
static int add_tc_layer(ppg_ctx* c, const Blob& blob, const char* name, const std::string& wkey,
                        const std::string& bnkey, const __half* in, int H, int W, int mode, int relu, void* out,
                        int out_ld, int npad) {
    std::string why;
    const Blob::T* w = blob.get(wkey + ".weight", why);
    const Blob::T* bs = w ? blob.get(wkey + ".bias", why) : nullptr;
    if (!w || !bs) return set_err(c, PPG_ERR_WEIGHTS, why);
    const int cout = w->d[0], cin = w->d[1], kh = w->d[2], kw = w->d[3], taps = kh * kw;
    const int N = npad ? npad : cout;
    if ((taps != 9 && taps != 1) || cin % 64 != 0 || N % 16 != 0 || N > 256 || N < cout)
        return set_err(c, PPG_ERR_WEIGHTS, std::string("unsupported conv shape for ") + name);
    std::vector<float> scale(cout, 1.f), shift(cout);
    for (int o = 0; o < cout; o++) shift[o] = bs->p[o];
    if (!bnkey.empty()) {
        const Blob::T* g = blob.get(bnkey + ".weight", why);
        const Blob::T* be = g ? blob.get(bnkey + ".bias", why) : nullptr;
        const Blob::T* mu = be ? blob.get(bnkey + ".running_mean", why) : nullptr;
        const Blob::T* var = mu ? blob.get(bnkey + ".running_var", why) : nullptr;
        if (!var) return set_err(c, PPG_ERR_WEIGHTS, why);
        for (int o = 0; o < cout; o++) {  // eval-mode BatchNorm2d, eps 1e-5, folded into the conv
            const double s = (double)g->p[o] / sqrt((double)var->p[o] + 1e-5);
            scale[o] = (float)s;
            shift[o] = (float)(((double)bs->p[o] - (double)mu->p[o]) * s + (double)be->p[o]);
        }
    }
    std::vector<__half> hw((size_t)taps * N * cin, __float2half(0.f));
    for (int o = 0; o < cout; o++)
        for (int i = 0; i < cin; i++)
            for (int t = 0; t < taps; t++)
                hw[((size_t)t * N + o) * cin + i] = __float2half(w->p[((size_t)o * cin + i) * taps + t] * scale[o]);
    std::vector<float> hb(N, 0.f);
    for (int o = 0; o < cout; o++) hb[o] = shift[o];
    TcLayerInfo li;
    memset(&li, 0, sizeof(li));
    li.name = name;
    li.in = in;
    li.H = H;
    li.W = W;
    li.cin = cin;
    li.cout = cout;
    li.N = N;
    li.taps = taps;
    li.mode = mode;
    li.relu = relu;
    li.out = out;
    li.out_ld = out_ld;
    PPG_CUDA(c, dalloc(&li.w, hw.size()));
    PPG_CUDA(c, dalloc(&li.bias, hb.size()));
    PPG_CUDA(c, cudaMemcpy(li.w, hw.data(), hw.size() * sizeof(__half), cudaMemcpyHostToDevice));
    PPG_CUDA(c, cudaMemcpy(li.bias, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice));
    conv_tc_plan(li.L, c->maxB, H, W, cin, N, taps, mode, relu, li.bias, out, out_ld);
    li.L.wgt = li.w;
    memset(&li.L.hb, 0, sizeof(li.L.hb));
    memcpy(li.L.hb.v, hb.data(), hb.size() * sizeof(float));
    if (!make_act_map(&li.L.mapA, in, c->maxB, H, W, cin, li.L.box_w, li.L.box_h) ||
        !make_kmajor_map(&li.L.mapB, li.w, (uint64_t)taps * N, cin, li.L.v3 == 3 ? 128u : (uint32_t)N, false))
        return set_err(c, PPG_ERR_CUDA, std::string("cuTensorMapEncodeTiled failed for ") + name);
    c->tc.push_back(li);
    return PPG_OK;
}

void stage_mark(ppg_ctx* c, const char* name) {
    if (!c->profiling) return;
    if (c->n_ev >= (int)c->ev.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->ev.push_back(e);
        c->ev_names.push_back(name);
    }
    c->ev_names[c->n_ev] = name;
    cudaEventRecord(c->ev[c->n_ev], c->st);
    c->n_ev++;
}

static void fill_out(const ppg_ctx* c, int f, ppg_frame_out* o) {
    const OutLayout& L = c->post.lay;
    const uint8_t* base = c->h_out + (size_t)f * L.total;
    const int* hdr = reinterpret_cast<const int*>(base + L.hdr);
    o->n_kp = hdr[HDR_NKP];
    o->n_edges = hdr[HDR_NEDGES];
    o->n_colines = hdr[HDR_NCOL];
    o->status = (uint32_t)hdr[HDR_STATUS];
    o->n_candidates = hdr[HDR_NCAND];
    o->n_pairs_tested_ok = hdr[HDR_NPASS];
    o->n_candidate_lines = hdr[HDR_NLINES];
    o->nms_rounds = hdr[HDR_NMS_ROUNDS];
    o->kp_x = reinterpret_cast<const float*>(base + L.kp_x);
    o->kp_y = reinterpret_cast<const float*>(base + L.kp_y);
    o->kp_px = reinterpret_cast<const int32_t*>(base + L.px);
    o->kp_py = reinterpret_cast<const int32_t*>(base + L.py);
    o->kp_score = reinterpret_cast<const float*>(base + L.score);
    o->kp_xun = reinterpret_cast<const float*>(base + L.xun);
    o->kp_yun = reinterpret_cast<const float*>(base + L.yun);
    o->kp_out = base + L.kout;
    o->edge_start = reinterpret_cast<const int32_t*>(base + L.edge_s);
    o->edge_end = reinterpret_cast<const int32_t*>(base + L.edge_e);
    o->edge_score = reinterpret_cast<const float*>(base + L.edge_score);
    o->conn_off = reinterpret_cast<const int32_t*>(base + L.conn_off);
    o->conn_idx = reinterpret_cast<const int32_t*>(base + L.conn_idx);
    o->col_off = reinterpret_cast<const int32_t*>(base + L.col_off);
    o->col_pairs = reinterpret_cast<const int32_t*>(base + L.col_pairs);
    o->desc = reinterpret_cast<const float*>(base + L.desc);
    for (int k = 0; k < 8; k++) o->diag[k] = k < HDR_WORDS - HDR_DIAG ? hdr[HDR_DIAG + k] : 0;
}

static PostMark post_mark(ppg_ctx* c) {
    PostMark m;
    m.fn = [](void* user, const char* name) { stage_mark(static_cast<ppg_ctx*>(user), name); };
    m.user = c;
    return m;
}

// Post-processing launches shared by ppg_run and ppg_extract_from_maps.
static int run_post(ppg_ctx* c, int n) {
    PostParams p = c->post;
    p.B = n;
    if (c->maps_from_caller) {
        p.prob = c->prob_in;
        p.heat_raw = c->heat_in;
        p.desc = c->desc_in;
    }
    const PostMark mk = post_mark(c);
    PPG_CUDA(c, post_keypoints_launch(p, c->st, &c->launches, mk));
    PPG_CUDA(c, post_heat_launch(p, c->st, &c->launches, mk));
    PPG_CUDA(c, post_lines_launch(p, c->st, &c->launches, mk));
    PPG_CUDA(c, post_desc_launch(p, c->st, &c->launches, mk));
    return PPG_OK;
}

}  // namespace ppg

// =================================================================================================
extern "C" {

int ppg_api_version(void) { return PPG_API_VERSION; }

void ppg_default_config(ppg_config* c) {
    memset(c, 0, sizeof(*c));
    c->junction_thresh = 1.0f / 128.0f;  // PPGExtractor.cpp:44-53
    c->junction_nms_radius = 4;
    c->junction_max_num = 500;
    c->line_valid_thresh = 1.0e-2f;
    c->line_valid_ratio = 0.3f;
    c->line_dist_thresh = 2.0f;
    c->heatmap_refine_sz = 16;
    c->line_heatmap_thresh = 0.2f;
    c->line_inlier_rate = 0.8f;
    c->th_low = 0.7f;  // Matcher.cpp:12-13
    c->th_high = 0.8f;
    c->max_batch = 1;
    c->max_edges = 4096;
    c->max_colines = 2048;
    c->max_map_points = 65536;
}

const char* ppg_last_error(const ppg_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

void ppg_destroy(ppg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->dev);
    if (c->st) cudaStreamSynchronize(c->st);
    for (auto& g : c->graphs) cudaGraphExecDestroy(g.second);
    for (cudaEvent_t e : {c->ev_feat, c->ev_heat, c->ev_dmap, c->ev_kp, c->ev_desc})
        if (e) cudaEventDestroy(e);
    if (c->st2) cudaStreamDestroy(c->st2);
    if (c->st3) cudaStreamDestroy(c->st3);
    comm_destroy(c);
    assoc_destroy(c);
    bow_destroy(c);
    for (auto& l : c->tc) {
        cudaFree(l.w);
        cudaFree(l.bias);
    }
    void* bufs[] = {c->w1a,  c->b1a,     c->we3,      c->be3,      c->we1,        c->be1b,       c->gray,
                    c->a1,   c->a2,      c->a3,       c->a4,       c->a5,         c->a6,         c->a7,
                    c->feat, c->p1,      c->d1,       c->e1,       c->e2,         c->jlogits,    c->desc,
                    c->prob, c->heat_raw, c->heat_ref, c->heat_final, c->prob_in,  c->heat_in,    c->desc_in,
                    c->undist_lut, c->remap_lut, c->d_out, c->post.state, c->post.state2, c->post.cand, c->post.counters,
                    c->post.pair_bits, c->post.row_cnt, c->post.l_score, c->post.l_edge, c->post.row_prefix,
                    c->post.row_off, c->post.c_se, c->post.c_dist, c->post.c_dirf, c->post.c_dirb, c->post.inter,
                    c->post.inter_cnt, c->post.inter_off, c->post.inter_pool, c->post.alive_g, c->post.sym_bits};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (c->h_gray) cudaFreeHost(c->h_gray);
    if (c->h_out) cudaFreeHost(c->h_out);
    for (auto e : c->ev) cudaEventDestroy(e);
    if (c->t0) cudaEventDestroy(c->t0);
    if (c->t1) cudaEventDestroy(c->t1);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

int ppg_create(const ppg_config* cfg, ppg_ctx** out) {
    if (!cfg || !out) return set_err(nullptr, PPG_ERR_ARG, "null argument");
    *out = nullptr;
    const int W = cfg->width, H = cfg->height;
    if (W <= 0 || H <= 0 || W % 16 || H % 16)
        return set_err(nullptr, PPG_ERR_ARG, "width and height must be positive multiples of 16");
    if (cfg->heatmap_refine_sz != 16) return set_err(nullptr, PPG_ERR_ARG, "only heatmap_refine_sz = 16 is supported");
    if (cfg->junction_max_num < 1 || cfg->junction_max_num > POST_MAX_KP)
        return set_err(nullptr, PPG_ERR_ARG, "junction_max_num must be in [1, 1024]");
    if (cfg->junction_nms_radius < 1 || cfg->junction_nms_radius > 8)
        return set_err(nullptr, PPG_ERR_ARG, "junction_nms_radius must be in [1, 8]");
    if (cfg->max_batch < 1 || cfg->max_edges < 1 || cfg->max_colines < 1)
        return set_err(nullptr, PPG_ERR_ARG, "max_batch / max_edges / max_colines must be >= 1");
    if (!cfg->weights_path) return set_err(nullptr, PPG_ERR_WEIGHTS, "weights_path is null");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return set_err(nullptr, PPG_ERR_CUDA,
                       std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(ce));
    if (cfg->device < 0 || cfg->device >= ndev) return set_err(nullptr, PPG_ERR_ARG, "bad device ordinal");
    cudaDeviceProp prop;
    if ((ce = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) return cuda_fail(nullptr, ce, "props");
    if (prop.major != 10)
        return set_err(nullptr, PPG_ERR_CUDA, "this library is built for sm_100a (B200) only");

    Blob blob;
    std::string why;
    if (!blob.load(cfg->weights_path, why)) return set_err(nullptr, PPG_ERR_WEIGHTS, why);

    ppg_ctx* c = new ppg_ctx();
    struct Guard {
        ppg_ctx* c;
        bool ok = false;
        ~Guard() {
            if (!ok) {
                g_create_err = c->err;
                ppg_destroy(c);
            }
        }
    } guard{c};
    c->cfg = *cfg;
    c->weights_path = cfg->weights_path;
    c->cfg.weights_path = c->weights_path.c_str();
    c->dev = cfg->device;
    c->num_sms = prop.multiProcessorCount;
    if (const char* e = getenv("PPG_GRAPH")) c->use_graph = atoi(e) != 0;
    // PPG_GRAPH_LARGE=1: replay large batches as one graph too (one driver call per step instead of ~35: for hosts
    // whose cores are the bottleneck; on an otherwise idle host eager launches were measured 3 % faster at batch 32)
    if (const char* e = getenv("PPG_GRAPH_LARGE")) c->graph_large = atoi(e) != 0;
    c->H = H;
    c->W = W;
    c->Hc = H / 8;
    c->Wc = W / 8;
    c->maxB = cfg->max_batch;
    const int B = c->maxB, Hc = c->Hc, Wc = c->Wc;
    PPG_CUDA(c, cudaSetDevice(c->dev));
    PPG_CUDA(c, cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    PPG_CUDA(c, cudaStreamCreateWithFlags(&c->st2, cudaStreamNonBlocking));
    PPG_CUDA(c, cudaStreamCreateWithFlags(&c->st3, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&c->ev_feat, &c->ev_heat, &c->ev_dmap, &c->ev_kp, &c->ev_desc})
        PPG_CUDA(c, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    if (const char* e = getenv("PPG_FORK")) c->use_fork = atoi(e) != 0;
    PPG_CUDA(c, cudaEventCreate(&c->t0));
    PPG_CUDA(c, cudaEventCreate(&c->t1));

    // ---- activations
    const size_t HW = (size_t)H * W;
    PPG_CUDA(c, dalloc(&c->gray, B * HW));
    PPG_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&c->h_gray), B * HW));
    // conv1a's full-resolution 64-channel map (46 MB per EuRoC frame) is not materialised when conv1b computes it on
    // the fly (the default, conv_t64.cu): one frame's worth is kept for ppg_selftest_conv
    const char* kenv = getenv("PPG_CONV_KERNEL");
    const bool fuse_conv1a = !kenv || atoi(kenv) >= 5;
    PPG_CUDA(c, dalloc(&c->a1, (fuse_conv1a ? (size_t)1 : (size_t)B) * HW * 64));
    PPG_CUDA(c, dalloc(&c->a2, B * HW / 4 * 64));
    PPG_CUDA(c, dalloc(&c->a3, B * HW / 4 * 64));
    PPG_CUDA(c, dalloc(&c->a4, B * HW / 16 * 64));
    PPG_CUDA(c, dalloc(&c->a5, B * HW / 16 * 128));
    PPG_CUDA(c, dalloc(&c->a6, B * HW / 64 * 128));
    PPG_CUDA(c, dalloc(&c->a7, B * HW / 64 * 128));
    PPG_CUDA(c, dalloc(&c->feat, B * HW / 64 * 128));
    PPG_CUDA(c, dalloc(&c->p1, B * HW / 64 * 256));
    PPG_CUDA(c, dalloc(&c->d1, B * HW / 64 * 256));
    PPG_CUDA(c, dalloc(&c->e1, B * HW / 16 * 64));
    PPG_CUDA(c, dalloc(&c->e2, B * HW / 4 * 16));
    PPG_CUDA(c, dalloc(&c->desc, B * HW / 64 * 256));
    PPG_CUDA(c, dalloc(&c->prob, B * HW));
    PPG_CUDA(c, dalloc(&c->heat_raw, B * HW));
    PPG_CUDA(c, dalloc(&c->heat_ref, B * HW));
    PPG_CUDA(c, dalloc(&c->heat_final, B * HW));

    // ---- weights
    {
        const Blob::T* w = blob.get("backbone.conv1a.weight", why);
        const Blob::T* b = w ? blob.get("backbone.conv1a.bias", why) : nullptr;
        if (!b || w->count() != 576) return set_err(c, PPG_ERR_WEIGHTS, why.empty() ? "conv1a shape" : why);
        PPG_CUDA(c, dalloc(&c->w1a, 576));
        PPG_CUDA(c, dalloc(&c->b1a, 64));
        PPG_CUDA(c, cudaMemcpy(c->w1a, w->p, 576 * 4, cudaMemcpyHostToDevice));
        PPG_CUDA(c, cudaMemcpy(c->b1a, b->p, 64 * 4, cudaMemcpyHostToDevice));
    }
    int rc;
#define ADD(name, wkey, bn, in, h, w, mode, relu, outp, ld, npad)                                           \
    if ((rc = add_tc_layer(c, blob, name, wkey, bn, in, h, w, mode, relu, outp, ld, npad)) != PPG_OK) return rc;
    ADD("conv1b", "backbone.conv1b", "", c->a1, H, W, EPI_F16_POOL, 1, c->a2, 64, 0)
    {
        // conv1a runs inside conv1b's producer warps (conv_t64.cu) unless PPG_CONV_KERNEL asks for an older arrangement
        if (fuse_conv1a && c->tc.back().L.v3 == 1) conv_t64_fuse_conv1a(c->tc.back().L, c->gray, c->w1a, c->b1a);
        if (fuse_conv1a && c->tc.back().L.v3 != 2) return set_err(c, PPG_ERR_ARG, "conv1b cannot run the fused kernel");
    }
    ADD("conv2a", "backbone.conv2a", "", c->a2, H / 2, W / 2, EPI_F16, 1, c->a3, 64, 0)
    ADD("conv2b", "backbone.conv2b", "", c->a3, H / 2, W / 2, EPI_F16_POOL, 1, c->a4, 64, 0)
    ADD("conv3a", "backbone.conv3a", "", c->a4, H / 4, W / 4, EPI_F16, 1, c->a5, 128, 0)
    ADD("conv3b", "backbone.conv3b", "", c->a5, H / 4, W / 4, EPI_F16_POOL, 1, c->a6, 128, 0)
    ADD("conv4a", "backbone.conv4a", "", c->a6, Hc, Wc, EPI_F16, 1, c->a7, 128, 0)
    ADD("conv4b", "backbone.conv4b", "", c->a7, Hc, Wc, EPI_F16, 1, c->feat, 128, 0)
    ADD("convPa", "junction.convPa", "", c->feat, Hc, Wc, EPI_F16, 1, c->p1, 256, 0)
    // convPb's epilogue is the softmax + depth-to-space: it writes the H x W junction map, not logits
    ADD("convPb", "junction.convPb", "", c->p1, Hc, Wc, EPI_SOFTMAX_D2S, 0, c->prob, 80, 80)
    ADD("convDa", "descriptor.convDa", "", c->feat, Hc, Wc, EPI_F16, 1, c->d1, 256, 0)
    ADD("convDb", "descriptor.convDb", "", c->d1, Hc, Wc, EPI_F32, 0, c->desc, 256, 0)
    ADD("edge0", "edge.conv_block_lst.0.0", "edge.conv_block_lst.0.1", c->feat, Hc, Wc, EPI_F16_PS2, 1, c->e1, 64, 0)
    ADD("edge1", "edge.conv_block_lst.1.0", "edge.conv_block_lst.1.1", c->e1, 2 * Hc, 2 * Wc, EPI_F16_PS2, 1, c->e2,
        16, 0)
#undef ADD
    {  // edge tail: conv3x3 16->16 + BN folded, reordered to [co][ky][kx][ci]; conv1x1 4->2
        const Blob::T* w = blob.get("edge.conv_block_lst.2.0.weight", why);
        const Blob::T* b = w ? blob.get("edge.conv_block_lst.2.0.bias", why) : nullptr;
        const Blob::T* g = b ? blob.get("edge.conv_block_lst.2.1.weight", why) : nullptr;
        const Blob::T* be = g ? blob.get("edge.conv_block_lst.2.1.bias", why) : nullptr;
        const Blob::T* mu = be ? blob.get("edge.conv_block_lst.2.1.running_mean", why) : nullptr;
        const Blob::T* var = mu ? blob.get("edge.conv_block_lst.2.1.running_var", why) : nullptr;
        const Blob::T* w1 = var ? blob.get("edge.conv_block_lst.3.weight", why) : nullptr;
        const Blob::T* b1 = w1 ? blob.get("edge.conv_block_lst.3.bias", why) : nullptr;
        if (!b1 || w->count() != 16 * 16 * 9 || w1->count() != 8) return set_err(c, PPG_ERR_WEIGHTS, why);
        std::vector<float> w3(16 * 9 * 16), b3(16);
        for (int co = 0; co < 16; co++) {
            const double s = (double)g->p[co] / sqrt((double)var->p[co] + 1e-5);
            b3[co] = (float)(((double)b->p[co] - (double)mu->p[co]) * s + (double)be->p[co]);
            for (int ci = 0; ci < 16; ci++)
                for (int t = 0; t < 9; t++)
                    w3[(co * 9 + t) * 16 + ci] = (float)((double)w->p[(co * 16 + ci) * 9 + t] * s);
        }
        PPG_CUDA(c, dalloc(&c->we3, w3.size()));
        PPG_CUDA(c, dalloc(&c->be3, 16));
        PPG_CUDA(c, dalloc(&c->we1, 8));
        PPG_CUDA(c, dalloc(&c->be1b, 2));
        PPG_CUDA(c, cudaMemcpy(c->we3, w3.data(), w3.size() * 4, cudaMemcpyHostToDevice));
        PPG_CUDA(c, cudaMemcpy(c->be3, b3.data(), 64, cudaMemcpyHostToDevice));
        PPG_CUDA(c, cudaMemcpy(c->we1, w1->p, 32, cudaMemcpyHostToDevice));
        PPG_CUDA(c, cudaMemcpy(c->be1b, b1->p, 8, cudaMemcpyHostToDevice));
    }

    // ---- camera tables (PPGExtractor.cpp:58-74)
    {
        std::vector<float2> lut(HW);
        for (int v = 0; v < H; v++)
            for (int u = 0; u < W; u++) {
                float x, y;
                if (cfg->fisheye)
                    undistort_fisheye(cfg->K, cfg->D, u, v, &x, &y);
                else
                    undistort_pinhole(cfg->K, cfg->D, u, v, &x, &y);
                lut[(size_t)v * W + u] = make_float2(x, y);
            }
        PPG_CUDA(c, dalloc(&c->undist_lut, HW));
        PPG_CUDA(c, cudaMemcpy(c->undist_lut, lut.data(), HW * sizeof(float2), cudaMemcpyHostToDevice));
        const bool do_remap = cfg->D[0] != 0.0f;  // :261
        if (do_remap) {
            std::vector<int2> rl(HW);
            for (int v = 0; v < H; v++)
                for (int u = 0; u < W; u++) {
                    float mx, my;
                    undistort_map_px(cfg->K, cfg->D, cfg->fisheye, u, v, &mx, &my);
                    // 1/32 fixed point as cv::remap; packed for remap_kernel (post.cu): offset of the top-left texel,
                    // the two 5-bit fractions and which of the four texels lie inside the image
                    const int sx = (int)lrint((double)(mx * 32.0f)), sy = (int)lrint((double)(my * 32.0f));
                    const int ix = sx >> 5, iy = sy >> 5;
                    const bool x0 = ix >= 0 && ix < W, x1 = ix + 1 >= 0 && ix + 1 < W;
                    const bool y0 = iy >= 0 && iy < H, y1 = iy + 1 >= 0 && iy + 1 < H;
                    const unsigned bits = (unsigned)(sx & 31) | ((unsigned)(sy & 31) << 5) | ((unsigned)(y0 && x0) << 10) |
                                          ((unsigned)(y0 && x1) << 11) | ((unsigned)(y1 && x0) << 12) |
                                          ((unsigned)(y1 && x1) << 13);
                    const long long off = (long long)iy * W + ix;
                    const bool any = (bits >> 10) != 0;  // far outside: the offset is never dereferenced
                    rl[(size_t)v * W + u] = make_int2(any ? (int)off : 0, (int)bits);
                }
            PPG_CUDA(c, dalloc(&c->remap_lut, HW));
            PPG_CUDA(c, cudaMemcpy(c->remap_lut, rl.data(), HW * sizeof(int2), cudaMemcpyHostToDevice));
        }
        // GeometricCamera::InitializeImageBounds (GeometricCamera.cpp:26-61)
        if (!cfg->fisheye) {
            float ux[4], uy[4];
            const double cx4[4] = {0, (double)W, 0, (double)W}, cy4[4] = {0, 0, (double)H, (double)H};
            for (int k = 0; k < 4; k++) undistort_pinhole(cfg->K, cfg->D, cx4[k], cy4[k], &ux[k], &uy[k]);
            c->minX = (int)fminf(ux[0], ux[2]);
            c->maxX = (int)fmaxf(ux[1], ux[3]);
            c->minY = (int)fminf(uy[0], uy[1]);
            c->maxY = (int)fmaxf(uy[2], uy[3]);
        } else {
            c->minX = 0;
            c->minY = 0;
            c->maxX = W;
            c->maxY = H;
        }
        c->wInv = 64.0f / (float)(c->maxX - c->minX);
        c->hInv = 48.0f / (float)(c->maxY - c->minY);
    }

    // ---- post-processing parameters and scratch
    PostParams& p = c->post;
    memset(&p, 0, sizeof(p));
    p.B = B;
    p.H = H;
    p.W = W;
    p.Hc = Hc;
    p.Wc = Wc;
    p.fisheye = cfg->fisheye;
    p.do_remap = cfg->D[0] != 0.0f;
    p.junction_thresh = cfg->junction_thresh;
    p.nms_radius = cfg->junction_nms_radius;
    p.max_kp = cfg->junction_max_num;
    p.line_valid_thresh = cfg->line_valid_thresh;
    p.line_valid_ratio = cfg->line_valid_ratio;
    p.line_dist_thresh = cfg->line_dist_thresh;
    p.line_heatmap_thresh = cfg->line_heatmap_thresh;
    p.line_inlier_rate = cfg->line_inlier_rate;
    p.inv_scale = 1.0f / sqrtf((float)(H * H + W * W));  // :74
    p.prob = c->prob;
    p.heat_raw = c->heat_raw;
    p.heat_ref = c->heat_ref;
    p.heat_final = c->heat_final;
    p.desc = c->desc;
    p.undist_lut = c->undist_lut;
    p.remap_lut = c->remap_lut;
    post_plan_nms(p);
    if (const char* e = getenv("PPG_NMS_GLOBAL"))  // A/B switch: force the global-memory NMS variant
        if (atoi(e)) {
            p.nms_smem = 0;
            p.acc_cap = 8192;
        }
    p.pair_words = (p.max_kp + 31) / 32;
    p.pair_cap = 16384;
    {
        const size_t budget = 226 * 1024;
        const size_t fixed = post_lines_fixed_smem(p.max_kp, p.pair_words) + 64;
        int deg = fixed < budget ? (int)((budget - fixed) / ((size_t)p.max_kp * 2)) : 0;
        if (deg > 128) deg = 128;
        p.deg_cap = deg;
        if (deg < 8 || post_lines_filter_smem(p) > budget)
            return set_err(c, PPG_ERR_ARG, "junction_max_num too large for the line-graph kernels");
    }
    p.lay = make_out_layout(p.max_kp, cfg->max_edges, cfg->max_colines);
    PPG_CUDA(c, dalloc(&p.state, B * HW));
    PPG_CUDA(c, dalloc(&p.state2, B * HW / 4 + 16));
    PPG_CUDA(c, dalloc(&p.cand, B * HW));
    PPG_CUDA(c, dalloc(&p.counters, (size_t)B * 8));
    PPG_CUDA(c, dalloc(&p.pair_bits, (size_t)B * p.max_kp * p.pair_words));
    PPG_CUDA(c, dalloc(&p.sym_bits, (size_t)B * p.max_kp * p.pair_words));
    PPG_CUDA(c, dalloc(&p.row_cnt, (size_t)B * p.max_kp));
    PPG_CUDA(c, dalloc(&p.row_prefix, (size_t)B * p.max_kp * p.pair_words));
    PPG_CUDA(c, dalloc(&p.row_off, (size_t)B * (p.max_kp + 1)));
    PPG_CUDA(c, dalloc(&p.c_se, (size_t)B * p.pair_cap));
    PPG_CUDA(c, dalloc(&p.c_dist, (size_t)B * p.pair_cap));
    PPG_CUDA(c, dalloc(&p.c_dirf, (size_t)B * p.pair_cap));
    PPG_CUDA(c, dalloc(&p.c_dirb, (size_t)B * p.pair_cap));
    PPG_CUDA(c, dalloc(&p.inter, (size_t)B * p.pair_cap * 32));
    PPG_CUDA(c, dalloc(&p.inter_cnt, (size_t)B * p.pair_cap));
    PPG_CUDA(c, dalloc(&p.inter_off, (size_t)B * p.pair_cap));
    p.pool_cap = p.pair_cap * 8;
    PPG_CUDA(c, dalloc(&p.inter_pool, (size_t)B * p.pool_cap));
    PPG_CUDA(c, dalloc(&p.alive_g, (size_t)B * p.max_kp * p.pair_words));
    PPG_CUDA(c, dalloc(&p.l_score, (size_t)B * p.pair_cap));
    PPG_CUDA(c, dalloc(&p.l_edge, (size_t)B * p.pair_cap));
    PPG_CUDA(c, dalloc(&c->d_out, (size_t)B * p.lay.total));
    PPG_CUDA(c, cudaMemset(c->d_out, 0, (size_t)B * p.lay.total));
    PPG_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&c->h_out), (size_t)B * p.lay.total));
    memset(c->h_out, 0, (size_t)B * p.lay.total);
    p.out = c->d_out;
    p.scan_fused = 0;
    // PPG_FUSE_SCAN=1: the junction head's epilogue (convPb: softmax + depth-to-space) also does the threshold scan of
    // detectKeyPoint -- candidate list, NMS state map and counters come out of the registers that hold the
    // probabilities.  Measured on B200: convPb 0.037 -> 0.061 ms, scan + NMS 0.107 -> 0.082 ms per 32 frames, i.e. no
    // gain (the epilogue is the critical path of that small launch), so the separate scan kernel stays the default.
    c->fuse_scan = false;
    if (const char* e = getenv("PPG_FUSE_SCAN")) c->fuse_scan = atoi(e) != 0;
    if (c->fuse_scan)
        for (auto& l : c->tc)
            if (l.L.p.mode == EPI_SOFTMAX_D2S) {
                l.L.p.scan_cand = p.cand;
                l.L.p.scan_counters = p.counters;
                l.L.p.scan_state2 = p.nms_smem ? p.state2 : nullptr;
                l.L.p.scan_state = p.nms_smem ? nullptr : p.state;
                l.L.p.scan_thresh = p.junction_thresh;
                l.L.p.scan_radius = p.nms_radius;
            }
    PPG_CUDA(c, post_init_attrs(p));
    PPG_CUDA(c, cudaDeviceSynchronize());
    guard.ok = true;
    *out = c;
    return PPG_OK;
}

// -------------------------------------------------------------------------------------------------
static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

int ppg_upload_frames(ppg_ctx* c, const uint8_t* const* gray, const int* stride, int n) {
    if (!c || !gray || n < 1 || n > c->maxB) return set_err(c, PPG_ERR_ARG, "ppg_upload_frames: bad arguments");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    const size_t HW = (size_t)c->H * c->W;
    bool pinned = true, packed = true;
    for (int f = 0; f < n; f++) {
        if (!gray[f]) return set_err(c, PPG_ERR_ARG, "null frame pointer");
        const int s = stride ? stride[f] : c->W;
        if (s < c->W) return set_err(c, PPG_ERR_ARG, "row stride smaller than the image width");
        pinned = pinned && is_pinned(gray[f]);
        packed = packed && s == c->W && (f == 0 || gray[f] == gray[f - 1] + HW);
    }
    if (pinned) {
        // the caller's buffers are page-locked: DMA from where they are, nothing to wait for
        if (packed) {
            PPG_CUDA(c, cudaMemcpyAsync(c->gray, gray[0], n * HW, cudaMemcpyHostToDevice, c->st));
        } else {
            for (int f = 0; f < n; f++)
                PPG_CUDA(c, cudaMemcpy2DAsync(c->gray + f * HW, c->W, gray[f], stride ? stride[f] : c->W, c->W, c->H,
                                              cudaMemcpyHostToDevice, c->st));
        }
        return PPG_OK;
    }
    // pageable frames: through the pinned staging buffer; the previous batch's DMA out of it must be done first
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    for (int f = 0; f < n; f++) {
        const int s = stride ? stride[f] : c->W;
        uint8_t* dst = c->h_gray + f * HW;
        if (s == c->W)
            memcpy(dst, gray[f], HW);
        else
            for (int y = 0; y < c->H; y++) memcpy(dst + (size_t)y * c->W, gray[f] + (size_t)y * s, c->W);
    }
    PPG_CUDA(c, cudaMemcpyAsync(c->gray, c->h_gray, n * HW, cudaMemcpyHostToDevice, c->st));
    return PPG_OK;
}

static int enqueue_run(ppg_ctx* c, int n);

int ppg_run(ppg_ctx* c, int n) {
    if (!c || n < 1 || n > c->maxB) return set_err(c, PPG_ERR_ARG, "ppg_run: bad frame count");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    c->maps_from_caller = false;
    c->last_batch = n;
    c->n_ev = 0;
    // Graphs pay at small batches, where the ~30 launches are latency (p50 at batch 1: 0.775 -> 0.734 ms); at batch
    // 32 with several contexts in flight eager launches interleave better across streams (9.6 k vs 9.3 k frames/s).
    if (!c->use_graph || c->profiling || (n > 8 && !c->graph_large)) return enqueue_run(c, n);
    // The launch sequence of a batch size is fixed (same kernels, same parameters): replay it as one CUDA graph.
    // The first call with a size runs eagerly (one-time kernel attributes are set on that path), the second one is
    // captured.
    auto it = c->graphs.find(n);
    if (it != c->graphs.end()) {
        PPG_CUDA(c, cudaGraphLaunch(it->second, c->st));
        c->launches += c->graph_launches[n];
        return PPG_OK;
    }
    if (c->run_calls[n]++ == 0) return enqueue_run(c, n);
    const long long l0 = c->launches;
    PPG_CUDA(c, cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue_run(c, n);
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(c->st, &g);
    if (rc != PPG_OK || ce != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        c->launches = l0;
        c->use_graph = false;  // capture not possible here: stay on the eager path
        cudaGetLastError();
        return enqueue_run(c, n);
    }
    cudaGraphExec_t ex = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    if (ie != cudaSuccess || !ex) {
        c->launches = l0;
        c->use_graph = false;
        cudaGetLastError();
        return enqueue_run(c, n);
    }
    c->graphs[n] = ex;
    c->graph_launches[n] = (int)(c->launches - l0);
    c->launches = l0;
    PPG_CUDA(c, cudaGraphLaunch(ex, c->st));
    c->launches += c->graph_launches[n];
    return PPG_OK;
}

static int enqueue_run(ppg_ctx* c, int n) {
    stage_mark(c, "start");
    if (c->tc.empty() || c->tc[0].L.v3 != 2) {  // otherwise conv1b computes it on the fly
        PPG_CUDA(c, conv1a_tc_launch(c->gray, c->w1a, c->b1a, c->a1, n, c->H, c->W, c->st));
        c->launches++;
        stage_mark(c, "conv1a");
    }
    // Small batches leave most of the GPU idle inside every kernel, so the branches of the network and of the
    // post-processing that do not depend on one another are forked onto side streams:
    //   st : backbone -> junction head -> keypoints (scan, NMS)   ...join heat... -> point-pair graph  ...join desc
    //   st2: (after the backbone) edge head -> edge tail -> refine / remap
    //   st3: (after the backbone) descriptor head; (after the keypoints) descriptor sampling
    // Large batches fill the GPU by themselves and stay on one stream (as do profiling runs: stage events).
    const bool fork = c->use_fork && !c->profiling && n <= 8 && c->st2 && c->st3;
    auto stream_of = [&](const char* name) {
        if (!fork) return c->st;
        if (!strncmp(name, "edge", 4)) return c->st2;
        if (!strncmp(name, "convD", 5)) return c->st3;
        return c->st;
    };
    bool forked = false;
    if (c->fuse_scan)  // the junction head's epilogue appends candidates: counters start at zero
        PPG_CUDA(c, cudaMemsetAsync(c->post.counters, 0, sizeof(int) * 8 * n, c->st));
    for (auto& l : c->tc) {
        cudaStream_t s = stream_of(l.name);
        const bool head = !strncmp(l.name, "convP", 5) || !strncmp(l.name, "convD", 5) || !strncmp(l.name, "edge", 4);
        if (fork && !forked && head) {  // first head layer: the backbone (feature map) is complete here
            PPG_CUDA(c, cudaEventRecord(c->ev_feat, c->st));
            PPG_CUDA(c, cudaStreamWaitEvent(c->st2, c->ev_feat, 0));
            PPG_CUDA(c, cudaStreamWaitEvent(c->st3, c->ev_feat, 0));
            forked = true;
        }
        PPG_CUDA(c, conv_tc_launch(l.L, n, c->num_sms, s));
        c->launches++;
        stage_mark(c, l.name);
    }
    if (fork && !forked) {  // no head layer ran on a side stream (cannot happen with the shipped networks)
        PPG_CUDA(c, cudaEventRecord(c->ev_feat, c->st));
        PPG_CUDA(c, cudaStreamWaitEvent(c->st2, c->ev_feat, 0));
        PPG_CUDA(c, cudaStreamWaitEvent(c->st3, c->ev_feat, 0));
    }
    PostParams p = c->post;
    p.B = n;
    p.scan_fused = c->fuse_scan ? 1 : 0;
    cudaStream_t s_heat = fork ? c->st2 : c->st, s_desc = fork ? c->st3 : c->st;
    PPG_CUDA(c, edge_tail_tc_launch(c->e2, c->we3, c->be3, c->we1, c->be1b, c->heat_raw, n, c->H / 2, c->W / 2, s_heat));
    c->launches++;
    stage_mark(c, "edge_tail");
    const PostMark mk = post_mark(c);  // profiling runs are single-stream: every mark lands on c->st
    PPG_CUDA(c, post_keypoints_launch(p, c->st, &c->launches, mk));
    PPG_CUDA(c, post_heat_launch(p, s_heat, &c->launches, mk));
    if (fork) {
        PPG_CUDA(c, cudaEventRecord(c->ev_heat, c->st2));
        PPG_CUDA(c, cudaEventRecord(c->ev_kp, c->st));
        PPG_CUDA(c, cudaStreamWaitEvent(c->st, c->ev_heat, 0));   // the graph needs keypoints + heat map
        PPG_CUDA(c, cudaStreamWaitEvent(c->st3, c->ev_kp, 0));    // the sampling needs keypoints + descriptor map
    }
    PPG_CUDA(c, post_lines_launch(p, c->st, &c->launches, mk));
    PPG_CUDA(c, post_desc_launch(p, s_desc, &c->launches, mk));
    if (fork) {
        PPG_CUDA(c, cudaEventRecord(c->ev_desc, c->st3));
        PPG_CUDA(c, cudaStreamWaitEvent(c->st, c->ev_desc, 0));   // join: everything is ordered on st again
    }
    return PPG_OK;
}

int ppg_sync(ppg_ctx* c) {
    if (!c) return PPG_ERR_ARG;
    PPG_CUDA(c, cudaSetDevice(c->dev));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    return PPG_OK;
}

int ppg_download(ppg_ctx* c, int n, ppg_frame_out* out) {
    if (!c || !out || n < 1 || n > c->maxB) return set_err(c, PPG_ERR_ARG, "ppg_download: bad arguments");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    const OutLayout& L = c->post.lay;
    // small records first (they say how many descriptor rows are live), one strided copy for all frames; then
    // only the live descriptor rows
    PPG_CUDA(c, cudaMemcpy2DAsync(c->h_out, L.total, c->d_out, L.total, L.small_total, n, cudaMemcpyDeviceToHost,
                                  c->st));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    for (int f = 0; f < n; f++) {
        const int nk = reinterpret_cast<const int*>(c->h_out + (size_t)f * L.total + L.hdr)[HDR_NKP];
        if (nk > 0)
            PPG_CUDA(c, cudaMemcpyAsync(c->h_out + (size_t)f * L.total + L.desc, c->d_out + (size_t)f * L.total + L.desc,
                                        (size_t)nk * 256 * 4, cudaMemcpyDeviceToHost, c->st));
    }
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    int rc = PPG_OK;
    for (int f = 0; f < n; f++) {
        fill_out(c, f, &out[f]);
        if (out[f].status) rc = PPG_ERR_CAPACITY;
    }
    if (rc != PPG_OK) set_err(c, rc, "a per-frame capacity was exceeded (see ppg_frame_out.status)");
    return rc;
}

int ppg_extract_async(ppg_ctx* c, const uint8_t* const* gray, const int* stride, int n) {
    int rc = ppg_upload_frames(c, gray, stride, n);
    if (rc != PPG_OK) return rc;
    if ((rc = ppg_run(c, n)) != PPG_OK) return rc;
    // the whole record of every frame (descriptor area included) in one strided copy: the live descriptor rows are only
    // known once the header is on the host, and a host round trip in the middle is what this entry point avoids
    const OutLayout& L = c->post.lay;
    PPG_CUDA(c, cudaMemcpy2DAsync(c->h_out, L.total, c->d_out, L.total, L.total, n, cudaMemcpyDeviceToHost, c->st));
    return PPG_OK;
}

int ppg_extract_wait(ppg_ctx* c, int n, ppg_frame_out* out) {
    if (!c || !out || n < 1 || n > c->maxB) return set_err(c, PPG_ERR_ARG, "ppg_extract_wait: bad arguments");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    int rc = PPG_OK;
    for (int f = 0; f < n; f++) {
        fill_out(c, f, &out[f]);
        if (out[f].status) rc = PPG_ERR_CAPACITY;
    }
    if (rc != PPG_OK) set_err(c, rc, "a per-frame capacity was exceeded (see ppg_frame_out.status)");
    return rc;
}

long long ppg_record_bytes(const ppg_ctx* c) { return c ? (long long)c->post.lay.total : 0; }

void* ppg_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void ppg_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
int ppg_host_register(void* p, size_t bytes) {
    if (!p || cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        return PPG_ERR_CUDA;
    }
    return PPG_OK;
}
int ppg_host_unregister(void* p) {
    if (!p || cudaHostUnregister(p) != cudaSuccess) {
        cudaGetLastError();
        return PPG_ERR_CUDA;
    }
    return PPG_OK;
}

int ppg_extract(ppg_ctx* c, const uint8_t* const* gray, const int* stride, int n, ppg_frame_out* out) {
    int rc = ppg_upload_frames(c, gray, stride, n);
    if (rc != PPG_OK) return rc;
    if ((rc = ppg_run(c, n)) != PPG_OK) return rc;
    return ppg_download(c, n, out);
}

int ppg_extract_from_maps(ppg_ctx* c, const float* prob, const float* heat, const float* desc_chw, int n,
                          ppg_frame_out* out) {
    if (!c || !prob || !heat || !desc_chw || !out || n < 1 || n > c->maxB)
        return set_err(c, PPG_ERR_ARG, "ppg_extract_from_maps: bad arguments");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    const size_t HW = (size_t)c->H * c->W, hw = (size_t)c->Hc * c->Wc;
    if (!c->prob_in) {
        PPG_CUDA(c, dalloc(&c->prob_in, c->maxB * HW));
        PPG_CUDA(c, dalloc(&c->heat_in, c->maxB * HW));
        PPG_CUDA(c, dalloc(&c->desc_in, c->maxB * hw * 256));
    }
    std::vector<float> nhwc((size_t)n * hw * 256);  // CHW (LibTorch) -> HWC (what the sampling kernel reads)
    for (int f = 0; f < n; f++)
        for (int ch = 0; ch < 256; ch++) {
            const float* src = desc_chw + ((size_t)f * 256 + ch) * hw;
            float* dst = nhwc.data() + (size_t)f * hw * 256 + ch;
            for (size_t i = 0; i < hw; i++) dst[i * 256] = src[i];
        }
    PPG_CUDA(c, cudaMemcpyAsync(c->prob_in, prob, n * HW * 4, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(c->heat_in, heat, n * HW * 4, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(c->desc_in, nhwc.data(), nhwc.size() * 4, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    c->maps_from_caller = true;
    c->last_batch = n;
    c->n_ev = 0;
    stage_mark(c, "start");
    int rc = run_post(c, n);
    if (rc != PPG_OK) return rc;
    return ppg_download(c, n, out);
}

int ppg_get_maps(ppg_ctx* c, int frame, float* prob, float* heat_raw, float* heat_final, float* desc_chw,
                 float* feature_chw) {
    if (!c || frame < 0 || frame >= c->maxB) return set_err(c, PPG_ERR_ARG, "ppg_get_maps: bad frame");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    const size_t HW = (size_t)c->H * c->W, hw = (size_t)c->Hc * c->Wc;
    const bool ext = c->maps_from_caller;
    if (prob) PPG_CUDA(c, cudaMemcpy(prob, (ext ? c->prob_in : c->prob) + frame * HW, HW * 4, cudaMemcpyDeviceToHost));
    if (heat_raw)
        PPG_CUDA(c, cudaMemcpy(heat_raw, (ext ? c->heat_in : c->heat_raw) + frame * HW, HW * 4, cudaMemcpyDeviceToHost));
    if (heat_final) PPG_CUDA(c, cudaMemcpy(heat_final, c->heat_final + frame * HW, HW * 4, cudaMemcpyDeviceToHost));
    if (desc_chw) {
        std::vector<float> t(hw * 256);
        PPG_CUDA(c, cudaMemcpy(t.data(), (ext ? c->desc_in : c->desc) + frame * hw * 256, hw * 256 * 4,
                               cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < hw; i++)
            for (int ch = 0; ch < 256; ch++) desc_chw[(size_t)ch * hw + i] = t[i * 256 + ch];
    }
    if (feature_chw) {
        std::vector<__half> t(hw * 128);
        PPG_CUDA(c, cudaMemcpy(t.data(), c->feat + frame * hw * 128, hw * 128 * 2, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < hw; i++)
            for (int ch = 0; ch < 128; ch++) feature_chw[(size_t)ch * hw + i] = __half2float(t[i * 128 + ch]);
    }
    return PPG_OK;
}

int ppg_get_layer_output(ppg_ctx* c, const char* name, int frame, void* dst, size_t max_bytes, size_t* bytes) {
    if (!c || !name || !bytes || frame < 0 || frame >= c->maxB) return set_err(c, PPG_ERR_ARG, "ppg_get_layer_output: bad argument");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    for (auto& l : c->tc) {
        if (strcmp(l.name, name)) continue;
        size_t px = (size_t)l.H * l.W, el = 2;
        if (l.mode == EPI_F16_POOL) px /= 4;
        if (l.mode == EPI_F16_PS2) px *= 4;
        if (l.mode == EPI_F32) el = 4;
        if (l.mode == EPI_SOFTMAX_D2S) px *= 64, el = 4;
        const size_t sz = px * (l.mode == EPI_SOFTMAX_D2S ? 1 : l.out_ld) * el;
        *bytes = sz;
        const size_t n = sz < max_bytes ? sz : max_bytes;
        if (dst && n)
            PPG_CUDA(c, cudaMemcpy(dst, static_cast<const uint8_t*>(l.out) + (size_t)frame * sz, n, cudaMemcpyDeviceToHost));
        return PPG_OK;
    }
    return set_err(c, PPG_ERR_ARG, "ppg_get_layer_output: no such layer");
}

// Validation of the tcgen05 convolution against a plain CUDA-core convolution over the same fp16 operands.
int ppg_selftest_conv(ppg_ctx* c, int max_layers, const char** names, float* max_abs_diff, float* max_abs_ref,
                      int* n_layers) {
    if (!c || !names || !max_abs_diff || !max_abs_ref || !n_layers) return set_err(c, PPG_ERR_ARG, "null argument");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    if (!c->tc.empty() && c->tc[0].L.v3 == 2) {
        // conv1b computed conv1a on the fly: give the check its input map from the stand-alone conv1a kernel (frame 0)
        PPG_CUDA(c, conv1a_tc_launch(c->gray, c->w1a, c->b1a, c->a1, 1, c->H, c->W, c->st));
        c->launches++;
    }
    int nl = 0;
    for (auto& l : c->tc) {
        if (nl >= max_layers) break;
        const size_t npix = (size_t)l.H * l.W;
        float* ref = nullptr;
        PPG_CUDA(c, dalloc(&ref, npix * l.N));
        PPG_CUDA(c, conv_ref_launch(l.in, l.w, l.bias, ref, 1, l.H, l.W, l.cin, l.N, l.taps, l.relu, c->st));
        c->launches++;
        std::vector<float> r(npix * l.N);
        PPG_CUDA(c, cudaMemcpyAsync(r.data(), ref, r.size() * 4, cudaMemcpyDeviceToHost, c->st));
        PPG_CUDA(c, cudaStreamSynchronize(c->st));
        cudaFree(ref);
        float md = 0.f, mr = 0.f;
        auto acc = [&](float got, float want) {
            const float d = fabsf(got - want);
            if (!(d <= md)) md = d;  // NaN-propagating max
            if (fabsf(want) > mr) mr = fabsf(want);
        };
        if (l.mode == EPI_SOFTMAX_D2S) {  // compare the probabilities with the softmax of the reference logits
            const int Wf = l.W * 8;
            std::vector<float> o((size_t)l.H * 8 * Wf);
            PPG_CUDA(c, cudaMemcpy(o.data(), l.out, o.size() * 4, cudaMemcpyDeviceToHost));
            for (int y = 0; y < l.H; y++)
                for (int x = 0; x < l.W; x++) {
                    const float* lg = &r[((size_t)y * l.W + x) * l.N];
                    float m = lg[0];
                    for (int ch = 1; ch < 65; ch++) m = fmaxf(m, lg[ch]);
                    float sum = 0.f;
                    for (int ch = 0; ch < 65; ch++) sum += expf(lg[ch] - m);
                    for (int ch = 0; ch < 64; ch++)
                        acc(o[((size_t)(8 * y + ch / 8)) * Wf + 8 * x + ch % 8], expf(lg[ch] - m) / sum);
                }
        } else if (l.mode == EPI_F32) {
            std::vector<float> o(npix * l.out_ld);
            PPG_CUDA(c, cudaMemcpy(o.data(), l.out, o.size() * 4, cudaMemcpyDeviceToHost));
            for (size_t px = 0; px < npix; px++)
                for (int ch = 0; ch < l.N; ch++) acc(o[px * l.out_ld + ch], r[px * l.N + ch]);
        } else if (l.mode == EPI_F16) {
            std::vector<__half> o(npix * l.out_ld);
            PPG_CUDA(c, cudaMemcpy(o.data(), l.out, o.size() * 2, cudaMemcpyDeviceToHost));
            for (size_t px = 0; px < npix; px++)
                for (int ch = 0; ch < l.N; ch++) acc(__half2float(o[px * l.out_ld + ch]), r[px * l.N + ch]);
        } else if (l.mode == EPI_F16_POOL) {
            const int Ho = l.H / 2, Wo = l.W / 2;
            std::vector<__half> o((size_t)Ho * Wo * l.out_ld);
            PPG_CUDA(c, cudaMemcpy(o.data(), l.out, o.size() * 2, cudaMemcpyDeviceToHost));
            for (int y = 0; y < Ho; y++)
                for (int x = 0; x < Wo; x++)
                    for (int ch = 0; ch < l.N; ch++) {
                        float m = -INFINITY;
                        for (int i = 0; i < 2; i++)
                            for (int j = 0; j < 2; j++)
                                m = fmaxf(m, r[((size_t)(2 * y + i) * l.W + 2 * x + j) * l.N + ch]);
                        acc(__half2float(o[((size_t)y * Wo + x) * l.out_ld + ch]), m);
                    }
        } else {  // EPI_F16_PS2
            const int Ho = l.H * 2, Wo = l.W * 2;
            std::vector<__half> o((size_t)Ho * Wo * l.out_ld);
            PPG_CUDA(c, cudaMemcpy(o.data(), l.out, o.size() * 2, cudaMemcpyDeviceToHost));
            for (int y = 0; y < l.H; y++)
                for (int x = 0; x < l.W; x++)
                    for (int ch = 0; ch < l.N; ch++) {
                        const int oc = ch >> 2, i = (ch >> 1) & 1, j = ch & 1;
                        acc(__half2float(o[((size_t)(2 * y + i) * Wo + 2 * x + j) * l.out_ld + oc]),
                            r[((size_t)y * l.W + x) * l.N + ch]);
                    }
        }
        names[nl] = l.name;
        max_abs_diff[nl] = md;
        max_abs_ref[nl] = mr;
        nl++;
    }
    *n_layers = nl;
    return PPG_OK;
}

int ppg_set_profiling(ppg_ctx* c, int on) {
    if (!c) return PPG_ERR_ARG;
    c->profiling = on != 0;
    c->n_ev = 0;
    return PPG_OK;
}

int ppg_get_stage_times(ppg_ctx* c, int max_stages, const char** names, float* ms, int* n_stages) {
    if (!c || !names || !ms || !n_stages) return set_err(c, PPG_ERR_ARG, "null argument");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    int n = 0;
    for (int i = 1; i < c->n_ev && n < max_stages; i++, n++) {
        names[n] = c->ev_names[i];
        PPG_CUDA(c, cudaEventElapsedTime(&ms[n], c->ev[i - 1], c->ev[i]));
    }
    *n_stages = n;
    return PPG_OK;
}

long long ppg_launch_count(const ppg_ctx* c) { return c ? c->launches : 0; }

int ppg_timer_start(ppg_ctx* c) {
    if (!c) return PPG_ERR_ARG;
    PPG_CUDA(c, cudaSetDevice(c->dev));
    PPG_CUDA(c, cudaEventRecord(c->t0, c->st));
    return PPG_OK;
}
int ppg_timer_stop(ppg_ctx* c, float* ms) {
    if (!c || !ms) return PPG_ERR_ARG;
    PPG_CUDA(c, cudaSetDevice(c->dev));
    PPG_CUDA(c, cudaEventRecord(c->t1, c->st));
    PPG_CUDA(c, cudaEventSynchronize(c->t1));
    PPG_CUDA(c, cudaEventElapsedTime(ms, c->t0, c->t1));
    return PPG_OK;
}

void* ppg_stream(ppg_ctx* c) { return c ? (void*)c->st : nullptr; }

}  // extern "C"
