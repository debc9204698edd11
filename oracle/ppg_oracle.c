/*
 * ppg_oracle.c -- CPU restatement of the PPG-SLAM front-end post-processing and association.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this.  The product (ppg_slam_b200/, include/) never does.
 *
 * Parity status: the reference ships no tests / golden vectors.  What pins this file
 * instead -- first of all the reference's own C++ compiled from /root/reference (oracle/ref_build.py,
 * tests/test_ref_pin.py, tests/test_ref_pin_kb8.py: extractor post-processing, Frame grid, ExtendMapMatches,
 * SearchForInitialization, SearchForTriangulation with both cameras, SearchByBoW x 2, CheckInFrustum), and:
 *   - the four OpenCV routines are checked bit-for-bit against cv2 4.13 (tests/test_oracle_cv.py);
 *   - the descriptor sampler is checked against torch.grid_sampler + F.normalize;
 *   - everything else follows the cited reference lines statement by statement.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC ppg_oracle.c -lm   (oracle/build.py)
 * -ffp-contract=off matters: the reference is built without -march (SSE2, no FMA; CMakeLists.txt:8-9).
 *
 * Two documented divergences from the literal reference (DESIGN.md "Oracle"):
 *   1. std::sort ties (PPGExtractor.cpp:180, Matcher.cpp:221) are broken by a stable order
 *      (raster index / input order); std::sort leaves them unspecified.
 *   2. libm float transcendentals (atan2f :283, sin(double) :330, sinf :413) are evaluated as the
 *      correctly rounded float of the double routine ((float)atan2((double)y,(double)x) ...), which
 *      is what a correctly-rounded libm (glibc >= 2.41) returns; PPGO_LIBM_FLOAT=1 at compile time
 *      switches to the literal float libm calls so tests can show the graph does not depend on it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PPGO_PI 3.1415926535897932384626433832795 /* CV_PI  */
#define PPGO_2PI 6.283185307179586476925286766559 /* CV_2PI */

#ifndef PPGO_LIBM_FLOAT
#define PPGO_LIBM_FLOAT 0
#endif

static float o_atan2f(float y, float x) {
#if PPGO_LIBM_FLOAT
    return atan2f(y, x);
#else
    return (float)atan2((double)y, (double)x);
#endif
}
static float o_sinf(float a) {
#if PPGO_LIBM_FLOAT
    return sinf(a);
#else
    return (float)sin((double)a);
#endif
}
static float o_tanf(float a) { /* std::tan(float), KannalaBrandt8.cpp:87 */
#if PPGO_LIBM_FLOAT
    return tanf(a);
#else
    return (float)tan((double)a);
#endif
}

typedef struct {
    int width, height;
    float K[9]; /* row-major fx 0 cx 0 fy cy 0 0 1 (toK(), Pinhole.cpp / KannalaBrandt8.cpp:136) */
    float D[4]; /* pinhole: k1 k2 p1 p2; KB8: k0..k3 (toD()) */
    int fisheye;
    /* static tunables, PPGExtractor.cpp:44-53 */
    float junction_thresh;   /* 1/128 */
    int junction_nms_radius; /* 4     */
    int junction_max_num;    /* 500   */
    float line_valid_thresh; /* 0.01  */
    float line_valid_ratio;  /* 0.3   */
    float line_dist_thresh;  /* 2.0   */
    int heatmap_refine_sz;   /* 16    */
    float line_heatmap_thresh; /* 0.2 */
    float line_inlier_rate;  /* 0.8   */
} ppgo_cfg;

void ppgo_default_cfg(ppgo_cfg *c) {
    memset(c, 0, sizeof(*c));
    c->junction_thresh = 1.0f / 128.0f;
    c->junction_nms_radius = 4;
    c->junction_max_num = 500;
    c->line_valid_thresh = 1.0e-2f;
    c->line_valid_ratio = 0.3f;
    c->line_dist_thresh = 2.0f;
    c->heatmap_refine_sz = 16;
    c->line_heatmap_thresh = 0.2f;
    c->line_inlier_rate = 0.8f;
}

/* ------------------------------------------------------------------------------------------- */
/* OpenCV routines the extractor calls (source not under /root/reference: OpenCV >= 4.2,       */
/* README.md:26; restated from the published algorithm, pinned against cv2 4.13 in the tests). */
/* ------------------------------------------------------------------------------------------- */

/* cv::undistortPoints(pts,K,D,noArray(),K) with D=(k1,k2,p1,p2): PPGExtractor.cpp:223 and
 * GeometricCamera.cpp:40.  5 fixed-point iterations in double, output cast to float. */
void ppgo_undistort_points_pinhole(const float *K, const float *D, const float *xy, int n, float *out) {
    double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    double ifx = 1. / fx, ify = 1. / fy;
    double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3];
    for (int i = 0; i < n; i++) {
        double u = xy[2 * i], v = xy[2 * i + 1];
        double x = (u - cx) * ifx, y = (v - cy) * ify;
        double x0 = x, y0 = y;
        for (int it = 0; it < 5; it++) {
            double r2 = x * x + y * y;
            double icdist = (1 + ((0. * r2 + 0.) * r2 + 0.) * r2) / (1 + ((0. * r2 + k2) * r2 + k1) * r2);
            if (icdist < 0) {
                x = (u - cx) * ifx;
                y = (v - cy) * ify;
                break;
            }
            double dX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x) + 0. * r2 + 0. * r2 * r2;
            double dY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y + 0. * r2 + 0. * r2 * r2;
            x = (x0 - dX) * icdist;
            y = (y0 - dY) * icdist;
        }
        /* P = K, R = I */
        double xx = fx * x + 0. * y + cx, yy = 0. * x + fy * y + cy, ww = 1. / (0. * x + 0. * y + 1.);
        out[2 * i] = (float)(xx * ww);
        out[2 * i + 1] = (float)(yy * ww);
    }
}

/* cv::fisheye::undistortPoints(pts,pts,K,D,Mat(),K): PPGExtractor.cpp:221.  Newton, <=10 its, eps 1e-8. */
void ppgo_undistort_points_fisheye(const float *K, const float *D, const float *xy, int n, float *out) {
    double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    double k[4] = {D[0], D[1], D[2], D[3]};
    for (int i = 0; i < n; i++) {
        double px = xy[2 * i], py = xy[2 * i + 1];
        double pwx = (px - cx) / fx, pwy = (py - cy) / fy;
        double theta_d = sqrt(pwx * pwx + pwy * pwy);
        theta_d = fmin(fmax(-PPGO_PI / 2., theta_d), PPGO_PI / 2.);
        int converged = 0;
        double theta = theta_d, scale = 0.0;
        if (fabs(theta_d) > 1e-8) {
            for (int j = 0; j < 10; j++) {
                double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
                double k0t2 = k[0] * t2, k1t4 = k[1] * t4, k2t6 = k[2] * t6, k3t8 = k[3] * t8;
                double fix = (theta * (1 + k0t2 + k1t4 + k2t6 + k3t8) - theta_d) /
                             (1 + 3 * k0t2 + 5 * k1t4 + 7 * k2t6 + 9 * k3t8);
                theta = theta - fix;
                if (fabs(fix) < 1e-8) {
                    converged = 1;
                    break;
                }
            }
            scale = tan(theta) / theta_d;
        } else {
            converged = 1;
        }
        int flipped = ((theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0));
        if (converged && !flipped) {
            double pux = pwx * scale, puy = pwy * scale;
            double pr0 = fx * pux + 0. * puy + cx * 1.0, pr1 = 0. * pux + fy * puy + cy * 1.0;
            double pr2 = 0. * pux + 0. * puy + 1.0;
            out[2 * i] = (float)(pr0 / pr2);
            out[2 * i + 1] = (float)(pr1 / pr2);
        } else {
            out[2 * i] = -1000000.0f;
            out[2 * i + 1] = -1000000.0f;
        }
    }
}

/* cv::initUndistortRectifyMap(K,D,I,K,size,CV_32F,mX,mY): PPGExtractor.cpp:69 (pinhole). */
void ppgo_init_undistort_map_pinhole(const float *K, const float *D, int W, int H, float *mx, float *my) {
    double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3];
    for (int v = 0; v < H; v++)
        for (int u = 0; u < W; u++) {
            double x = ((double)u - cx) / fx, y = ((double)v - cy) / fy;
            double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
            double kr = 1 + ((0. * r2 + k2) * r2 + k1) * r2;
            double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2);
            double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy;
            mx[v * W + u] = (float)(fx * xd + cx);
            my[v * W + u] = (float)(fy * yd + cy);
        }
}

/* cv::fisheye::initUndistortRectifyMap(K,D,I,K,size,CV_32F,mX,mY): PPGExtractor.cpp:66.
 * (Computed by the reference but unused for every shipped config because D[0]==0, :261.) */
void ppgo_init_undistort_map_fisheye(const float *K, const float *D, int W, int H, float *mx, float *my) {
    double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    double k[4] = {D[0], D[1], D[2], D[3]};
    for (int v = 0; v < H; v++)
        for (int u = 0; u < W; u++) {
            double x = ((double)u - cx) / fx, y = ((double)v - cy) / fy;
            double r = sqrt(x * x + y * y);
            double theta = atan(r);
            double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
            double theta_d = theta * (1 + k[0] * t2 + k[1] * t4 + k[2] * t6 + k[3] * t8);
            double scale = (r == 0) ? 1.0 : theta_d / r;
            mx[v * W + u] = (float)(fx * x * scale + cx);
            my[v * W + u] = (float)(fy * y * scale + cy);
        }
}

/* cv::remap(src,dst,mX,mY,INTER_LINEAR) with BORDER_CONSTANT 0: PPGExtractor.cpp:262.
 * Fixed-point (1/32) bilinear exactly as cv2 4.13 evaluates it for CV_32F. */
void ppgo_remap_linear(const float *src, int W, int H, const float *mx, const float *my, float *dst) {
    for (int i = 0; i < W * H; i++) {
        int sx = (int)lrint((double)(mx[i] * 32.0f)), sy = (int)lrint((double)(my[i] * 32.0f));
        int ix = sx >> 5, iy = sy >> 5;
        float fx = (float)(sx & 31) * (1.0f / 32.0f), fy = (float)(sy & 31) * (1.0f / 32.0f);
        float w00 = (1.0f - fy) * (1.0f - fx), w01 = (1.0f - fy) * fx, w10 = fy * (1.0f - fx), w11 = fy * fx;
        float s00 = 0, s01 = 0, s10 = 0, s11 = 0;
        if (iy >= 0 && iy < H) {
            if (ix >= 0 && ix < W) s00 = src[iy * W + ix];
            if (ix + 1 >= 0 && ix + 1 < W) s01 = src[iy * W + ix + 1];
        }
        if (iy + 1 >= 0 && iy + 1 < H) {
            if (ix >= 0 && ix < W) s10 = src[(iy + 1) * W + ix];
            if (ix + 1 >= 0 && ix + 1 < W) s11 = src[(iy + 1) * W + ix + 1];
        }
        dst[i] = ((s00 * w00 + s01 * w01) + s10 * w10) + s11 * w11;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* PPGExtractor::detectKeyPoint, feature/src/PPGExtractor.cpp:158-234 (from the H x W prob map) */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    float score;
    int idx;
} cand_t;
static int cand_cmp(const void *a, const void *b) {
    const cand_t *x = a, *y = b;
    if (x->score > y->score) return -1; /* :180 descending by score */
    if (x->score < y->score) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx); /* divergence 1: ties by raster index */
}

/* -> N.  x,y integer pixel, score, (xun,yun) = mPosUn, out = mbOut.  n_cand receives the number of
 * pixels >= threshold (diagnostic). */
int ppgo_detect_keypoints(const ppgo_cfg *c, const float *prob, int *kx, int *ky, float *kscore, float *kxun,
                          float *kyun, uint8_t *kout, int *n_cand) {
    int W = c->width, H = c->height, R = c->junction_nms_radius;
    cand_t *cand = malloc(sizeof(cand_t) * (size_t)W * H);
    int nc = 0;
    for (int i = 0; i < H; i++) /* :168-177 */
        for (int j = 0; j < W; j++) {
            float s = prob[i * W + j];
            if (s < c->junction_thresh) continue;
            cand[nc].score = s;
            cand[nc].idx = i * W + j;
            nc++;
        }
    if (n_cand) *n_cand = nc;
    qsort(cand, nc, sizeof(cand_t), cand_cmp);
    unsigned char *flag = calloc((size_t)W * H, 1); /* :181 */
    int n = 0;
    for (int t = 0; t < nc; t++) { /* :185-206 */
        int px = cand[t].idx % W, py = cand[t].idx / W;
        if (px < R || px > (W - R - 1) || py < R || py > (H - R - 1) || flag[py * W + px] != 0) continue;
        flag[py * W + px] = 1;
        kx[n] = px;
        ky[n] = py;
        kscore[n] = cand[t].score;
        n++;
        if ((unsigned)n + 1 > (unsigned)c->junction_max_num) break;
        for (int i = py - R; i <= py + R; i++)
            for (int j = px - R; j <= px + R; j++) {
                if (i < 0 || i > H || j < 0 || j > W) continue;
                flag[i * W + j] = (unsigned char)-1;
            }
    }
    free(flag);
    free(cand);
    if (n == 0) return 0; /* :210 */
    float *pts = malloc(sizeof(float) * 2 * n), *und = malloc(sizeof(float) * 2 * n);
    for (int i = 0; i < n; i++) {
        pts[2 * i] = (float)kx[i];
        pts[2 * i + 1] = (float)ky[i];
    }
    if (c->fisheye)
        ppgo_undistort_points_fisheye(c->K, c->D, pts, n, und);
    else
        ppgo_undistort_points_pinhole(c->K, c->D, pts, n, und);
    for (int i = 0; i < n; i++) { /* :226-233 */
        float u = und[2 * i], v = und[2 * i + 1];
        kout[i] = 1; /* KeyPointEx ctor: mbOut(true) */
        if (u >= 1 && u < W - 1 && v >= 1 && v < H - 1) kout[i] = 0;
        kxun[i] = u;
        kyun[i] = v;
    }
    free(pts);
    free(und);
    return n;
}

/* ------------------------------------------------------------------------------------------- */
/* PPGExtractor::refineHeatMap :540-578 applied tile by tile as in detectLines :243-256          */
/* ------------------------------------------------------------------------------------------- */
static int fdesc_cmp(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x < y) - (x > y);
}
static void refine_tile(const ppgo_cfg *c, float *s, int stride, int segH, int segW) {
    float *v = malloc(sizeof(float) * (size_t)segH * segW);
    size_t n = 0;
    for (int i = 0; i < segH; i++)
        for (int j = 0; j < segW; j++)
            if (s[i * stride + j] > c->line_valid_thresh) v[n++] = s[i * stride + j];
    int valCount = (int)(c->line_valid_ratio * (float)n); /* :554 float * size_t */
    if (valCount < 1) {
        free(v);
        return;
    }
    if ((double)n >= (double)(segH * segW) * 0.9 && (double)v[(size_t)((double)n * 0.9)] > 0.1) { /* :557 */
        for (int i = 0; i < segH; i++)
            for (int j = 0; j < segW; j++) s[i * stride + j] = 0.0f;
        free(v);
        return;
    }
    qsort(v, n, sizeof(float), fdesc_cmp); /* :562 descending */
    double acc = 0.0;
    for (int i = 0; i < valCount; i++) acc += (double)v[i];       /* std::accumulate(...,0.0) */
    float ave = (float)(acc / (double)(float)valCount);            /* :563 */
    for (int i = 0; i < segH; i++)
        for (int j = 0; j < segW; j++) {
            float cur = s[i * stride + j];
            if (cur > c->line_valid_thresh) {
                float ns = cur / ave;
                s[i * stride + j] = ((double)ns > 1.0) ? 1.0f : ns;
            } else
                s[i * stride + j] = 0.0f;
        }
    free(v);
}

void ppgo_refine_heat(const ppgo_cfg *c, float *heat) {
    int W = c->width, H = c->height, S = c->heatmap_refine_sz;
    int gx = W / S, gy = H / S;
    for (int i = 0; i < gy; i++)
        for (int j = 0; j < gx; j++) {
            int h0 = i * S, w0 = j * S;
            int h1 = (i == gy - 1) ? H : (i + 1) * S, w1 = (j == gx - 1) ? W : (j + 1) * S;
            refine_tile(c, heat + h0 * W + w0, W, h1 - h0, w1 - w0);
        }
}

/* ------------------------------------------------------------------------------------------- */
/* PPGExtractor::detectLines :259-441 from the refined (and, for pinhole, remapped) heat map     */
/* ------------------------------------------------------------------------------------------- */
static const float invSampleGapTable[4] = {0.3333, 0.200, 0.1427, 0.1111}; /* :19 */

typedef struct {
    int s, e;
    int bad;
    float lscore;
} line_t;
typedef struct {
    int *v;
    int n, cap;
} ivec;
static void ivec_push(ivec *a, int x) {
    if (a->n == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 8;
        a->v = realloc(a->v, sizeof(int) * a->cap);
    }
    a->v[a->n++] = x;
}

static float bilinear(const float *M, int W, float ptX, float ptY) { /* :580-589 */
    int x1 = (int)ptX, x2 = x1 + 1, y1 = (int)ptY, y2 = y1 + 1;
    float data1 = ((float)x2 - ptX) * M[y1 * W + x1] + (ptX - (float)x1) * M[y1 * W + x2];
    float data2 = ((float)x2 - ptX) * M[y2 * W + x1] + (ptX - (float)x1) * M[y2 * W + x2];
    return ((float)y2 - ptY) * data1 + (ptY - (float)y1) * data2;
}

static void seg_params(float dist, float invScale, int *segNum, float *step) {
    int lenLevel = (int)((double)(dist * invScale) * 4.0);       /* :485 */
    *segNum = (int)(dist * invSampleGapTable[lenLevel]);          /* :486 */
    *step = (float)(1.0 / (double)(float)(*segNum));              /* :487 */
}

static float inlier_rate(const ppgo_cfg *c, const float *heat, float invScale, float psx, float psy, float pex,
                         float pey) { /* :461-498 */
    int W = c->width;
    float dx = psx - pex, dy = psy - pey;
    float dist = sqrtf(dx * dx + dy * dy);
    int segNum;
    float step;
    seg_params(dist, invScale, &segNum, &step);
    int cnt = 0;
    for (int i = 1; i < segNum; i++) {
        float sx = (psx * step) * (float)i + (pex * step) * (float)(segNum - i);
        float sy = (psy * step) * (float)i + (pey * step) * (float)(segNum - i);
        int posx = (int)((double)sx + 0.5), posy = (int)((double)sy + 0.5);
        if (heat[posy * W + posx] > c->line_heatmap_thresh) cnt++;
    }
    return (float)cnt / (float)(segNum - 1);
}

static float line_score(const ppgo_cfg *c, const float *heat, float invScale, float psx, float psy, float pex,
                        float pey) { /* :500-513 */
    int W = c->width;
    float dx = psx - pex, dy = psy - pey;
    float dist = sqrtf(dx * dx + dy * dy);
    int segNum;
    float step;
    seg_params(dist, invScale, &segNum, &step);
    float sum = 0.f;
    for (int i = 1; i < segNum; i++) {
        float sx = (psx * step) * (float)i + (pex * step) * (float)(segNum - i);
        float sy = (psy * step) * (float)i + (pey * step) * (float)(segNum - i);
        sum += bilinear(heat, W, sx, sy);
    }
    return sum / (float)(segNum - 1);
}

/* dirMat/distMat entries are pure functions of the two endpoints (:268-288); evaluate on demand. */
static float dist_of(const float *xun, const float *yun, int a, int b) {
    int i = a < b ? a : b, j = a < b ? b : a;
    float dx = xun[j] - xun[i], dy = yun[j] - yun[i];
    return sqrtf(dx * dx + dy * dy);
}
static float dir_of(const float *xun, const float *yun, int a, int b) {
    int i = a < b ? a : b, j = a < b ? b : a;
    float dx = xun[j] - xun[i], dy = yun[j] - yun[i];
    float dist = sqrtf(dx * dx + dy * dy);
    float d = o_atan2f(dy / dist, dx / dist); /* :280-283 */
    if (a < b) return d;
    float r = (float)((double)d - PPGO_PI);   /* :284 */
    if ((double)r < -PPGO_PI) r = (float)((double)r + PPGO_2PI); /* :285-286 */
    return r;
}

/* one adjacency scan of the overlap filter, :316-335 / :338-357.  p = shared endpoint, q = new other. */
static int overlap_scan(const ppgo_cfg *c, const float *xun, const float *yun, line_t *lines, const ivec *adj, int p,
                        int q) {
    int isOverlap = 0;
    for (int t = 0; t < adj->n; t++) {
        line_t *lo = &lines[adj->v[t]];
        if (lo->bad) continue;
        int pid_old = (p == lo->s) ? lo->e : lo->s;
        float a = dir_of(xun, yun, p, q) - dir_of(xun, yun, p, pid_old);
        if ((double)a < -PPGO_PI) a = (float)((double)a + PPGO_2PI);
        if ((double)a > PPGO_PI) a = (float)((double)a - PPGO_2PI);
        a = fabsf(a);
        if ((double)a > 0.2 * PPGO_PI) continue;
        float distNew = dist_of(xun, yun, p, q), distOld = dist_of(xun, yun, p, pid_old);
#if PPGO_LIBM_FLOAT == 2
        float s = sinf(a);
#else
        float s = (float)sin((double)a); /* :330 unqualified sin -> double overload */
#endif
        if (distNew <= distOld && distNew * s < c->line_dist_thresh) lo->bad = 1;
        if (distOld < distNew && distOld * s < c->line_dist_thresh) isOverlap = 1;
    }
    return isOverlap;
}

/* Outputs (caller-allocated, capacities in brackets):
 *   edge_s/edge_e/edge_score [max_edges]  final mvKeyEdges (:433-441), returns E (or -1 if > max_edges)
 *   conn_off [n+1], conn_idx [2*max_edges]     CSR of KeyPointEx::mvConnected (final edge indices)
 *   col_off [n+1], col_pairs [2*max_col]       CSR of KeyPointEx::mvColine (p1,p2 pairs), *n_col total pairs
 *   stats[0]=pairs passing the 3-point test, stats[1]=candidateLines.size(), stats[2]=bad after filter */
int ppgo_detect_lines(const ppgo_cfg *c, const float *heat, int n, const float *xun, const float *yun,
                      const uint8_t *kout, int max_edges, int *edge_s, int *edge_e, float *edge_score, int *conn_off,
                      int *conn_idx, int max_col, int *col_off, int *col_pairs, int *n_col, int *stats) {
    int W = c->width, H = c->height;
    float invScale = 1.0f / sqrtf((float)(H * H + W * W)); /* :74 */
    line_t *lines = NULL;
    int nl = 0, capl = 0;
    ivec *adj = calloc(n > 0 ? n : 1, sizeof(ivec));
    int npass = 0;
    for (int i = 0; i < n; i++) { /* :293-365 */
        if (kout[i]) continue;
        for (int j = i + 1; j < n; j++) {
            if (kout[j]) continue;
            float c1x = xun[j] * 0.2f + xun[i] * 0.8f, c1y = yun[j] * 0.2f + yun[i] * 0.8f;
            float c2x = xun[j] * 0.8f + xun[i] * 0.2f, c2y = yun[j] * 0.8f + yun[i] * 0.2f;
            float c3x = xun[j] * 0.5f + xun[i] * 0.5f, c3y = yun[j] * 0.5f + yun[i] * 0.5f;
            if (heat[(int)((double)c1y + 0.5) * W + (int)((double)c1x + 0.5)] < c->line_heatmap_thresh) continue;
            if (heat[(int)((double)c2y + 0.5) * W + (int)((double)c2x + 0.5)] < c->line_heatmap_thresh) continue;
            if (heat[(int)((double)c3y + 0.5) * W + (int)((double)c3x + 0.5)] < c->line_heatmap_thresh) continue;
            npass++;
            if (overlap_scan(c, xun, yun, lines, &adj[i], i, j)) continue;
            if (overlap_scan(c, xun, yun, lines, &adj[j], j, i)) continue;
            if (nl == capl) {
                capl = capl ? capl * 2 : 256;
                lines = realloc(lines, sizeof(line_t) * capl);
            }
            lines[nl].s = i;
            lines[nl].e = j;
            lines[nl].bad = 0;
            lines[nl].lscore = 0.f;
            ivec_push(&adj[i], nl);
            ivec_push(&adj[j], nl);
            nl++;
        }
    }
    if (stats) {
        stats[0] = npass;
        stats[1] = nl;
        int nb = 0;
        for (int i = 0; i < nl; i++) nb += lines[i].bad;
        stats[2] = nb;
    }
    for (int i = 0; i < n; i++) adj[i].n = 0; /* :366 */
    for (int i = 0; i < nl; i++) {            /* :367-389 */
        line_t *kl = &lines[i];
        if (kl->bad) continue;
        float psx = xun[kl->s], psy = yun[kl->s], pex = xun[kl->e], pey = yun[kl->e];
        float si = inlier_rate(c, heat, invScale, psx, psy, pex, pey);
        if (si < c->line_inlier_rate) {
            kl->bad = 1;
            continue;
        }
        float sh = line_score(c, heat, invScale, psx, psy, pex, pey);
        if (sh < c->line_heatmap_thresh) {
            kl->bad = 1;
            continue;
        }
        kl->lscore = si * sh;
        ivec_push(&adj[kl->s], i);
        ivec_push(&adj[kl->e], i);
    }
    /* colinearity :392-432 */
    int ncol = 0, col_overflow = 0;
    for (int p = 0; p < n; p++) {
        col_off[p] = ncol;
        int m = adj[p].n;
        int *ti = malloc(sizeof(int) * (m > 0 ? m : 1));
        memcpy(ti, adj[p].v, sizeof(int) * m);
        while (m > 1) {
            double minPD = 1e9;
            int bp1 = -1, bp2 = -1, bestId = -1;
            line_t *kl1 = &lines[ti[m - 1]];
            if (kl1->bad) {
                m--;
                continue;
            }
            for (int i = 0; i < m - 1; i++) {
                line_t *kl2 = &lines[ti[i]];
                if (kl2->bad) continue;
                int p1 = (p == kl1->s) ? kl1->e : kl1->s, p2 = (p == kl2->s) ? kl2->e : kl2->s;
                float ad = dir_of(xun, yun, p, p1) - dir_of(xun, yun, p, p2);
                double pd = 0.5 * (double)(dist_of(xun, yun, p, p1) + dist_of(xun, yun, p, p2)) *
                            (double)fabsf(o_sinf(ad)); /* :413 */
                if (minPD > pd) {
                    minPD = pd;
                    bestId = i;
                    bp1 = p1;
                    bp2 = p2;
                }
            }
            if (minPD > (double)c->line_dist_thresh) {
                m--;
                continue;
            }
            if (ncol < max_col) {
                col_pairs[2 * ncol] = bp1;
                col_pairs[2 * ncol + 1] = bp2;
            } else
                col_overflow = 1;
            ncol++;
            m--;
            ti[bestId] = ti[m - 1];
            m--;
        }
        free(ti);
    }
    col_off[n] = ncol;
    *n_col = col_overflow ? -1 : ncol;
    /* final edges + mvConnected :433-441 */
    int E = 0, overflow = 0;
    ivec *fin = calloc(n > 0 ? n : 1, sizeof(ivec));
    for (int i = 0; i < nl; i++) {
        if (lines[i].bad) continue;
        if (E < max_edges) {
            edge_s[E] = lines[i].s;
            edge_e[E] = lines[i].e;
            edge_score[E] = lines[i].lscore;
        } else
            overflow = 1;
        ivec_push(&fin[lines[i].s], E);
        ivec_push(&fin[lines[i].e], E);
        E++;
    }
    int k = 0;
    for (int p = 0; p < n; p++) {
        conn_off[p] = k;
        for (int t = 0; t < fin[p].n; t++) {
            if (!overflow) conn_idx[k] = fin[p].v[t];
            k++;
        }
        free(fin[p].v);
        free(adj[p].v);
    }
    conn_off[n] = k;
    free(fin);
    free(adj);
    free(lines);
    return overflow ? -1 : E;
}

/* ------------------------------------------------------------------------------------------- */
/* PPGExtractor::genPointDescriptor :515-538.  desc is the raw dense map, CHW (256,Hc,Wc) fp32.  */
/* grid_sampler(bilinear, zeros, align_corners=false) + F.normalize(dim=1) restated from ATen.    */
/* ------------------------------------------------------------------------------------------- */
void ppgo_sample_descriptors(const ppgo_cfg *c, const float *desc, int C, int Hc, int Wc, int n, const int *kx,
                             const int *ky, float *out) {
    if (n < 10) { /* :520-524 */
        memset(out, 0, sizeof(float) * (size_t)n * C);
        return;
    }
    for (int i = 0; i < n; i++) {
        float gx = (float)((double)((float)kx[i] / (float)c->width) * 2. - 1.);  /* :530 */
        float gy = (float)((double)((float)ky[i] / (float)c->height) * 2. - 1.); /* :531 */
        float ix = ((gx + 1.f) * (float)Wc - 1.f) / 2.f, iy = ((gy + 1.f) * (float)Hc - 1.f) / 2.f;
        float x0f = floorf(ix), y0f = floorf(iy);
        int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
        float nw = ((float)x1 - ix) * ((float)y1 - iy), ne = (ix - (float)x0) * ((float)y1 - iy);
        float sw = ((float)x1 - ix) * (iy - (float)y0), se = (ix - (float)x0) * (iy - (float)y0);
        double ss = 0.0;
        float *o = out + (size_t)i * C;
        for (int ch = 0; ch < C; ch++) {
            const float *p = desc + (size_t)ch * Hc * Wc;
            float r = 0.f;
            if (y0 >= 0 && y0 < Hc && x0 >= 0 && x0 < Wc) r += p[y0 * Wc + x0] * nw;
            if (y0 >= 0 && y0 < Hc && x1 >= 0 && x1 < Wc) r += p[y0 * Wc + x1] * ne;
            if (y1 >= 0 && y1 < Hc && x0 >= 0 && x0 < Wc) r += p[y1 * Wc + x0] * sw;
            if (y1 >= 0 && y1 < Hc && x1 >= 0 && x1 < Wc) r += p[y1 * Wc + x1] * se;
            o[ch] = r;
            ss += (double)r * (double)r;
        }
        float nrm = (float)sqrt(ss);
        if (nrm < 1e-12f) nrm = 1e-12f;
        for (int ch = 0; ch < C; ch++) o[ch] = o[ch] / nrm;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* Association: DescriptorDistance, the Frame grid, the ExtendMapMatches search core             */
/* ------------------------------------------------------------------------------------------- */

/* DescriptorDistance, feature/src/MapPoint.cpp:22-29: ||a-b||_2 over 256 floats.  Eigen's vectorised
 * reduction order is unspecified; the contract shared with the CUDA path is the fixed order below:
 * 32 strided partial sums (lane l adds elements l, l+32, ... in order), then a butterfly
 * (xor 16,8,4,2,1) -- every lane ends with the same value. */
float ppgo_descriptor_distance(const float *a, const float *b, int dim) {
    float part[32];
    for (int l = 0; l < 32; l++) {
        float s = 0.f;
        for (int k = l; k < dim; k += 32) {
            float d = a[k] - b[k];
            s = s + d * d;
        }
        part[l] = s;
    }
    for (int m = 16; m >= 1; m >>= 1) {
        float t[32];
        for (int l = 0; l < 32; l++) t[l] = part[l] + part[l ^ m];
        memcpy(part, t, sizeof(t));
    }
    return sqrtf(part[0]);
}

typedef struct {
    int minX, minY, maxX, maxY; /* mnMinX.. GeometricCamera.cpp:26-61 */
    float wInv, hInv;           /* mfGridElementWidthInv/HeightInv */
} ppgo_bounds;

/* GeometricCamera::InitializeImageBounds, sensors/src/GeometricCamera.cpp:26-61 */
void ppgo_image_bounds(const ppgo_cfg *c, ppgo_bounds *b) {
    if (!c->fisheye) {
        float cor[8] = {0.f, 0.f, (float)c->width, 0.f, 0.f, (float)c->height, (float)c->width, (float)c->height};
        float u[8];
        ppgo_undistort_points_pinhole(c->K, c->D, cor, 4, u);
        b->minX = (int)fminf(u[0], u[4]);
        b->maxX = (int)fmaxf(u[2], u[6]);
        b->minY = (int)fminf(u[1], u[3]);
        b->maxY = (int)fmaxf(u[5], u[7]);
    } else {
        b->minX = 0;
        b->minY = 0;
        b->maxX = c->width;
        b->maxY = c->height;
    }
    b->wInv = 64.0f / (float)(b->maxX - b->minX);
    b->hInv = 48.0f / (float)(b->maxY - b->minY);
}

/* Frame::PosInGrid, map/src/Frame.cpp:317-327.  -> 1 if indexable */
int ppgo_pos_in_grid(const ppgo_bounds *b, float x, float y, int *px, int *py) {
    *px = (int)roundf((x - (float)b->minX) * b->wInv);
    *py = (int)roundf((y - (float)b->minY) * b->hInv);
    if (*px < 0 || *px >= 64 || *py < 0 || *py >= 48) return 0;
    return 1;
}

/* Frame::AssignFeaturesToGrid + GetFeaturesInArea, map/src/Frame.cpp:138-156, 262-315, literally.
 * grid_off[64*48+1] / grid_idx[n] is the cell lists in (ix*48+iy) order.  -> count, indices into out. */
void ppgo_grid_build(const ppgo_bounds *b, int n, const float *x, const float *y, int *grid_off, int *grid_idx) {
    int *cell = malloc(sizeof(int) * (n > 0 ? n : 1));
    memset(grid_off, 0, sizeof(int) * (64 * 48 + 1));
    for (int i = 0; i < n; i++) {
        int px, py;
        cell[i] = ppgo_pos_in_grid(b, x[i], y[i], &px, &py) ? px * 48 + py : -1;
        if (cell[i] >= 0) grid_off[cell[i] + 1]++;
    }
    for (int k = 0; k < 64 * 48; k++) grid_off[k + 1] += grid_off[k];
    int *fill = calloc(64 * 48, sizeof(int));
    for (int i = 0; i < n; i++)
        if (cell[i] >= 0) grid_idx[grid_off[cell[i]] + fill[cell[i]]++] = i;
    free(fill);
    free(cell);
}

int ppgo_features_in_area(const ppgo_bounds *b, const int *grid_off, const int *grid_idx, const float *kx,
                          const float *ky, float x, float y, float r, int *out) {
    int cnt = 0;
    int nMinCellX = (int)floorf((x - (float)b->minX - r) * b->wInv);
    if (nMinCellX < 0) nMinCellX = 0;
    if (nMinCellX >= 64) return 0;
    int nMaxCellX = (int)ceilf((x - (float)b->minX + r) * b->wInv);
    if (nMaxCellX > 63) nMaxCellX = 63;
    if (nMaxCellX < 0) return 0;
    int nMinCellY = (int)floorf((y - (float)b->minY - r) * b->hInv);
    if (nMinCellY < 0) nMinCellY = 0;
    if (nMinCellY >= 48) return 0;
    int nMaxCellY = (int)ceilf((y - (float)b->minY + r) * b->hInv);
    if (nMaxCellY > 47) nMaxCellY = 47;
    if (nMaxCellY < 0) return 0;
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            int c0 = grid_off[ix * 48 + iy], c1 = grid_off[ix * 48 + iy + 1];
            for (int t = c0; t < c1; t++) {
                int id = grid_idx[t];
                float dx = kx[id] - x, dy = ky[id] - y;
                if (fabsf(dx) < r && fabsf(dy) < r) out[cnt++] = id;
            }
        }
    return cnt;
}

/* Search core of the projection matchers for one map point with the frame state frozen
 * (free_mask[idx]!=0 <=> keypoint idx is NOT skipped).  This is the data-parallel contract of
 * ppg_associate(); the whole function (sequential consumption + seed growing) is ppgo_extend_map_matches below.
 * Candidate order =
 * GetFeaturesInArea order (cell-major), ties keep the first (strict <).
 *   mode 0  Matcher::ExtendMapMatches, matching/src/Matcher.cpp:224-281: r = th * (viewCos > 0.998 ? 2.5 : 4),
 *           accept = !(best > TH_HIGH && best > ratio * second) (:276)
 *   mode 1  the "best only" cores: SearchByProjection(Cur, Last) :31-87 (max_dist = TH_HIGH),
 *           SearchByProjection(F, KF, sFound, th, descDist) :1337-1411 (max_dist = descDist), Fuse :897-1036
 *           (max_dist = TH_LOW, e2_max = 5.99: candidates with ex*ex + ey*ey > 5.99 are skipped, :1000-1005):
 *           r = th, accept = best <= max_dist.
 * -> accept flag; outputs best/second idx and distances (second_idx=-1, d=1e6 when absent). */
int ppgo_search_core(const ppgo_bounds *b, const int *grid_off, const int *grid_idx, int n, const float *kx,
                     const float *ky, const float *frame_desc, const uint8_t *free_mask, const float *mp_desc,
                     float proj_x, float proj_y, float view_cos, float th, float ratio, float th_high, int mode,
                     float max_dist, double e2_max, int *best_idx, int *second_idx, float *best_d,
                     float *second_d) {
    float bestDist = 1e6f, bestDist2 = 1e6f;
    int bestIdx = -1, bestIdx2 = -1;
    float r = th;
    if (mode == 0) {
        if ((double)view_cos > 0.998) /* :240-244: float compared with a double literal, r *= double */
            r = (float)((double)r * 2.5);
        else
            r = (float)((double)r * 4.0);
    }
    int *cand = malloc(sizeof(int) * (n > 0 ? n : 1));
    int nc = ppgo_features_in_area(b, grid_off, grid_idx, kx, ky, proj_x, proj_y, r, cand);
    for (int t = 0; t < nc; t++) {
        int idx = cand[t];
        if (!free_mask[idx]) continue;
        if (mode == 1 && e2_max > 0.0) { /* Fuse :1000-1005: float e2 compared with the double literal */
            float ex = proj_x - kx[idx], ey = proj_y - ky[idx];
            float e2 = ex * ex + ey * ey;
            if ((double)e2 > e2_max) continue;
        }
        float dist = ppgo_descriptor_distance(mp_desc, frame_desc + (size_t)idx * 256, 256);
        if (dist < bestDist) {
            bestDist2 = bestDist;
            bestIdx2 = bestIdx;
            bestDist = dist;
            bestIdx = idx;
        } else if (dist < bestDist2) {
            bestDist2 = dist;
            bestIdx2 = idx;
        }
    }
    free(cand);
    *best_idx = bestIdx;
    *second_idx = bestIdx2;
    *best_d = bestDist;
    *second_d = bestDist2;
    if (nc == 0 || bestIdx < 0) return 0;
    if (mode == 1) return bestDist <= max_dist ? 1 : 0; /* :78, :1399, :1016 */
    if (bestDist > th_high && bestDist > ratio * bestDist2) return 0; /* :276 */
    return 1;
}

void ppgo_search_all(const ppgo_cfg *c, int n, const float *kx, const float *ky, const float *frame_desc,
                     const uint8_t *free_mask, int m, const float *map_desc, const float *proj_uv,
                     const float *view_cos, float th, float ratio, float th_high, int mode, float max_dist,
                     double e2_max, int *best_idx, int *second_idx, float *best_d, float *second_d,
                     uint8_t *accept) {
    ppgo_bounds b;
    ppgo_image_bounds(c, &b);
    int *goff = malloc(sizeof(int) * (64 * 48 + 1)), *gidx = malloc(sizeof(int) * (n > 0 ? n : 1));
    ppgo_grid_build(&b, n, kx, ky, goff, gidx);
    for (int j = 0; j < m; j++)
        accept[j] = (uint8_t)ppgo_search_core(&b, goff, gidx, n, kx, ky, frame_desc, free_mask,
                                              map_desc + (size_t)j * 256, proj_uv[2 * j], proj_uv[2 * j + 1],
                                              view_cos[j], th, ratio, th_high, mode, max_dist, e2_max, &best_idx[j],
                                              &second_idx[j], &best_d[j], &second_d[j]);
    free(goff);
    free(gidx);
}

/* ------------------------------------------------------------------------------------------- */
/* MapPoint::ComputeDistinctiveDescriptors, feature/src/MapPoint.cpp:234-302: the representative */
/* descriptor of a map point = the observation with the least median distance to the others.    */
/* ------------------------------------------------------------------------------------------- */
static int ppgo_cmp_float(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* desc: n x 256 observation descriptors in the order the caller iterates mObservations.
 * -> BestIdx (:279-291: BestMedian starts at 1.0f, strict <, so index 0 when no median is below 1). */
int ppgo_distinctive_one(const float *desc, int n) {
    if (n <= 0) return -1;
    float *D = malloc(sizeof(float) * (size_t)n * n), *row = malloc(sizeof(float) * (size_t)n);
    for (int i = 0; i < n; i++) {
        D[(size_t)i * n + i] = 0.f; /* :270 */
        for (int j = i + 1; j < n; j++) {
            float d = ppgo_descriptor_distance(desc + (size_t)i * 256, desc + (size_t)j * 256, 256);
            D[(size_t)i * n + j] = d;
            D[(size_t)j * n + i] = d;
        }
    }
    float best_median = 1.0f;
    int best = 0;
    for (int i = 0; i < n; i++) {
        memcpy(row, D + (size_t)i * n, sizeof(float) * (size_t)n);
        qsort(row, (size_t)n, sizeof(float), ppgo_cmp_float); /* :284 std::sort */
        float median = row[(size_t)(0.5 * (double)(n - 1))];  /* :285 vDists[0.5 * (N - 1)] */
        if (median < best_median) {
            best_median = median;
            best = i;
        }
    }
    free(D);
    free(row);
    return best;
}

/* offsets[n_points + 1] into the packed descriptor array */
void ppgo_distinctive_all(const float *desc, const int *offsets, int n_points, int *best_idx) {
    for (int p = 0; p < n_points; p++)
        best_idx[p] = ppgo_distinctive_one(desc + (size_t)offsets[p] * 256, offsets[p + 1] - offsets[p]);
}


/* ------------------------------------------------------------------------------------------- */
/* Matcher::ExtendMapMatches, matching/src/Matcher.cpp:203-381, whole function: the sequential  */
/* walk over the candidate map points (search core above with the LIVE frame state) and the    */
/* seed growing over (map edges of pMP) x (key edges of the matched keypoint).                 */
/*                                                                                             */
/* POD form of the pointer graph.  The table has P rows = the candidate map points AND every   */
/* map point reachable as theOtherPt of one of their edges (its descriptor is read at :330):   */
/*   candidate[p]   !isBad() && mbTrackInView (:210-215)                                       */
/*   observed[p]    Observations() > 0 (:253)             bad[p]  isBad() (:227, :364)         */
/*   edge_off/edge_other/edge_ok   CSR of MapPoint::getEdges() in vector order: the row of     */
/*                  theOtherPt(pMP) (-1 = nullptr) and !isBad() && mbValid of each edge (:311)  */
/*   tracked[p]     in/out: mnTrackedbyFrame == F.mnId                                         */
/* Frame: keypoints (mvKeysUn[i].mPos), descriptors, key edges (mvKeyEdges start/end), CSR of  */
/*   mvConnected; kp_mp[i] in/out = F.mvpMapPoints[i] as a table row, -1 = nullptr, -2 = a map */
/*   point outside the table that has observations; kedge_me[e] in/out = F.mvpMapEdges[e] as   */
/*   the CSR position of the map edge (-1 = nullptr).                                          */
/* std::sort by getEdges().size() descending is unstable (:223-224): ties keep table order here */
/* (same documented divergence as the other sorts).  Requires ratio < 1 (with ratio >= 1 the    */
/* reference writes mvpMapPoints[-1] when every keypoint of a window is taken).                */
/* -> nmatches as the reference counts it (two increments per accepted map point, :281, :378). */
/* ------------------------------------------------------------------------------------------- */
static int emm_occupied(const int *kp_mp, const uint8_t *observed, int idx) {
    int r = kp_mp[idx];
    return r >= 0 ? (observed[r] != 0) : (r == -2); /* :253 */
}

typedef struct {
    int row, deg, pos;
} emm_key;

static int emm_cmp(const void *a, const void *b) {
    const emm_key *x = a, *y = b;
    if (x->deg != y->deg) return x->deg > y->deg ? -1 : 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);
}

int ppgo_extend_map_matches(const ppgo_cfg *c, int P, const float *map_desc, const uint8_t *candidate,
                            const uint8_t *observed, const uint8_t *bad, const int *edge_off, const int *edge_other,
                            const uint8_t *edge_ok, const float *proj_uv, const float *view_cos, uint8_t *tracked,
                            int n, const float *kx, const float *ky, const float *frame_desc, int *kp_mp,
                            const int *kedge_start, const int *kedge_end, const int *conn_off, const int *conn_idx,
                            int *kedge_me, float th, float ratio, float th_high) {
    int nmatches = 0;
    ppgo_bounds b;
    ppgo_image_bounds(c, &b);
    int *goff = malloc(sizeof(int) * (64 * 48 + 1)), *gidx = malloc(sizeof(int) * (n > 0 ? n : 1));
    ppgo_grid_build(&b, n, kx, ky, goff, gidx);
    emm_key *order = malloc(sizeof(emm_key) * (P > 0 ? P : 1));
    int nc = 0;
    for (int p = 0; p < P; p++) /* :210-215 */
        if (candidate[p] && !bad[p]) {
            order[nc].row = p;
            order[nc].deg = edge_off[p + 1] - edge_off[p];
            order[nc].pos = nc;
            nc++;
        }
    qsort(order, nc, sizeof(emm_key), emm_cmp); /* :223-224 */
    int *win = malloc(sizeof(int) * (n > 0 ? n : 1));
    int *queue = malloc(sizeof(int) * ((size_t)P + 2)); /* every push marks a new map point tracked */
    for (int t = 0; t < nc; t++) {
        const int p = order[t].row;
        if (tracked[p] || bad[p]) continue; /* :229 */
        const float *dmp = map_desc + (size_t)p * 256;
        float bestDist = 1e6f, bestDist2 = 1e6f;
        int bestIdx = -1;
        float r = th;
        if ((double)view_cos[p] > 0.998) /* :240-244 */
            r = (float)((double)r * 2.5);
        else
            r = (float)((double)r * 4.0);
        int nw = ppgo_features_in_area(&b, goff, gidx, kx, ky, proj_uv[2 * p], proj_uv[2 * p + 1], r, win);
        if (nw == 0) continue; /* :247-248 */
        for (int k = 0; k < nw; k++) {
            int idx = win[k];
            if (emm_occupied(kp_mp, observed, idx)) continue; /* :253 */
            float dist = ppgo_descriptor_distance(dmp, frame_desc + (size_t)idx * 256, 256);
            if (dist < bestDist) {
                bestDist2 = bestDist;
                bestDist = dist;
                bestIdx = idx;
            } else if (dist < bestDist2)
                bestDist2 = dist;
        }
        if (bestDist > th_high && bestDist > ratio * bestDist2) continue; /* :276 */
        if (bestIdx < 0) continue; /* unreachable for ratio < 1 (see header) */
        kp_mp[bestIdx] = p; /* :279 */
        tracked[p] = 1;
        nmatches++;
        /* seed growing, :287-377 */
        int qh = 0, qt = 0, qcap = P + 2;
        queue[qt++] = bestIdx;
        const int me0 = edge_off[p], nme = edge_off[p + 1] - edge_off[p];
        while (qh < qt) {
            const int keyID = queue[qh++];
            const int ke0 = conn_off[keyID], nke = conn_off[keyID + 1] - conn_off[keyID];
            if (nme == 0 || nke == 0) continue; /* :300-301 */
            float *weight = malloc(sizeof(float) * (size_t)nme * nke);
            for (int i = 0; i < nme * nke; i++) weight[i] = 1e6f; /* :308 */
            int *lx = malloc(sizeof(int) * nme), *ly = malloc(sizeof(int) * nke);
            int nlx = 0, nly = 0;
            for (int i = 0; i < nme; i++) { /* :312-318 */
                if (!edge_ok[me0 + i] || edge_other[me0 + i] < 0) continue;
                lx[nlx++] = i;
            }
            for (int j = 0; j < nke; j++) ly[nly++] = j; /* :321-322 */
            for (int a = 0; a < nlx; a++)
                for (int j = 0; j < nke; j++) { /* :324-340 */
                    const int i = lx[a];
                    const int po = edge_other[me0 + i];
                    const int e = conn_idx[ke0 + j];
                    const int ko = kedge_start[e] == keyID ? kedge_end[e] : kedge_start[e];
                    if (po == kp_mp[ko])
                        weight[i * nke + j] = -1.f;
                    else
                        weight[i * nke + j] =
                            ppgo_descriptor_distance(map_desc + (size_t)po * 256, frame_desc + (size_t)ko * 256, 256);
                }
            while (nlx > 0 && nly > 0) { /* :342-374 */
                int minlx = 0, minly = 0;
                float minWeight = 1e6f;
                for (int a = 0; a < nlx; a++)
                    for (int bb = 0; bb < nly; bb++)
                        if (weight[lx[a] * nke + ly[bb]] < minWeight) {
                            minWeight = weight[lx[a] * nke + ly[bb]];
                            minlx = a;
                            minly = bb;
                        }
                if (minWeight > th_high) break;
                const int mi = lx[minlx], kj = ly[minly];
                memmove(lx + minlx, lx + minlx + 1, sizeof(int) * (nlx - minlx - 1));
                nlx--;
                memmove(ly + minly, ly + minly + 1, sizeof(int) * (nly - minly - 1));
                nly--;
                const int po = edge_other[me0 + mi];
                const int e = conn_idx[ke0 + kj];
                const int ko = kedge_start[e] == keyID ? kedge_end[e] : kedge_start[e];
                if (po < 0 || bad[po] || tracked[po]) continue; /* :364-365 */
                kp_mp[ko] = po; /* :366 */
                kedge_me[e] = me0 + mi;
                tracked[po] = 1;
                if (qt < qcap) queue[qt++] = ko;
            }
            free(weight);
            free(lx);
            free(ly);
        }
        nmatches++; /* :378 */
    }
    free(queue);
    free(win);
    free(order);
    free(goff);
    free(gidx);
    return nmatches;
}

/* ------------------------------------------------------------------------------------------- */
/* Frame::CheckInFrustum, map/src/Frame.cpp:223-260, with Pinhole::project (sensors/src/        */
/* Pinhole.cpp:32-38) / KannalaBrandt8::project (sensors/src/KannalaBrandt8.cpp:44-59) and      */
/* GeometricCamera::IsInImage (sensors/src/GeometricCamera.cpp:21-24).  It produces the inputs  */
/* of ExtendMapMatches: mbTrackInView, mTrackProjX/Y, mTrackDepth, mTrackViewCos.               */
/* Eigen's evaluation order for the 3-vectors is taken as: a row of a 3x3 product and a dot /   */
/* squared norm are (a0*b0 + a1*b1) + a2*b2 (unrolled redux of a fixed-size expression), the    */
/* translation is added afterwards; no FMA (baseline x86-64).  Eigen is not available here: this */
/* order is the stand-in's (oracle/ref_standins), against which the reference's own Frame.cpp,   */
/* Pinhole.cpp and KannalaBrandt8.cpp run this function in tests/test_ref_pin_kb8.py.             */
/* KannalaBrandt8: atan2f as o_atan2f; `cos(psi)` / `sin(psi)` are unqualified calls on a float  */
/* -> double routine, the product is carried in double and rounded once on assignment (both     */
/* libm variants; see o_kb8_project).                                                           */
/* out = {u, v, depth, viewCos}; returns mbTrackInView.  Not-in-view rows keep the reference's  */
/* reset values (-1, -1, -1) and viewCos 0.                                                      */
/* ------------------------------------------------------------------------------------------- */
static float o_dot3(const float *a, const float *b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

/* KannalaBrandt8::project(const Eigen::Vector3f&), sensors/src/KannalaBrandt8.cpp:44-59.  cam8 = fx fy cx cy k0..k3. */
static void o_kb8_project(const float *cam8, const float *Pc, float *u, float *v) {
    const float x2y2 = Pc[0] * Pc[0] + Pc[1] * Pc[1];
    const float theta = o_atan2f(sqrtf(x2y2), Pc[2]);
    const float psi = o_atan2f(Pc[1], Pc[0]);
    const float theta2 = theta * theta, theta3 = theta * theta2, theta5 = theta3 * theta2;
    const float theta7 = theta5 * theta2, theta9 = theta7 * theta2;
    const float r = theta + cam8[4] * theta3 + cam8[5] * theta5 + cam8[6] * theta7 + cam8[7] * theta9;
    /* `cos(psi)` / `sin(psi)`: unqualified calls in a file without `using namespace std` -> ::cos(double); the product   */
    /* is carried in double and rounded once on assignment.  Confirmed by the reference's own file compiled first in  */
    /* oracle/ref_tu/matcher_tu.cpp: with PPGO_LIBM_FLOAT=1 (the literal atan2f) this function is bit-identical to it. */
    *u = (float)((double)(cam8[0] * r) * cos((double)psi) + (double)cam8[2]);
    *v = (float)((double)(cam8[1] * r) * sin((double)psi) + (double)cam8[3]);
}

int ppgo_check_in_frustum(const ppgo_cfg *c, const ppgo_bounds *b, const float *Rcw, const float *tcw,
                          const float *Ow, const float *P, const float *Pn, float minD, float maxD, float cos_limit,
                          float *out) {
    out[0] = -1.f; /* :226-228 */
    out[1] = -1.f;
    out[2] = -1.f;
    out[3] = 0.f;
    float Pc[3];
    for (int i = 0; i < 3; i++) Pc[i] = o_dot3(Rcw + 3 * i, P) + tcw[i]; /* :232 */
    if (Pc[2] < 0.0f) return 0;                                         /* :234 */
    const float fx = c->K[0], fy = c->K[4], cx = c->K[2], cy = c->K[5];
    float u, v;
    if (!c->fisheye) { /* Pinhole.cpp:35-36 */
        u = fx * Pc[0] / Pc[2] + cx;
        v = fy * Pc[1] / Pc[2] + cy;
    } else { /* KannalaBrandt8.cpp:46-58 */
        const float cam8[8] = {fx, fy, cx, cy, c->D[0], c->D[1], c->D[2], c->D[3]};
        o_kb8_project(cam8, Pc, &u, &v);
    }
    if (!(u >= (float)b->minX && u < (float)b->maxX && v >= (float)b->minY && v < (float)b->maxY)) return 0; /* :238 */
    float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]}; /* :243 */
    const float dist = sqrtf(o_dot3(PO, PO));                  /* :244 */
    if (dist < minD || dist > maxD) return 0;                  /* :245 */
    const float viewCos = o_dot3(PO, Pn) / dist;               /* :249 */
    if (viewCos < cos_limit) return 0;                         /* :250 */
    out[0] = u;
    out[1] = v;
    out[2] = dist;
    out[3] = viewCos;
    return 1;
}

void ppgo_check_in_frustum_all(const ppgo_cfg *c, const float *Rcw, const float *tcw, const float *Ow, int m,
                               const float *world_pos, const float *normal, const float *min_dist,
                               const float *max_dist, float cos_limit, uint8_t *in_view, float *proj_uv,
                               float *depth, float *view_cos) {
    ppgo_bounds b;
    ppgo_image_bounds(c, &b);
    for (int j = 0; j < m; j++) {
        float o[4];
        in_view[j] = (uint8_t)ppgo_check_in_frustum(c, &b, Rcw, tcw, Ow, world_pos + 3 * j, normal + 3 * j,
                                                    min_dist[j], max_dist[j], cos_limit, o);
        proj_uv[2 * j] = o[0];
        proj_uv[2 * j + 1] = o[1];
        depth[j] = o[2];
        view_cos[j] = o[3];
    }
}

/* ------------------------------------------------------------------------------------------- */
/* DBoW3::Vocabulary::transform(features, BowVector&, FeatureVector&, levelsup) as called from  */
/* Frame::ComputeBoW (map/src/Frame.cpp:331-340, levelsup = 4) and KeyFrame construction        */
/* (:127-131).  DBoW3 is a third-party dependency (find_package(DBoW3), CMakeLists.txt:18,      */
/* unpinned, not under /root/reference); this restates its published algorithm:                  */
/*   per feature: descend from the root, at every level take the child with the least           */
/*     DescManip::distance = sum_i (float)((a_i - b_i) * (a_i - b_i)) accumulated in double in   */
/*     index order, strict < so the first child wins ties; word = leaf's word id, weight = leaf's */
/*     weight; node for the FeatureVector = the node reached at level L - levelsup, or the root   */
/*     (0) when that level is <= 0 (the reference: L = 3, levelsup = 4 -> always the root);        */
/*   BowVector: features with weight > 0 add their weight to their word (std::map, accumulated in */
/*     feature order), then normalize: L1 (scoring 0, 2, 3, 4), L2 (scoring 1) over the words in  */
/*     ascending id, or no norm (scoring 5, DOT_PRODUCT): every value divided by the word count.  */
/* Weighting TF_IDF (0) / TF (1) only (IDF / BINARY use addIfNotExist).  "Parity unpinned": no    */
/* DBoW3 here; the reference's own vocabulary files pin the file format (ppg_slam_b200/            */
/* vocabulary.py).  Outputs: per feature word / weight / node (-1 when weight <= 0), BowVector    */
/* as sorted (word, value) pairs -> returns their number.                                         */
/* ------------------------------------------------------------------------------------------- */
static double bow_distance(const float *a, const float *b, int dim) {
    double sqd = 0.;
    for (int i = 0; i < dim; i++) sqd += (a[i] - b[i]) * (a[i] - b[i]); /* float product, double sum */
    return sqd;
}

int ppgo_bow_transform(int k, int L, int scoring, const int *children, const double *node_weight,
                       const int *node_word, const float *node_desc, int dim, int n, const float *feat, int levelsup,
                       int *f_word, double *f_weight, int *f_node, int *bow_word, double *bow_value) {
    const int nid_level = L - levelsup;
    int nb = 0;
    for (int i = 0; i < n; i++) {
        int final_id = 0, current_level = 0, nid = 0;
        do {
            ++current_level;
            double best_d = 1.7976931348623157e308;
            const int *ch = children + (size_t)final_id * k;
            int next = final_id;
            for (int c = 0; c < k && ch[c] >= 0; c++) {
                double d = bow_distance(feat + (size_t)i * dim, node_desc + (size_t)ch[c] * dim, dim);
                if (d < best_d) {
                    best_d = d;
                    next = ch[c];
                }
            }
            final_id = next;
            if (current_level == nid_level) nid = final_id;
        } while (children[(size_t)final_id * k] >= 0);
        f_word[i] = node_word[final_id];
        f_weight[i] = node_weight[final_id];
        f_node[i] = f_weight[i] > 0 ? nid : -1;
        if (f_weight[i] > 0) { /* v.addWeight(id, w) */
            int lo = 0;
            while (lo < nb && bow_word[lo] < f_word[i]) lo++;
            if (lo < nb && bow_word[lo] == f_word[i])
                bow_value[lo] += f_weight[i];
            else {
                memmove(bow_word + lo + 1, bow_word + lo, sizeof(int) * (nb - lo));
                memmove(bow_value + lo + 1, bow_value + lo, sizeof(double) * (nb - lo));
                bow_word[lo] = f_word[i];
                bow_value[lo] = f_weight[i];
                nb++;
            }
        }
    }
    if (nb > 0) {
        if (scoring == 5) { /* !mustNormalize */
            const double nd = (double)nb;
            for (int j = 0; j < nb; j++) bow_value[j] /= nd;
        } else {
            double norm = 0.0;
            if (scoring == 1) {
                for (int j = 0; j < nb; j++) norm += bow_value[j] * bow_value[j];
                norm = sqrt(norm);
            } else
                for (int j = 0; j < nb; j++) norm += fabs(bow_value[j]);
            if (norm > 0.0)
                for (int j = 0; j < nb; j++) bow_value[j] /= norm;
        }
    }
    return nb;
}

/* Search core of the bag-of-words matchers for one keyframe feature with the frame state frozen:
 * Matcher::SearchByBoW(KF, F, ...) matching/src/Matcher.cpp:393-477 (inner loop :421-461) and SearchByBoW(KF1, KF2)
 * :663-754.  Candidates = the frame features listed under the same FeatureVector node, in ascending index (vIndicesF
 * is filled in feature order), skipped when free_mask[idx] == 0 (:434 / :707-711); best / second by strict <;
 * accept = best <= max_dist && best < ratio * second (:456-458). */
int ppgo_search_node_core(int n, const float *frame_desc, const uint8_t *free_mask, const int *kp_node,
                          const float *row_desc, int row_node, float ratio, float max_dist, int *best_idx,
                          int *second_idx, float *best_d, float *second_d) {
    float bestDist1 = 1e6f, bestDist2 = 1e6f;
    int bestIdxF = -1, bestIdx2 = -1;
    if (row_node >= 0)
        for (int idx = 0; idx < n; idx++) {
            if (kp_node[idx] != row_node || !free_mask[idx]) continue;
            float dist = ppgo_descriptor_distance(row_desc, frame_desc + (size_t)idx * 256, 256);
            if (dist < bestDist1) {
                bestDist2 = bestDist1;
                bestIdx2 = bestIdxF;
                bestDist1 = dist;
                bestIdxF = idx;
            } else if (dist < bestDist2) {
                bestDist2 = dist;
                bestIdx2 = idx;
            }
        }
    *best_idx = bestIdxF;
    *second_idx = bestIdx2;
    *best_d = bestDist1;
    *second_d = bestDist2;
    if (bestIdxF < 0) return 0;
    return (bestDist1 <= max_dist && bestDist1 < ratio * bestDist2) ? 1 : 0;
}

void ppgo_search_node_all(int n, const float *frame_desc, const uint8_t *free_mask, const int *kp_node, int m,
                          const float *row_desc, const int *row_node, float ratio, float max_dist, int *best_idx,
                          int *second_idx, float *best_d, float *second_d, uint8_t *accept) {
    for (int j = 0; j < m; j++)
        accept[j] = (uint8_t)ppgo_search_node_core(n, frame_desc, free_mask, kp_node, row_desc + (size_t)j * 256,
                                                   row_node[j], ratio, max_dist, &best_idx[j], &second_idx[j],
                                                   &best_d[j], &second_d[j]);
}

/* Matcher::SearchByBoW whole, matching/src/Matcher.cpp:393-477 (KF vs Frame: strict = 0, `bestDist1 <= TH_LOW`) and
 * :663-754 (KF1 vs KF2: strict = 1, `bestDist1 < TH_LOW`).  Rows = the keyframe features that hold a good map point
 * (the callers' `if (!pMP) continue; if (pMP->isBad()) continue;`), listed in the order the reference visits them;
 * kp_node = FeatureVector node of every feature of the other side (-1: not listed / no good map point).  The two
 * FeatureVector maps are walked in ascending node id; kp_row[idx] = the row matched to feature idx (vpMapPointMatches /
 * vbMatched2), -1 otherwise.  -> nmatches. */
int ppgo_search_by_bow(int n, const float *frame_desc, const int *kp_node, int m, const float *row_desc,
                       const int *row_node, float ratio, float max_dist, int strict, int *kp_row) {
    int nmatches = 0;
    for (int i = 0; i < n; i++) kp_row[i] = -1;
    int max_node = -1;
    for (int j = 0; j < m; j++)
        if (row_node[j] > max_node) max_node = row_node[j];
    for (int node = 0; node <= max_node; node++) { /* KFit->first == Fit->first, ascending */
        for (int j = 0; j < m; j++) {
            if (row_node[j] != node) continue;
            const float *dKF = row_desc + (size_t)j * 256;
            float bestDist1 = 1e6f, bestDist2 = 1e6f;
            int bestIdxF = -1;
            for (int idx = 0; idx < n; idx++) {
                if (kp_node[idx] != node) continue;
                if (kp_row[idx] >= 0) continue; /* :434 / :707 */
                float dist = ppgo_descriptor_distance(dKF, frame_desc + (size_t)idx * 256, 256);
                if (dist < bestDist1) {
                    bestDist2 = bestDist1;
                    bestDist1 = dist;
                    bestIdxF = idx;
                } else if (dist < bestDist2)
                    bestDist2 = dist;
            }
            if (strict ? bestDist1 < max_dist : bestDist1 <= max_dist) /* :733 / :456 */
                if (bestDist1 < ratio * bestDist2) {
                    kp_row[bestIdxF] = j;
                    nmatches++;
                }
        }
    }
    return nmatches;
}

/* ------------------------------------------------------------------------------------------- */
/* Matcher::SearchForInitialization, matching/src/Matcher.cpp:582-651 (called from               */
/* system/src/Tracking.cpp:525 with windowSize = 50).  For every feature i1 of F1 in index order:   */
/* window of radius windowSize around vbPrevMatched[i1] in F2 (Frame::GetFeaturesInArea), best /     */
/* second best DescriptorDistance over the window features that are not yet matched, accept iff       */
/* best <= TH_LOW && best < ratio * second.                                                           */
/* Quirk kept: vMatchedDistance is a vector<int> -- the stored distance of a matched F2 feature        */
/* truncates to 0 (every accepted distance is <= TH_LOW < 1), so `vMatchedDistance[i2] <= dist` (:613)  */
/* is true for every later candidate: a matched F2 feature is never offered again, and the             */
/* "steal the match back" branch (:634-638) is unreachable.  INT_MAX <= dist (unmatched) is false.      */
/* matches12[i1] = index in F2 or -1; prev (n1 x 2, in/out) gets the matched feature's position (:646). */
/* -> nmatches.                                                                                          */
/* ------------------------------------------------------------------------------------------- */
int ppgo_search_for_initialization(const ppgo_cfg *c, int n1, const float *desc1, float *prev, int n2,
                                   const float *kx2, const float *ky2, const float *desc2, int window, float ratio,
                                   float th_low, int *matches12) {
    ppgo_bounds b;
    ppgo_image_bounds(c, &b);
    int *goff = malloc(sizeof(int) * (64 * 48 + 1)), *gidx = malloc(sizeof(int) * (n2 > 0 ? n2 : 1));
    ppgo_grid_build(&b, n2, kx2, ky2, goff, gidx);
    int *win = malloc(sizeof(int) * (n2 > 0 ? n2 : 1));
    int *matched_dist = malloc(sizeof(int) * (n2 > 0 ? n2 : 1)); /* vMatchedDistance: vector<int> */
    int *matches21 = malloc(sizeof(int) * (n2 > 0 ? n2 : 1));
    for (int i = 0; i < n2; i++) {
        matched_dist[i] = 2147483647;
        matches21[i] = -1;
    }
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    int nmatches = 0;
    for (int i1 = 0; i1 < n1; i1++) {
        const int nw = ppgo_features_in_area(&b, goff, gidx, kx2, ky2, prev[2 * i1], prev[2 * i1 + 1], (float)window, win);
        if (nw == 0) continue;
        float bestDist = 1e6f, bestDist2 = 1e6f;
        int bestIdx2 = -1;
        for (int k = 0; k < nw; k++) {
            const int i2 = win[k];
            const float dist = ppgo_descriptor_distance(desc1 + (size_t)i1 * 256, desc2 + (size_t)i2 * 256, 256);
            if ((float)matched_dist[i2] <= dist) continue; /* :613, int -> float */
            if (dist < bestDist) {
                bestDist2 = bestDist;
                bestDist = dist;
                bestIdx2 = i2;
            } else if (dist < bestDist2)
                bestDist2 = dist;
        }
        if (bestDist <= th_low) {
            if (bestDist < bestDist2 * ratio) {
                if (matches21[bestIdx2] >= 0) { /* unreachable, see header */
                    matches12[matches21[bestIdx2]] = -1;
                    nmatches--;
                }
                matches12[i1] = bestIdx2;
                matches21[bestIdx2] = i1;
                matched_dist[bestIdx2] = (int)bestDist;
                nmatches++;
            }
        }
    }
    for (int i1 = 0; i1 < n1; i1++)
        if (matches12[i1] >= 0) {
            prev[2 * i1] = kx2[matches12[i1]];
            prev[2 * i1 + 1] = ky2[matches12[i1]];
        }
    free(goff);
    free(gidx);
    free(win);
    free(matched_dist);
    free(matches21);
    return nmatches;
}

/* ------------------------------------------------------------------------------------------- */
/* Matcher::SearchForTriangulation, matching/src/Matcher.cpp:767-885, for the pinhole camera       */
/* (epipolar test of sensors/src/Pinhole.cpp:98-114).  The FeatureVectors are given per feature:    */
/* node[i] = the vocabulary node the feature hangs under (every feature of a DBoW3 FeatureVector    */
/* belongs to exactly one node, its indices in ascending order), -1 = none.  The merge loop of      */
/* :801-869 pairs the nodes present in both key frames, i.e. node1[i1] == node2[i2].                */
/* For every feature i1 of KF1 without a map point (:812-818): over the features i2 of KF2 in the   */
/* same node, in index order, without a map point (:831-836; vbMatched2 is never set in the         */
/* reference, so "already matched" never fires): dist = DescriptorDistance; skip if dist > TH_LOW   */
/* or dist > bestDist (:842, a tie replaces the earlier candidate); skip if closer than 10 px to    */
/* the epipole (:846-847, Eigen norm of a 2-vector = sqrt(x*x + y*y)); keep if the epipolar line    */
/* of kp1 passes within dsqr < 3.84 of kp2 (Pinhole.cpp:106-113: a, b, c, num, den in float, left   */
/* to right, the comparison in double).  F12 (row-major) = K1^-T * [t12]x * R12 * K2^-1 and the     */
/* epipole are the caller's (they depend on the poses only, :776-788, Pinhole.cpp:101-104).         */
/* match12[i1] = i2 or -1 -> nmatches.                                                               */
/* ------------------------------------------------------------------------------------------- */
int ppgo_search_for_triangulation(int n1, const float *desc1, const int *node1, const unsigned char *has_mp1,
                                  const float *pos1, int n2, const float *desc2, const int *node2,
                                  const unsigned char *has_mp2, const float *pos2, const float *F12, const float *ep,
                                  float th_low, int *match12) {
    int nmatches = 0;
    for (int i1 = 0; i1 < n1; i1++) {
        match12[i1] = -1;
        if (has_mp1[i1] || node1[i1] < 0) continue;
        const float x1 = pos1[2 * i1], y1 = pos1[2 * i1 + 1];
        const float a = x1 * F12[0] + y1 * F12[3] + F12[6];
        const float b = x1 * F12[1] + y1 * F12[4] + F12[7];
        const float c = x1 * F12[2] + y1 * F12[5] + F12[8];
        const float den = a * a + b * b;
        float bestDist = th_low;
        int bestIdx2 = -1;
        for (int i2 = 0; i2 < n2; i2++) {
            if (node2[i2] != node1[i1] || has_mp2[i2]) continue;
            const float dist = ppgo_descriptor_distance(desc1 + (size_t)i1 * 256, desc2 + (size_t)i2 * 256, 256);
            if (dist > th_low || dist > bestDist) continue;
            const float x2 = pos2[2 * i2], y2 = pos2[2 * i2 + 1];
            const float ex = ep[0] - x2, ey = ep[1] - y2;
            if (sqrtf(ex * ex + ey * ey) < 10.0f) continue;
            if (den == 0) continue;
            const float num = a * x2 + b * y2 + c;
            const float dsqr = num * num / den;
            if ((double)dsqr < 3.84) {
                bestIdx2 = i2;
                bestDist = dist;
            }
        }
        if (bestIdx2 >= 0) {
            match12[i1] = bestIdx2;
            nmatches++;
        }
    }
    return nmatches;
}

/* ------------------------------------------------------------------------------------------- */
/* KannalaBrandt8::epipolarConstrain = TriangulateMatches(...) > 0.0001f, sensors/src/              */
/* KannalaBrandt8.cpp:167-236: unproject both pixels (Newton iteration :62-91), reject small        */
/* parallax, triangulate by the null vector of the 4 x 4 DLT matrix (:224-236), require positive    */
/* depth in both cameras and a reprojection error below 5.991 px^2 in both images.  Everything is   */
/* float, left to right as written; Eigen's matrix-vector products and dots as (a0*b0 + a1*b1) +    */
/* a2*b2 (the stand-in's reading, see ppgo_check_in_frustum).                                       */
/* The one step that is NOT the reference's arithmetic is Eigen::JacobiSVD (:233): Eigen iterates   */
/* two-sided Jacobi rotations in float; ppgo_null_vector4 computes the same vector -- the right     */
/* singular vector of the smallest singular value -- by cyclic Jacobi on A^T A in double and        */
/* rounds it to float (sign: the reference divides by the last component, the sign cancels).        */
/* ------------------------------------------------------------------------------------------- */
void ppgo_null_vector4(const float *A, float *v4) { /* A row-major 4 x 4 */
    double M[4][4], V[4][4];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double acc = (double)A[i] * (double)A[j];
            for (int k = 1; k < 4; k++) acc = acc + (double)A[4 * k + i] * (double)A[4 * k + j];
            M[i][j] = acc;
            V[i][j] = i == j ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 12; sweep++) {
        double off = 0.0;
        for (int p = 0; p < 3; p++)
            for (int q = p + 1; q < 4; q++) off = off + fabs(M[p][q]);
        if (off == 0.0) break;
        for (int p = 0; p < 3; p++)
            for (int q = p + 1; q < 4; q++) {
                const double apq = M[p][q];
                if (apq == 0.0) continue;
                const double theta = (M[q][q] - M[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 4; k++) { /* columns p, q of M and V */
                    const double mkp = M[k][p], mkq = M[k][q];
                    M[k][p] = c * mkp - sn * mkq;
                    M[k][q] = sn * mkp + c * mkq;
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - sn * vkq;
                    V[k][q] = sn * vkp + c * vkq;
                }
                for (int k = 0; k < 4; k++) { /* rows p, q of M */
                    const double mpk = M[p][k], mqk = M[q][k];
                    M[p][k] = c * mpk - sn * mqk;
                    M[q][k] = sn * mpk + c * mqk;
                }
            }
    }
    int best = 0;
    for (int j = 1; j < 4; j++)
        if (M[j][j] < M[best][best]) best = j;
    for (int i = 0; i < 4; i++) v4[i] = (float)V[i][best];
}

/* KannalaBrandt8::unproject, :62-91 (precision = 1e-6f, :11; CV_PI / 2.f is a double narrowed by fmaxf / fminf) */
void ppgo_kb8_unproject(const float *cam8, const float *p2D, float *out3) {
    const float pw0 = (p2D[0] - cam8[2]) / cam8[0], pw1 = (p2D[1] - cam8[3]) / cam8[1];
    float scale = 1.f;
    float theta_d = sqrtf(pw0 * pw0 + pw1 * pw1);
    theta_d = fminf(fmaxf((float)(-PPGO_PI / 2.0), theta_d), (float)(PPGO_PI / 2.0));
    if ((double)theta_d > 1e-8) {
        float theta = theta_d;
        for (int j = 0; j < 10; j++) {
            const float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2,
                        theta8 = theta4 * theta4;
            const float k0_theta2 = cam8[4] * theta2, k1_theta4 = cam8[5] * theta4;
            const float k2_theta6 = cam8[6] * theta6, k3_theta8 = cam8[7] * theta8;
            const float theta_fix = (theta * (1 + k0_theta2 + k1_theta4 + k2_theta6 + k3_theta8) - theta_d) /
                                    (1 + 3 * k0_theta2 + 5 * k1_theta4 + 7 * k2_theta6 + 9 * k3_theta8);
            theta = theta - theta_fix;
            if (fabsf(theta_fix) < 1e-6f) break;
        }
        scale = o_tanf(theta) / theta_d;
    }
    out3[0] = pw0 * scale;
    out3[1] = pw1 * scale;
    out3[2] = 1.f;
}

void ppgo_kb8_project(const float *cam8, const float *P3, float *uv) { o_kb8_project(cam8, P3, uv, uv + 1); }

/* KannalaBrandt8::TriangulateMatches, :175-222, with r1 = unproject(kp1.mPos) from the caller.  R12 row-major.      */
/* Returns what the reference returns (z1, or -1 .. -5); x3D_out (3) may be NULL.                                   */
float ppgo_kb8_triangulate_matches(const float *cam8, const float *r1, const float *pos1, const float *pos2,
                                   const float *R12, const float *t12, float *x3D_out) {
    float r2[3];
    ppgo_kb8_unproject(cam8, pos2, r2);
    float r21[3];
    for (int i = 0; i < 3; i++) r21[i] = o_dot3(R12 + 3 * i, r2); /* :181 */
    const float cosParallaxRays = o_dot3(r1, r21) / (sqrtf(o_dot3(r1, r1)) * sqrtf(o_dot3(r21, r21)));
    if ((double)cosParallaxRays > 0.9998) return -1.f; /* :183 */
    float R21[9], T2[12]; /* Tcw2 = [R21 | -R21 * t12], :198-200; Tcw1 = [I | 0] */
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R21[3 * i + j] = R12[3 * j + i];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) T2[4 * i + j] = R21[3 * i + j];
        T2[4 * i + 3] = ((-R21[3 * i]) * t12[0] + (-R21[3 * i + 1]) * t12[1]) + (-R21[3 * i + 2]) * t12[2];
    }
    float A[16]; /* :227-231, coefficient-wise p * T.row(2) - T.row(k) */
    for (int j = 0; j < 4; j++) {
        const float t1r0 = j == 0 ? 1.f : 0.f, t1r1 = j == 1 ? 1.f : 0.f, t1r2 = j == 2 ? 1.f : 0.f;
        A[j] = r1[0] * t1r2 - t1r0;
        A[4 + j] = r1[1] * t1r2 - t1r1;
        A[8 + j] = r2[0] * T2[8 + j] - T2[j];
        A[12 + j] = r2[1] * T2[8 + j] - T2[4 + j];
    }
    float h[4];
    ppgo_null_vector4(A, h); /* :233-234 (see the note above) */
    const float x3D[3] = {h[0] / h[3], h[1] / h[3], h[2] / h[3]}; /* :235 */
    const float z1 = x3D[2];
    if (z1 <= 0) return -2.f; /* :205 */
    const float z2 = o_dot3(R21 + 6, x3D) + T2[11];
    if (z2 <= 0) return -3.f; /* :209 */
    float uv[2];
    o_kb8_project(cam8, x3D, uv, uv + 1);
    const float e10 = uv[0] - pos1[0], e11 = uv[1] - pos1[1];
    if ((double)(e10 * e10 + e11 * e11) > 5.991) return -4.f; /* :213-214 */
    float x3D2[3];
    for (int i = 0; i < 3; i++) x3D2[i] = o_dot3(R21 + 3 * i, x3D) + T2[4 * i + 3];
    o_kb8_project(cam8, x3D2, uv, uv + 1);
    const float e20 = uv[0] - pos2[0], e21 = uv[1] - pos2[1];
    if ((double)(e20 * e20 + e21 * e21) > 5.991) return -5.f; /* :218-219 */
    if (x3D_out) {
        x3D_out[0] = x3D[0];
        x3D_out[1] = x3D[1];
        x3D_out[2] = x3D[2];
    }
    return z1;
}

/* Matcher::SearchForTriangulation, matching/src/Matcher.cpp:767-885, for the KannalaBrandt8 camera: as           */
/* ppgo_search_for_triangulation with the epipolar test of KannalaBrandt8.cpp:167-172 (R12, t12 of :784-788 from     */
/* the caller, row-major).  pos = mvKeysUn[i].mPos.                                                                  */
int ppgo_search_for_triangulation_kb8(int n1, const float *desc1, const int *node1, const unsigned char *has_mp1,
                                      const float *pos1, int n2, const float *desc2, const int *node2,
                                      const unsigned char *has_mp2, const float *pos2, const float *cam8,
                                      const float *R12, const float *t12, const float *ep, float th_low, int *match12) {
    int nmatches = 0;
    for (int i1 = 0; i1 < n1; i1++) {
        match12[i1] = -1;
        if (has_mp1[i1] || node1[i1] < 0) continue;
        float r1[3];
        ppgo_kb8_unproject(cam8, pos1 + 2 * i1, r1);
        float bestDist = th_low;
        int bestIdx2 = -1;
        for (int i2 = 0; i2 < n2; i2++) {
            if (node2[i2] != node1[i1] || has_mp2[i2]) continue;
            const float dist = ppgo_descriptor_distance(desc1 + (size_t)i1 * 256, desc2 + (size_t)i2 * 256, 256);
            if (dist > th_low || dist > bestDist) continue;
            const float ex = ep[0] - pos2[2 * i2], ey = ep[1] - pos2[2 * i2 + 1];
            if (sqrtf(ex * ex + ey * ey) < 10.0f) continue;
            if (ppgo_kb8_triangulate_matches(cam8, r1, pos1 + 2 * i1, pos2 + 2 * i2, R12, t12, NULL) > 0.0001f) {
                bestIdx2 = i2;
                bestDist = dist;
            }
        }
        if (bestIdx2 >= 0) {
            match12[i1] = bestIdx2;
            nmatches++;
        }
    }
    return nmatches;
}

/* ------------------------------------------------------------------------------------------- */
/* Matcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th), matching/src/    */
/* Matcher.cpp:31-87 (tracking with the motion model, system/src/Tracking.cpp:811/817), and        */
/* Matcher::SearchByProjection(Frame &CurrentFrame, KeyFrame*, sAlreadyFound, th, descDist),        */
/* :1337-1411 (relocalisation, Tracking.cpp:1297/1311): the sequential part of both.  The rows are  */
/* the map points that passed the caller's projection tests (:38-56 / :1347-1371), in the order of  */
/* the reference's loop; proj_uv = mpCamera->project(Tcw * x3Dw).  For every row: window of radius  */
/* th (Frame::GetFeaturesInArea), best DescriptorDistance (strict <: first minimum in visiting       */
/* order) over the window features that are not occupied, accept iff best <= max_dist (TH_HIGH /    */
/* descDist) -> CurrentFrame.mvpMapPoints[bestIdx2] = pMP, which the later rows see.                */
/* Occupied (:71-73): mvpMapPoints[i2] && Observations() > 0 -- kp_mp[i2] >= 0 && observed[row],   */
/* or kp_mp[i2] == -2 (a map point outside the table with observations).  The relocalisation         */
/* variant tests mvpMapPoints[i2] alone (:1386): the caller passes observed = all ones and -2 for   */
/* every pre-assigned keypoint.  kp_mp (n) in / out; -> nmatches.                                    */
/* ------------------------------------------------------------------------------------------- */
int ppgo_search_by_projection(const ppgo_cfg *c, int n_rows, const float *map_desc, const float *proj_uv,
                              const unsigned char *observed, int n, const float *kx, const float *ky,
                              const float *frame_desc, int *kp_mp, float th, float max_dist) {
    ppgo_bounds b;
    ppgo_image_bounds(c, &b);
    int *goff = malloc(sizeof(int) * (64 * 48 + 1)), *gidx = malloc(sizeof(int) * (n > 0 ? n : 1));
    ppgo_grid_build(&b, n, kx, ky, goff, gidx);
    int *win = malloc(sizeof(int) * (n > 0 ? n : 1));
    int nmatches = 0;
    for (int r = 0; r < n_rows; r++) {
        const int nw = ppgo_features_in_area(&b, goff, gidx, kx, ky, proj_uv[2 * r], proj_uv[2 * r + 1], th, win);
        if (nw == 0) continue; /* :59-60 */
        float bestDist = 1e6f;
        int bestIdx2 = -1;
        for (int k = 0; k < nw; k++) {
            const int i2 = win[k];
            if (kp_mp[i2] == -2 || (kp_mp[i2] >= 0 && observed[kp_mp[i2]])) continue; /* :71-73 */
            const float dist = ppgo_descriptor_distance(map_desc + (size_t)r * 256, frame_desc + (size_t)i2 * 256, 256);
            if (dist < bestDist) {
                bestDist = dist;
                bestIdx2 = i2;
            }
        }
        if (bestDist <= max_dist) { /* :82 */
            kp_mp[bestIdx2] = r;
            nmatches++;
        }
    }
    free(goff);
    free(gidx);
    free(win);
    return nmatches;
}
