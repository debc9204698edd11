// ppg_shim.hpp -- header-only drop-in classes with the reference's signatures on top of the C ABI.
//
// Compile this INSIDE the PPG-SLAM tree in place of feature/src/PPGExtractor.cpp and of
// Matcher::ExtendMapMatches (matching/src/Matcher.cpp:203-381): it includes the reference's own headers for
// GeometricCamera / KeyPointEx / KeyEdge / Frame / MapPoint and needs OpenCV + Eigen exactly as they do.  It
// holds no numerical logic of its own besides translating containers <-> POD and the sequential consumption
// of the matches, which SURVEY.md s.8b keeps on the host.  (tests/test_shim_compiles.py builds it against tiny
// stand-in headers to keep it syntactically honest in a container without OpenCV/Eigen.)
#pragma once
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "ppg_b200.h"

#ifndef PPG_SHIM_NO_REFERENCE_HEADERS
#include "Frame.h"            // map/include
#include "GeometricCamera.h"  // sensors/include: KeyPointEx, GeometricCamera
#include "MapPoint.h"         // feature/include
#include "PPGGraph.h"         // feature/include: KeyEdge
#endif

namespace ppg_shim {

inline void check(int rc, ppg_ctx* ctx, const char* what) {
    if (rc != PPG_OK && rc != PPG_ERR_CAPACITY)
        throw std::runtime_error(std::string(what) + ": " + ppg_last_error(ctx));
}

// Replaces class PPGExtractor (feature/include/PPGExtractor.h:34-148).  Same constructor arguments, same run()
// signature and semantics (PPGExtractor.cpp:118-147), same public members / static tunables; the LibTorch
// member `normDesc` (unused outside the class) is gone.
class PPGExtractor {
public:
    PPGExtractor(GeometricCamera* pCam, std::string dataPath) {
        ppg_config cfg;
        ppg_default_config(&cfg);
        cfg.width = pCam->imWidth();
        cfg.height = pCam->imHeight();
        cv::Mat K = pCam->toK(), D = pCam->toD();  // PPGExtractor.cpp:58-62
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) cfg.K[3 * r + c] = K.at<float>(r, c);
        for (int i = 0; i < 4; i++) cfg.D[i] = D.at<float>(i);
        cfg.fisheye = pCam->mnType == GeometricCamera::CAM_FISHEYE;
        mWeights = dataPath + "/ppg_weights.bin";  // tools/export_weights.py output next to net/*.pt
        cfg.weights_path = mWeights.c_str();
        cfg.junction_thresh = JUNCTION_THRESH;
        cfg.junction_nms_radius = JUNCTION_NMS_RADIUS;
        cfg.junction_max_num = (int)JUNCTION_MAX_NUM;
        cfg.line_valid_thresh = LINE_VALID_THRESH;
        cfg.line_valid_ratio = LINE_VALID_RATIO;
        cfg.line_dist_thresh = LINE_DISTTHRESH;
        cfg.heatmap_refine_sz = HEATMAP_REFINE_SZ;
        cfg.line_heatmap_thresh = LINE_HEATMAP_THRESH;
        cfg.line_inlier_rate = LINE_INLIER_RATE;
        cfg.max_batch = 1;  // the SLAM front-end feeds one frame at a time (map/src/Frame.cpp:62)
        int rc = ppg_create(&cfg, &mCtx);
        if (rc != PPG_OK) throw std::runtime_error(std::string("ppg_create: ") + ppg_last_error(nullptr));
    }
    ~PPGExtractor() { ppg_destroy(mCtx); }
    PPGExtractor(const PPGExtractor&) = delete;
    PPGExtractor& operator=(const PPGExtractor&) = delete;

    void run(cv::Mat srcMat, std::vector<KeyPointEx>& _keypoints, std::vector<KeyPointEx>& _keypoints_un,
             std::vector<KeyEdge>& _keyedges, cv::Mat& _descriptors) {
        if (srcMat.channels() != 1) throw std::invalid_argument("PPGExtractor::run: single-channel image expected");
        const uint8_t* ptr = srcMat.data;
        const int stride = (int)srcMat.step;
        ppg_frame_out o;
        check(ppg_extract(mCtx, &ptr, &stride, 1, &o), mCtx, "ppg_extract");
        mvKeyPoints.clear();
        mvKeyPoints.reserve(o.n_kp);
        for (int i = 0; i < o.n_kp; i++) {
            KeyPointEx kp((float)o.kp_px[i], (float)o.kp_py[i], o.kp_score[i]);
            kp.mPosUn << o.kp_xun[i], o.kp_yun[i];
            kp.mbOut = o.kp_out[i] != 0;
            kp.mvConnected.assign(o.conn_idx + o.conn_off[i], o.conn_idx + o.conn_off[i + 1]);
            for (int k = o.col_off[i]; k < o.col_off[i + 1]; k++)
                kp.mvColine.emplace_back((unsigned)o.col_pairs[2 * k], (unsigned)o.col_pairs[2 * k + 1]);
            mvKeyPoints.push_back(kp);
        }
        mvKeyEdges.clear();
        mvKeyEdges.reserve(o.n_edges);
        for (int e = 0; e < o.n_edges; e++) {
            KeyEdge ke((unsigned)o.edge_start[e], (unsigned)o.edge_end[e]);
            ke.isBad = false;
            ke.lscore = o.edge_score[e];
            mvKeyEdges.push_back(ke);
        }
        _keypoints = mvKeyPoints;
        _keyedges = mvKeyEdges;
        _descriptors = cv::Mat(o.n_kp, PPG_DESC_DIM, CV_32FC1);
        for (int i = 0; i < o.n_kp; i++)
            std::memcpy(_descriptors.ptr<float>(i), o.desc + (size_t)i * PPG_DESC_DIM, PPG_DESC_DIM * sizeof(float));
        for (int i = 0; i < o.n_kp; i++) _keypoints[i].mPos << o.kp_x[i], o.kp_y[i];  // :141-145 (pinhole: mPosUn)
        _keypoints_un = _keypoints;                                                   // :146
    }

    const std::vector<KeyPointEx>& getKeyPoints() const { return mvKeyPoints; }
    const std::vector<KeyEdge>& getKeyEdges() const { return mvKeyEdges; }
    ppg_ctx* context() { return mCtx; }

public:
    std::vector<KeyPointEx> mvKeyPoints;
    std::vector<KeyEdge> mvKeyEdges;
    // static tunables, defaults of PPGExtractor.cpp:44-53 (set before constructing)
    static inline int DESC_DIM_SIZE = 256;
    static inline float JUNCTION_THRESH = 1.0f / 128.0f;
    static inline int JUNCTION_NMS_RADIUS = 4;
    static inline unsigned int JUNCTION_MAX_NUM = 500;
    static inline float LINE_VALID_THRESH = 1.0e-2f;
    static inline float LINE_VALID_RATIO = 0.3f;
    static inline float LINE_DISTTHRESH = 2.0f;
    static inline int HEATMAP_REFINE_SZ = 16;
    static inline float LINE_HEATMAP_THRESH = 0.2f;
    static inline float LINE_INLIER_RATE = 0.8f;

private:
    ppg_ctx* mCtx = nullptr;
    std::string mWeights;
};

// The data-parallel part of Matcher::ExtendMapMatches (Matcher.cpp:224-281): best / second-best keypoint of
// every candidate map point over its search window, in one GPU call with the frame state frozen.  The caller
// (the reference's ExtendMapMatches body) then walks the map points in its sorted order: a row whose best and
// second-best keypoints are both still unmatched keeps the GPU answer (it is what the sequential loop would
// find: removing OTHER candidates cannot change the two smallest); a row that lost one of them to an earlier
// map point is re-evaluated with the reference's own loop (Matcher.cpp:251-270) -- rare, and exact.
struct SearchResult {
    std::vector<int32_t> best_idx, second_idx;
    std::vector<float> best_dist, second_dist;
    std::vector<uint8_t> accept;
};

inline void upload_map_descriptors(ppg_ctx* ctx, const std::vector<MapPoint*>& mps) {
    std::vector<float> table(mps.size() * (size_t)PPG_DESC_DIM);
    for (size_t m = 0; m < mps.size(); m++) {
        const cv::Mat d = mps[m]->GetDescriptor();  // feature/src/MapPoint.cpp:304-308
        std::memcpy(&table[m * PPG_DESC_DIM], d.ptr<float>(0), PPG_DESC_DIM * sizeof(float));
    }
    check(ppg_upload_map(ctx, table.data(), (int)mps.size()), ctx, "ppg_upload_map");
}

inline SearchResult search_local_points(ppg_ctx* ctx, Frame& F, const std::vector<MapPoint*>& mps, float th,
                                        float nnratio) {
    const int N = (int)F.mvKeysUn.size(), M = (int)mps.size();
    std::vector<float> kx(N), ky(N), uv(2 * (size_t)M), vc(M);
    std::vector<uint8_t> free_mask(N);
    for (int i = 0; i < N; i++) {
        kx[i] = F.mvKeysUn[i].mPos[0];
        ky[i] = F.mvKeysUn[i].mPos[1];
        free_mask[i] = !(F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0);  // Matcher.cpp:253
    }
    for (int m = 0; m < M; m++) {
        uv[2 * m] = mps[m]->mTrackProjX;
        uv[2 * m + 1] = mps[m]->mTrackProjY;
        vc[m] = mps[m]->mTrackViewCos;
    }
    SearchResult r;
    r.best_idx.resize(M);
    r.second_idx.resize(M);
    r.best_dist.resize(M);
    r.second_dist.resize(M);
    r.accept.resize(M);
    ppg_assoc_in in{};  // mode = PPG_SEARCH_EXTEND_MAP
    in.n_kp = N;
    in.kp_x = kx.data();
    in.kp_y = ky.data();
    in.frame_desc = F.mDescriptors.ptr<float>(0);
    in.free_mask = free_mask.data();
    in.n_rows = M;
    in.proj_uv = uv.data();
    in.view_cos = vc.data();
    in.th = th;
    in.ratio = nnratio;
    ppg_assoc_out out{r.best_idx.data(), r.second_idx.data(), r.best_dist.data(), r.second_dist.data(),
                      r.accept.data()};
    check(ppg_associate(ctx, &in, &out), ctx, "ppg_associate");
    return r;
}

// The best-only projection cores share the same primitive with r = th and accept = best <= max_dist
// (PPG_SEARCH_WINDOW).  The caller keeps its own projection / visibility filters and passes the surviving rows:
//   Matcher::SearchByProjection(Cur, Last, th)               Matcher.cpp:31-87     max_dist = TH_HIGH,
//        free_mask[i] = !(Cur.mvpMapPoints[i] && Cur.mvpMapPoints[i]->Observations() > 0)               (:66-68)
//   Matcher::SearchByProjection(F, pKF, sFound, th, descDist) Matcher.cpp:1337-1411 max_dist = descDist,
//        free_mask[i] = !F.mvpMapPoints[i]                                                               (:1386)
//   Matcher::Fuse(pKF, vpMapPoints, th)                       Matcher.cpp:897-1036  max_dist = TH_LOW, e2_max = 5.99,
//        free_mask all ones (Fuse looks at every keypoint of the window, :994-1013)
//   Matcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming)   Matcher.cpp:479-568
//        max_dist = TH_LOW * ratioHamming, free_mask[i] = !vpMatched[i]
//   Matcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint)     Matcher.cpp:1038-1135 max_dist = TH_LOW, free_mask all ones
//   Matcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th)   Matcher.cpp:1149-1335 one call per direction (map points of
//        KF1 against the keypoints of KF2 and vice versa), max_dist = TH_HIGH, free_mask all ones; the mutual-consistency
//        check (:1316-1330) stays on the host
// `FrameLike` is Frame or KeyFrame (mvKeysUn, mDescriptors).  As in search_local_points the sequential consumption
// (a keypoint taken by an earlier row) is resolved by the caller with the reference's own inner loop.
template <class FrameLike>
inline SearchResult search_window(ppg_ctx* ctx, FrameLike& F, const std::vector<MapPoint*>& mps,
                                  const std::vector<float>& proj_uv, const std::vector<uint8_t>& free_mask, float th,
                                  float max_dist, double e2_max = 0.0) {
    const int N = (int)F.mvKeysUn.size(), M = (int)mps.size();
    std::vector<float> kx(N), ky(N), vc(M, 0.f);
    for (int i = 0; i < N; i++) {
        kx[i] = F.mvKeysUn[i].mPos[0];
        ky[i] = F.mvKeysUn[i].mPos[1];
    }
    SearchResult r;
    r.best_idx.resize(M);
    r.second_idx.resize(M);
    r.best_dist.resize(M);
    r.second_dist.resize(M);
    r.accept.resize(M);
    ppg_assoc_in in{};
    in.n_kp = N;
    in.kp_x = kx.data();
    in.kp_y = ky.data();
    in.frame_desc = F.mDescriptors.template ptr<float>(0);
    in.free_mask = free_mask.data();
    in.n_rows = M;
    in.proj_uv = proj_uv.data();
    in.view_cos = vc.data();
    in.th = th;
    in.mode = PPG_SEARCH_WINDOW;
    in.max_dist = max_dist;
    in.e2_max = e2_max;
    ppg_assoc_out out{r.best_idx.data(), r.second_idx.data(), r.best_dist.data(), r.second_dist.data(),
                      r.accept.data()};
    upload_map_descriptors(ctx, mps);
    check(ppg_associate(ctx, &in, &out), ctx, "ppg_associate");
    return r;
}

}  // namespace ppg_shim
