// ppg_shim.hpp -- header-only drop-in classes with the reference's signatures on top of the C ABI.
//
// Compile this INSIDE the PPG-SLAM tree in place of feature/src/PPGExtractor.cpp and of
// Matcher::ExtendMapMatches (matching/src/Matcher.cpp:203-381): it includes the reference's own headers for
// GeometricCamera / KeyPointEx / KeyEdge / Frame / MapPoint and needs OpenCV + Eigen exactly as they do.  It
// holds no numerical logic of its own: it translates containers / the pointer graph <-> POD.  ExtendMapMatches runs
// whole on the GPU (extend_map_matches below); search_local_points / search_window expose the frozen-state search
// core for the matchers that keep their own sequential loops.  (tests/test_shim_compiles.py builds it against tiny
// stand-in headers to keep it syntactically honest in a container without OpenCV/Eigen.)
#pragma once
#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <map>
#include <string>
#include <set>
#include <unordered_map>
#include <vector>

#include "ppg_b200.h"

#ifndef PPG_SHIM_NO_REFERENCE_HEADERS
#include "Frame.h"            // map/include
#include "GeometricCamera.h"  // sensors/include: KeyPointEx, GeometricCamera
#include "MapPoint.h"         // feature/include
#include "PPGGraph.h"         // feature/include: KeyEdge
#include "Matcher.h"          // matching/include: the base of ppg_shim::Matcher (host-side matchers stay there)
#endif

namespace ppg_shim {

// Every C-ABI failure throws -- PPG_ERR_CAPACITY included: a truncated result must never reach the SLAM caller
// silently.  The two places where a capacity overflow has a defined meaning handle the code themselves:
// PPGExtractor::run (records a frame status) and extend_map_matches (falls back to the caller's CPU loop).
inline void check(int rc, ppg_ctx* ctx, const char* what) {
    if (rc != PPG_OK) throw std::runtime_error(std::string(what) + ": " + ppg_last_error(ctx));
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
        const int rc = ppg_extract(mCtx, &ptr, &stride, 1, &o);
        if (rc != PPG_ERR_CAPACITY) check(rc, mCtx, "ppg_extract");
        // The reference has no capacities.  When one of the library's was exceeded (o.status: PPG_STATUS_* bits --
        // accepted candidates, candidate pairs, per-keypoint degree, max_edges, max_colines) the record is a truncated
        // graph: it is still handed out (a subset of the reference's edges), the status is kept for the caller to
        // inspect, and the first occurrence is reported.  Raise the ppg_config capacities if it ever shows up.
        mLastStatus = o.status;
        if (o.status && !mWarned) {
            std::fprintf(stderr, "ppg_shim::PPGExtractor::run: frame capacity exceeded (status 0x%x): %s\n", o.status,
                         ppg_last_error(mCtx));
            mWarned = true;
        }
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
    unsigned lastStatus() const { return mLastStatus; }  // ppg_frame_out.status of the last run(): 0 = nothing truncated

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
    unsigned mLastStatus = 0;
    bool mWarned = false;
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

// MSTracking::SearchLocalPoints (system/src/Tracking.cpp:978-1008) = Frame::CheckInFrustum over the local map +
// ExtendMapMatches.  The first half on the GPU: map geometry uploaded once per local-map update, one pose per frame.
// Writes mbTrackInView / mTrackProjX / mTrackProjY / mTrackDepth / mTrackViewCos back and calls IncreaseVisible()
// (map/src/Frame.cpp:252-259) so that code reading those members keeps working.  `rows` = the map points of the
// resident table, in table order.
inline void upload_map_geometry(ppg_ctx* ctx, const std::vector<MapPoint*>& rows) {
    const size_t M = rows.size();
    std::vector<float> P(3 * M), Nn(3 * M), dmin(M), dmax(M);
    for (size_t m = 0; m < M; m++) {
        const auto p = rows[m]->GetWorldPos();
        const auto n = rows[m]->GetNormal();
        for (int k = 0; k < 3; k++) {
            P[3 * m + k] = p[k];
            Nn[3 * m + k] = n[k];
        }
        dmin[m] = rows[m]->GetMinDistanceInvariance();
        dmax[m] = rows[m]->GetMaxDistanceInvariance();
    }
    check(ppg_upload_map_geometry(ctx, P.data(), Nn.data(), dmin.data(), dmax.data(), (int)M), ctx,
          "ppg_upload_map_geometry");
}

inline void check_in_frustum(ppg_ctx* ctx, Frame& F, const std::vector<MapPoint*>& rows, float viewingCosLimit,
                             float th, float nnratio) {
    float R[9], t[3], O[3];
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) R[3 * r + c] = F.mRcw(r, c);
        t[r] = F.mtcw[r];
        O[r] = F.mOw[r];
    }
    const int M = (int)rows.size();
    check(ppg_assoc_stage_poses(ctx, 1, M, R, t, O, viewingCosLimit, th, nnratio), ctx, "ppg_assoc_stage_poses");
    std::vector<uint8_t> in_view(M);
    std::vector<float> uv(2 * (size_t)M), depth(M), vc(M);
    check(ppg_frustum_fetch(ctx, 1, in_view.data(), uv.data(), depth.data(), vc.data()), ctx, "ppg_frustum_fetch");
    for (int m = 0; m < M; m++) {
        MapPoint* p = rows[m];
        p->mbTrackInView = in_view[m] != 0;
        p->mTrackProjX = uv[2 * m];
        p->mTrackProjY = uv[2 * m + 1];
        p->mTrackDepth = depth[m];
        if (in_view[m]) {
            p->mTrackViewCos = vc[m];
            p->IncreaseVisible();
        }
    }
}

// Frame::ComputeBoW (map/src/Frame.cpp:331-340): mpVoc->transform(descriptors, mBowVec, mFeatVec, 4) on the GPU.  The
// vocabulary is uploaded once (ppg_upload_vocabulary, from the blob tools/export_vocabulary.py writes next to the .gz).
// `FrameLike` is Frame or KeyFrame (mDescriptors, mBowVec, mFeatVec).
template <class FrameLike>
inline void compute_bow(ppg_ctx* ctx, FrameLike& F, int levelsup = 4) {
    if (!F.mBowVec.empty()) return;
    const int N = F.mDescriptors.rows;
    std::vector<int32_t> word(N > 0 ? N : 1), node(N > 0 ? N : 1), bword(N > 0 ? N : 1);
    std::vector<double> weight(N > 0 ? N : 1), bval(N > 0 ? N : 1);
    ppg_bow_out out{};
    out.word_id = word.data();
    out.word_weight = weight.data();
    out.node_id = node.data();
    out.bow_word = bword.data();
    out.bow_value = bval.data();
    check(ppg_bow_transform(ctx, F.mDescriptors.template ptr<float>(0), N, levelsup, &out), ctx, "ppg_bow_transform");
    for (int j = 0; j < out.n_bow; j++) F.mBowVec[(unsigned int)bword[j]] = bval[j];          // BowVector = map<WordId, WordValue>
    for (int i = 0; i < out.n_features; i++)
        if (node[i] >= 0) F.mFeatVec[(unsigned int)node[i]].push_back((unsigned int)i);       // FeatureVector::addFeature
}

// Matcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (matching/src/Matcher.cpp:393-477) whole on the GPU
// (ppg_search_by_bow): rows = the keyframe features that hold a good map point, in the order the reference visits them
// (FeatureVector node ascending, then feature index); the frame side gives its descriptors and the node of every
// listed feature.  `KeyFrameLike` needs GetMapPointMatches(), mFeatVec, mDescriptors.
template <class KeyFrameLike>
inline int search_by_bow(ppg_ctx* ctx, KeyFrameLike* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches,
                         float nnratio, float th_low) {
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    const int N = (int)F.mvKeysUn.size();
    vpMapPointMatches = std::vector<MapPoint*>(N, static_cast<MapPoint*>(nullptr));  // :396
    std::vector<int32_t> row_node, kp_node(N, -1);
    std::vector<MapPoint*> row_mp;
    std::vector<float> table;
    for (const auto& nf : pKF->mFeatVec)  // std::map: ascending node id
        for (unsigned int idx : nf.second) {
            MapPoint* pMP = vpMapPointsKF[idx];
            if (!pMP || pMP->isBad()) continue;  // :423-427
            row_node.push_back((int32_t)nf.first);
            row_mp.push_back(pMP);
            const float* d = pKF->mDescriptors.template ptr<float>((int)idx);
            table.insert(table.end(), d, d + PPG_DESC_DIM);
        }
    if (row_node.empty()) return 0;
    for (const auto& nf : F.mFeatVec)
        for (unsigned int idx : nf.second) kp_node[idx] = (int32_t)nf.first;
    check(ppg_upload_map(ctx, table.data(), (int)row_node.size()), ctx, "ppg_upload_map");
    std::vector<int32_t> kp_row(N > 0 ? N : 1, -1);
    ppg_bow_match_in in{};
    in.n_rows = (int)row_node.size();
    in.row_node = row_node.data();
    in.n_kp = N;
    in.frame_desc = F.mDescriptors.template ptr<float>(0);
    in.kp_node = kp_node.data();
    in.ratio = nnratio;
    in.max_dist = th_low;
    in.strict = 0;  // :456 `<=`; SearchByBoW(KF1, KF2) :733 uses `<` (strict = 1)
    ppg_bow_match_out out{};
    out.kp_row = kp_row.data();
    check(ppg_search_by_bow(ctx, &in, &out), ctx, "ppg_search_by_bow");
    for (int i = 0; i < N; i++)
        if (kp_row[i] >= 0) vpMapPointMatches[i] = row_mp[kp_row[i]];  // :460
    return out.nmatches;
}

// Matcher::SearchByBoW(KeyFrame*, KeyFrame*, ...) (matching/src/Matcher.cpp:663-754) whole on the GPU: the same call with
// strict = 1 (`bestDist1 < TH_LOW`, :733); rows = the features of KF1 with a good map point in FeatureVector order, the
// candidates of a row = the features of KF2 under the same node with a good map point that no earlier row has taken
// (vbMatched2, :709).  vpMatches12[i1] = the map point of the matched KF2 feature.
template <typename KeyFrameLike>
inline int search_by_bow(ppg_ctx* ctx, KeyFrameLike* pKF1, KeyFrameLike* pKF2, std::vector<MapPoint*>& vpMatches12,
                         float nnratio, float th_low) {
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const std::vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    vpMatches12 = std::vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(nullptr));  // :675
    const int n2 = (int)vpMapPoints2.size();
    std::vector<int32_t> row_node, row_feat, kp_node(n2 > 0 ? n2 : 1, -1);
    std::vector<float> table;
    for (const auto& nf : pKF1->mFeatVec)
        for (unsigned int idx : nf.second) {
            MapPoint* pMP = vpMapPoints1[idx];
            if (!pMP || pMP->isBad()) continue;  // :693-697
            row_node.push_back((int32_t)nf.first);
            row_feat.push_back((int32_t)idx);
            const float* d = pKF1->mDescriptors.template ptr<float>((int)idx);
            table.insert(table.end(), d, d + PPG_DESC_DIM);
        }
    if (row_node.empty() || n2 == 0) return 0;
    for (const auto& nf : pKF2->mFeatVec)
        for (unsigned int idx : nf.second) {
            MapPoint* pMP = vpMapPoints2[idx];
            if (pMP && !pMP->isBad()) kp_node[idx] = (int32_t)nf.first;  // :707-714
        }
    check(ppg_upload_map(ctx, table.data(), (int)row_node.size()), ctx, "ppg_upload_map");
    std::vector<int32_t> kp_row(n2, -1);
    ppg_bow_match_in in{};
    in.n_rows = (int)row_node.size();
    in.row_node = row_node.data();
    in.n_kp = n2;
    in.frame_desc = pKF2->mDescriptors.template ptr<float>(0);
    in.kp_node = kp_node.data();
    in.ratio = nnratio;
    in.max_dist = th_low;
    in.strict = 1;
    ppg_bow_match_out out{};
    out.kp_row = kp_row.data();
    check(ppg_search_by_bow(ctx, &in, &out), ctx, "ppg_search_by_bow");
    for (int i2 = 0; i2 < n2; i2++)
        if (kp_row[i2] >= 0) vpMatches12[row_feat[kp_row[i2]]] = vpMapPoints2[i2];  // :737
    return out.nmatches;
}

// The whole Matcher::ExtendMapMatches (matching/src/Matcher.cpp:203-381) in one GPU call: window search with the live
// frame state, assignment and seed growing all run in ppg_extend_map_matches; this function only flattens the
// pointer graph into the POD form of include/ppg_b200.h and writes the result back into the Frame:
//   table rows = the trackable map points (:210-215) in vpMapPoints order, then every map point reachable as
//   theOtherPt() of one of their edges (its descriptor is read at :330), then the map points the frame already holds
//   (needed for the pMP_o == F.mvpMapPoints[keyID_o] test at :327).
// `cpu_fallback` (optional) is the reference's own loop: it runs instead when the frame exceeds a device capacity (a map
// point / keypoint with more than PPG_EXTEND_MAX_DEGREE edges, a seed needing more than PPG_EXTEND_MAX_WEIGHTS
// weights) -- nothing has been written into F at that point; without it such a frame throws.
// ppg_shim::Matcher::ExtendMapMatches below passes ::Matcher::ExtendMapMatches.
inline int extend_map_matches(ppg_ctx* ctx, Frame& F, const std::vector<MapPoint*>& vpMapPoints, float th, float nnratio,
                              const std::function<int()>& cpu_fallback = nullptr) {
    std::vector<MapPoint*> rows;
    std::unordered_map<MapPoint*, int32_t> row_of;
    std::vector<uint8_t> candidate;
    auto add_row = [&](MapPoint* p, bool cand) -> int32_t {
        auto it = row_of.find(p);
        if (it != row_of.end()) return it->second;
        const int32_t r = (int32_t)rows.size();
        row_of.emplace(p, r);
        rows.push_back(p);
        candidate.push_back(cand ? 1 : 0);
        return r;
    };
    for (MapPoint* pMP : vpMapPoints)
        if (!(pMP->isBad() || !pMP->mbTrackInView)) add_row(pMP, true);  // :210-215
    const size_t n_cand = rows.size();
    if (n_cand == 0) return 0;  // no trackable map point: the reference's loop body never runs (Matcher.cpp:226)
    std::vector<std::vector<MapEdge*>> edges(n_cand);
    for (size_t r = 0; r < n_cand; r++) {
        edges[r] = rows[r]->getEdges();
        for (MapEdge* e : edges[r])
            if (MapPoint* o = e->theOtherPt(rows[r])) add_row(o, false);
    }
    const int N = (int)F.mvKeysUn.size();
    std::vector<int32_t> kp_mp(N, -1);
    for (int i = 0; i < N; i++)
        if (MapPoint* q = F.mvpMapPoints[i]) kp_mp[i] = add_row(q, false);
    const int P = (int)rows.size();
    std::vector<uint8_t> observed(P), bad(P), tracked(P), edge_ok;
    std::vector<int32_t> edge_off(P + 1, 0), edge_other;
    std::vector<MapEdge*> edge_ptr;
    std::vector<float> uv(2 * (size_t)P, 0.f), vc(P, 0.f);
    for (int r = 0; r < P; r++) {
        MapPoint* p = rows[r];
        observed[r] = p->Observations() > 0;                 // :253
        bad[r] = p->isBad();                                 // :227, :364
        tracked[r] = p->mnTrackedbyFrame == F.mnId;          // :227, :364
        uv[2 * r] = p->mTrackProjX;
        uv[2 * r + 1] = p->mTrackProjY;
        vc[r] = p->mTrackViewCos;
        if ((size_t)r < n_cand)
            for (MapEdge* e : edges[r]) {
                MapPoint* o = e->theOtherPt(p);
                edge_other.push_back(o ? row_of[o] : -1);
                edge_ok.push_back(!(e->isBad() || !e->mbValid));  // :311
                edge_ptr.push_back(e);
            }
        edge_off[r + 1] = (int32_t)edge_other.size();
    }
    upload_map_descriptors(ctx, rows);
    ppg_map_graph g{P, candidate.data(), observed.data(), bad.data(), edge_off.data(), edge_other.data(),
                    edge_ok.data()};
    check(ppg_upload_map_graph(ctx, &g), ctx, "ppg_upload_map_graph");
    const int E = (int)F.mvKeyEdges.size();
    std::vector<float> kx(N), ky(N);
    std::vector<int32_t> es(E), ee(E), conn_off(N + 1, 0), conn_idx, kedge_me(E, -1);
    for (int i = 0; i < N; i++) {
        kx[i] = F.mvKeysUn[i].mPos[0];
        ky[i] = F.mvKeysUn[i].mPos[1];
        for (unsigned int e : F.mvKeysUn[i].mvConnected) conn_idx.push_back((int32_t)e);
        conn_off[i + 1] = (int32_t)conn_idx.size();
    }
    for (int e = 0; e < E; e++) {
        es[e] = (int32_t)F.mvKeyEdges[e].startIdx;
        ee[e] = (int32_t)F.mvKeyEdges[e].endIdx;
    }
    ppg_extend_in in{};
    in.n_kp = N;
    in.kp_x = kx.data();
    in.kp_y = ky.data();
    in.frame_desc = F.mDescriptors.ptr<float>(0);
    in.kp_mp = kp_mp.data();
    in.n_edges = E;
    in.edge_start = es.data();
    in.edge_end = ee.data();
    in.conn_off = conn_off.data();
    in.conn_idx = conn_idx.data();
    in.kedge_me = nullptr;  // only written, never read, by the reference (:367)
    in.proj_uv = uv.data();
    in.view_cos = vc.data();
    in.tracked = tracked.data();
    in.th = th;
    in.ratio = nnratio;
    std::vector<int32_t> out_kp(N > 0 ? N : 1), out_ke(E > 0 ? E : 1);
    std::vector<uint8_t> out_tr(P);
    ppg_extend_out out{};
    out.kp_mp = out_kp.data();
    out.kedge_me = out_ke.data();
    out.tracked = out_tr.data();
    const int rc = ppg_extend_map_matches(ctx, &in, &out);
    if (rc == PPG_ERR_CAPACITY && cpu_fallback) return cpu_fallback();  // F untouched so far
    check(rc, ctx, "ppg_extend_map_matches");
    for (int i = 0; i < N; i++)
        if (out_kp[i] != kp_mp[i]) F.mvpMapPoints[i] = rows[out_kp[i]];          // :279, :366
    for (int e = 0; e < E; e++)
        if (out_ke[e] >= 0) F.mvpMapEdges[e] = edge_ptr[out_ke[e]];              // :367
    for (int r = 0; r < P; r++)
        if (out_tr[r] && !tracked[r]) rows[r]->mnTrackedbyFrame = F.mnId;        // :280, :368
    return out.nmatches;
}

// Matcher::SearchForInitialization (matching/src/Matcher.cpp:582-651) whole on the GPU (ppg_search_for_initialization).
inline int search_for_initialization(ppg_ctx* ctx, Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched,
                                     std::vector<int>& vnMatches12, int windowSize, float nnratio) {
    const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
    vnMatches12 = std::vector<int>(n1, -1);  // :585
    if (n1 == 0 || n2 == 0) return 0;
    check(ppg_upload_map(ctx, F1.mDescriptors.ptr<float>(0), n1), ctx, "ppg_upload_map");
    std::vector<float> prev(2 * (size_t)n1), kx(n2), ky(n2);
    for (int i = 0; i < n1; i++) {
        prev[2 * i] = vbPrevMatched[i].x;
        prev[2 * i + 1] = vbPrevMatched[i].y;
    }
    for (int i = 0; i < n2; i++) {
        kx[i] = F2.mvKeysUn[i].mPos[0];
        ky[i] = F2.mvKeysUn[i].mPos[1];
    }
    std::vector<int32_t> m12(n1, -1);
    ppg_init_match_in in{};
    in.n1 = n1;
    in.prev_matched = prev.data();
    in.n2 = n2;
    in.kp2_x = kx.data();
    in.kp2_y = ky.data();
    in.desc2 = F2.mDescriptors.ptr<float>(0);
    in.window = windowSize;
    in.ratio = nnratio;
    ppg_init_match_out out{};
    out.matches12 = m12.data();
    out.prev_matched = prev.data();
    check(ppg_search_for_initialization(ctx, &in, &out), ctx, "ppg_search_for_initialization");
    for (int i = 0; i < n1; i++) {
        vnMatches12[i] = m12[i];
        if (m12[i] >= 0) vbPrevMatched[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]);  // :644-647
    }
    return out.nmatches;
}

// Matcher::SearchByProjection(CurrentFrame, LastFrame, th) (matching/src/Matcher.cpp:31-87; every frame tracked with the
// motion model, system/src/Tracking.cpp:811) and SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, descDist)
// (:1337-1411; relocalisation) whole on the GPU (ppg_search_by_projection).  The projection tests of the reference's loop
// (:38-56 / :1347-1371) run here with the reference's own classes and expressions, in loop order; the map points that pass
// are the rows.  The sequential part -- window, best free keypoint, accept, occupy -- runs on the device.
namespace detail {
// keys / descriptors / slots: mvKeysUn, mDescriptors and the map-point vector (F.mvpMapPoints or vpMatched) of the frame
// or key frame that is searched
inline int projection_match(ppg_ctx* ctx, int N, const std::vector<KeyPointEx>& keys, const cv::Mat& descriptors,
                            std::vector<MapPoint*>& slots, const std::vector<MapPoint*>& rows,
                            const std::vector<float>& proj_uv, bool pointer_occupies, float th, float max_dist) {
    const int M = (int)rows.size();
    if (M == 0 || N <= 0) return 0;
    std::vector<float> table((size_t)M * 256), kx(N), ky(N);
    std::vector<uint8_t> observed(M);
    std::unordered_map<MapPoint*, int> row_of;
    for (int r = 0; r < M; r++) {
        const cv::Mat d = rows[r]->GetDescriptor();
        std::memcpy(&table[(size_t)r * 256], d.ptr<float>(0), 1024);
        observed[r] = rows[r]->Observations() > 0 ? 1 : 0;
        row_of.emplace(rows[r], r);  // a map point listed twice keeps its first row (it cannot be tracked twice anyway)
    }
    std::vector<int32_t> kp_mp(N, -1), out_kp(N, -1);
    for (int i = 0; i < N; i++) {
        kx[i] = keys[i].mPos[0];
        ky[i] = keys[i].mPos[1];
        MapPoint* m = slots[i];
        if (!m) continue;
        if (pointer_occupies) {  // :1386 tests the pointer alone
            kp_mp[i] = -2;
            continue;
        }
        auto it = row_of.find(m);
        kp_mp[i] = it != row_of.end() ? it->second : (m->Observations() > 0 ? -2 : -1);  // :71-73
    }
    check(ppg_upload_map(ctx, table.data(), M), ctx, "ppg_upload_map");
    ppg_projection_match_in in{};
    in.n_rows = M;
    in.proj_uv = proj_uv.data();
    in.observed = pointer_occupies ? nullptr : observed.data();
    in.n = N;
    in.kp_x = kx.data();
    in.kp_y = ky.data();
    in.desc = descriptors.ptr<float>(0);
    in.kp_mp = kp_mp.data();
    in.th = th;
    in.max_dist = max_dist;
    ppg_projection_match_out out{};
    out.kp_mp = out_kp.data();
    check(ppg_search_by_projection(ctx, &in, &out), ctx, "ppg_search_by_projection");
    for (int i = 0; i < N; i++)
        if (out_kp[i] >= 0 && out_kp[i] != kp_mp[i]) slots[i] = rows[out_kp[i]];  // :84 / :1403 / :561
    return out.nmatches;
}
}  // namespace detail

inline int search_by_projection(ppg_ctx* ctx, Frame& CurrentFrame, const Frame& LastFrame, float th, float th_high) {
    const SE3f Tcw = CurrentFrame.GetPose();
    std::vector<MapPoint*> rows;
    std::vector<float> uvs;
    for (int i = 0; i < LastFrame.N; i++) {
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;  // :38-41
        Eigen::Vector3f x3Dw = pMP->GetWorldPos();
        Eigen::Vector3f x3Dc = Tcw * x3Dw;
        const float invzc = 1.0 / x3Dc(2);
        if (invzc < 0) continue;  // :49-50
        Eigen::Vector2f uv = CurrentFrame.mpCamera->project(x3Dc);
        if (!CurrentFrame.mpCamera->IsInImage(uv(0), uv(1))) continue;  // :54-55
        rows.push_back(pMP);
        uvs.push_back(uv(0));
        uvs.push_back(uv(1));
    }
    return detail::projection_match(ctx, CurrentFrame.N, CurrentFrame.mvKeysUn, CurrentFrame.mDescriptors,
                                    CurrentFrame.mvpMapPoints, rows, uvs, false, th, th_high);
}

inline int search_by_projection(ppg_ctx* ctx, Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound,
                                float th, float descDist) {
    const SE3f Tcw = CurrentFrame.GetPose();
    Eigen::Vector3f Ow = Tcw.inverse().translation();
    const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    std::vector<MapPoint*> rows;
    std::vector<float> uvs;
    for (size_t i = 0, iend = vpMPs.size(); i < iend; i++) {
        MapPoint* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;  // :1350-1352
        Eigen::Vector3f x3Dw = pMP->GetWorldPos();
        Eigen::Vector3f x3Dc = Tcw * x3Dw;
        const Eigen::Vector2f uv = CurrentFrame.mpCamera->project(x3Dc);
        if (!CurrentFrame.mpCamera->IsInImage(uv(0), uv(1))) continue;  // :1360-1361
        Eigen::Vector3f PO = x3Dw - Ow;
        float dist3D = PO.norm();
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        if (dist3D < minDistance || dist3D > maxDistance) continue;  // :1370-1371
        rows.push_back(pMP);
        uvs.push_back(uv(0));
        uvs.push_back(uv(1));
    }
    return detail::projection_match(ctx, CurrentFrame.N, CurrentFrame.mvKeysUn, CurrentFrame.mDescriptors,
                                    CurrentFrame.mvpMapPoints, rows, uvs, true, th, descDist);
}

// Matcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (matching/src/Matcher.cpp:479-568; loop
// closing, system/src/LoopClosing.cpp:585 / :614 / :792) whole: the same sequential structure over a key frame's
// keypoints -- a key-frame feature that holds a match is skipped (:543-544), an accepted point occupies its feature
// (:561).  The loop head (:491-525: bad / already found, depth, image, distance band, viewing angle) runs here.
inline int search_by_projection(ppg_ctx* ctx, GeometricCamera* cam, KeyFrame* pKF, Sim3f& Scw,
                                const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th,
                                float ratioHamming, float th_low) {
    SE3f Tcw = SE3f(Scw.rotationMatrix(), Scw.translation() / Scw.scale());
    Eigen::Vector3f Ow = Tcw.inverse().translation();
    std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint*>(nullptr));
    std::vector<MapPoint*> rows;
    std::vector<float> uvs;
    for (int iMP = 0, iendMP = (int)vpPoints.size(); iMP < iendMP; iMP++) {
        MapPoint* pMP = vpPoints[iMP];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;  // :491-492
        Eigen::Vector3f p3Dw = pMP->GetWorldPos();
        Eigen::Vector3f p3Dc = Tcw * p3Dw;
        if (p3Dc(2) < 0.0) continue;  // :499-500
        const Eigen::Vector2f uv = cam->project(p3Dc);
        if (!pKF->mpCamera->IsInImage(uv(0), uv(1))) continue;  // :506-507
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        Eigen::Vector3f PO = p3Dw - Ow;
        const float dist = PO.norm();
        if (dist < minDistance || dist > maxDistance) continue;  // :515-516
        Eigen::Vector3f Pn = pMP->GetNormal();
        if (PO.dot(Pn) < 0.5 * dist) continue;  // :522-523
        rows.push_back(pMP);
        uvs.push_back(uv(0));
        uvs.push_back(uv(1));
    }
    return detail::projection_match(ctx, pKF->N, pKF->mvKeysUn, pKF->mDescriptors, vpMatched, rows, uvs, true, (float)th,
                                    th_low * ratioHamming);
}

// Matcher::SearchForTriangulation (matching/src/Matcher.cpp:767-885) whole on the GPU (ppg_search_for_triangulation)
// for the reference's two camera models.  What depends on the two poses only -- the epipole, R12 / t12 and F12 -- is
// computed here with the reference's own classes and expressions (:776-788, sensors/src/Pinhole.cpp:101-104); the
// per-pair work (descriptor distances under the same vocabulary node, epipole exclusion, and the camera's
// epipolarConstrain: the epipolar distance of Pinhole, the two-view triangulation of KannalaBrandt8) runs on the device.
inline int search_for_triangulation(ppg_ctx* ctx, GeometricCamera* cam, KeyFrame* pKF1, KeyFrame* pKF2,
                                    std::vector<std::pair<size_t, size_t>>& vMatchedPairs, float th_low) {
    SE3f T1w = pKF1->GetPose();
    SE3f T2w = pKF2->GetPose();
    SE3f Tw2 = pKF2->GetPoseInverse();
    Eigen::Vector3f Cw = pKF1->GetCameraCenter();
    Eigen::Vector3f C2 = T2w * Cw;
    Eigen::Vector2f ep = cam->project(C2);
    SE3f T12 = T1w * Tw2;
    Eigen::Matrix3f R12 = T12.rotationMatrix();
    Eigen::Vector3f t12 = T12.translation();
    Eigen::Matrix3f t12x = SO3f::hat(t12);
    Eigen::Matrix3f K1 = cam->toK_();
    Eigen::Matrix3f K2 = cam->toK_();
    Eigen::Matrix3f F12 = K1.transpose().inverse() * t12x * R12 * K2.inverse();

    const int n1 = pKF1->N, n2 = pKF2->N;
    vMatchedPairs.clear();
    if (n1 <= 0 || n2 <= 0) return 0;
    std::vector<float> pos1(2 * (size_t)n1), pos2(2 * (size_t)n2);
    std::vector<int32_t> node1(n1, -1), node2(n2, -1), m12(n1, -1);
    std::vector<uint8_t> mp1(n1), mp2(n2);
    for (int i = 0; i < n1; i++) {
        pos1[2 * i] = pKF1->mvKeysUn[i].mPos[0];
        pos1[2 * i + 1] = pKF1->mvKeysUn[i].mPos[1];
        mp1[i] = pKF1->GetMapPoint(i) ? 1 : 0;
    }
    for (int i = 0; i < n2; i++) {
        pos2[2 * i] = pKF2->mvKeysUn[i].mPos[0];
        pos2[2 * i + 1] = pKF2->mvKeysUn[i].mPos[1];
        mp2[i] = pKF2->GetMapPoint(i) ? 1 : 0;
    }
    for (const auto& nf : pKF1->mFeatVec)
        for (unsigned int i : nf.second) node1[i] = (int32_t)nf.first;
    for (const auto& nf : pKF2->mFeatVec)
        for (unsigned int i : nf.second) node2[i] = (int32_t)nf.first;
    ppg_triangulation_match_in in{};
    in.n1 = n1;
    in.n2 = n2;
    in.desc1 = pKF1->mDescriptors.ptr<float>(0);
    in.desc2 = pKF2->mDescriptors.ptr<float>(0);
    in.node1 = node1.data();
    in.node2 = node2.data();
    in.has_mp1 = mp1.data();
    in.has_mp2 = mp2.data();
    in.pos1 = pos1.data();
    in.pos2 = pos2.data();
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) in.F12[3 * r + c] = F12(r, c);
    in.epipole[0] = ep[0];
    in.epipole[1] = ep[1];
    in.th_low = th_low;
    in.camera_model = cam->mnType == GeometricCamera::CAM_FISHEYE ? 1 : 0;
    for (int i = 0; i < 8; i++) in.cam8[i] = i < (int)cam->mvParameters.size() ? cam->mvParameters[i] : 0.f;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) in.R12[3 * r + c] = R12(r, c);
        in.t12[r] = t12[r];
    }
    ppg_triangulation_match_out out{};
    out.match12 = m12.data();
    check(ppg_search_for_triangulation(ctx, &in, &out), ctx, "ppg_search_for_triangulation");
    vMatchedPairs.reserve(out.nmatches);
    for (int i = 0; i < n1; i++)
        if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)m12[i]));  // :876-881
    return out.nmatches;
}

#ifndef PPG_SHIM_NO_MATCHER_CLASS
// Replaces class Matcher (matching/include/Matcher.h:20-64) for its callers: the same twelve signatures, constants and
// public members.  Eight matchers run on the GPU --
//   ExtendMapMatches          image <-> map association of MSTracking::SearchLocalPoints (system/src/Tracking.cpp:1007)
//   SearchByBoW(KF, F)        relocalisation / reference-keyframe tracking
//   SearchByBoW(KF, KF)       loop / merge candidates
//   SearchForInitialization   monocular initialisation (Tracking.cpp:525)
//   SearchByProjection(F, F)  tracking with the motion model, every frame (Tracking.cpp:811 / :817)
//   SearchByProjection(F, KF, sFound, ...)  relocalisation (Tracking.cpp:1297 / :1311)
//   SearchByProjection(KF, Scw, ...)        loop closing (LoopClosing.cpp:585 / :614 / :792)
//   SearchForTriangulation    new map points in LocalMapping (Pinhole: closed-form epipolar distance;
//                             KannalaBrandt8: the two-view triangulation of its epipolarConstrain, per pair on the device)
// -- and the other four (SearchByProjection(F, vpMapPoints, th), SearchBySim3, Fuse x 2) are the reference's own
// host code, inherited unchanged from ::Matcher (their window-search cores are available as search_window above for
// callers that want them on the device).
// The ctx is the one the frame's PPGExtractor owns (PPGExtractor::context()).
class Matcher : public ::Matcher {
public:
    Matcher(ppg_ctx* ctx, GeometricCamera* pCam, float nnratio = 0.6) : ::Matcher(pCam, nnratio), mCtx(ctx) {}

    int ExtendMapMatches(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th) {
        return extend_map_matches(mCtx, F, vpMapPoints, th, mfNNratio,
                                  [&]() { return ::Matcher::ExtendMapMatches(F, vpMapPoints, th); });
    }
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
        return search_by_bow(mCtx, pKF, F, vpMapPointMatches, mfNNratio, TH_LOW);
    }
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
        return search_by_bow(mCtx, pKF1, pKF2, vpMatches12, mfNNratio, TH_LOW);
    }
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched,
                                std::vector<int>& vnMatches12, int windowSize = 10) {
        return search_for_initialization(mCtx, F1, F2, vbPrevMatched, vnMatches12, windowSize, mfNNratio);
    }
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<std::pair<size_t, size_t>>& vMatchedPairs,
                               const bool bCoarse = false) {
        if (mpCamera->mnType != GeometricCamera::CAM_PINHOLE && mpCamera->mnType != GeometricCamera::CAM_FISHEYE)
            return ::Matcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bCoarse);  // a camera model of the caller's own
        return search_for_triangulation(mCtx, mpCamera, pKF1, pKF2, vMatchedPairs, TH_LOW);
    }
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th) {
        return search_by_projection(mCtx, CurrentFrame, LastFrame, th, TH_HIGH);
    }
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th,
                           const float descDist) {
        return search_by_projection(mCtx, CurrentFrame, pKF, sAlreadyFound, th, descDist);
    }
    int SearchByProjection(KeyFrame* pKF, Sim3f& Scw, const std::vector<MapPoint*>& vpPoints,
                           std::vector<MapPoint*>& vpMatched, int th, float ratioHamming = 1.0) {
        return search_by_projection(mCtx, mpCamera, pKF, Scw, vpPoints, vpMatched, th, ratioHamming, TH_LOW);
    }
    // host-side matchers of the reference, unchanged (SearchByProjection(F, vpMapPoints, th) included)
    using ::Matcher::Fuse;
    using ::Matcher::SearchByProjection;
    using ::Matcher::SearchBySim3;

private:
    ppg_ctx* mCtx;
};
#endif

}  // namespace ppg_shim
