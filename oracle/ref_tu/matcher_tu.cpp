// Reference-pinning harness, part 2: the reference's OWN matcher and map sources compiled as they lie under
// /root/reference (matching/src/Matcher.cpp, feature/src/MapPoint.cpp, map/src/Frame.cpp, feature/src/PPGGraph.cpp,
// sensors/src/GeometricCamera.cpp, sensors/src/Pinhole.cpp) against the stand-ins of oracle/ref_standins.
// TEST INFRASTRUCTURE ONLY.
//
// ref_extend_map_matches builds the pointer graph the reference works on -- MapPoint / MapEdge / Frame / KeyPointEx /
// KeyEdge OBJECTS -- from the flat arrays the C ABI and the oracle use (the inverse of the flattening in
// include/ppg_shim.hpp), calls the real Matcher::ExtendMapMatches (Matcher.cpp:203-381), which in turn calls the real
// Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (Frame.cpp:138-156, 262-327), MapPoint::isBad / getEdges /
// GetDescriptor / Observations, MapEdge::isBad / theOtherPt, KeyEdge::theOtherPid and DescriptorDistance
// (MapPoint.cpp:22-29), and reads F.mvpMapPoints / F.mvpMapEdges / mnTrackedbyFrame back.
// `private -> public` lets the harness fill the members a running system would have filled; no source is modified.
#include <algorithm>
#include <atomic>
#include <cassert>
#include <chrono>
#include <deque>
#include <iostream>
#include <list>
#include <map>
#include <mutex>
#include <numeric>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <unistd.h>

#include <torch/script.h>
#include <torch/torch.h>

#include "cv_eigen_standin.hpp"
#include "DBoW3/DBoW3.h"

#define private public
#define protected public
#include REF_FILE(sensors/src/KannalaBrandt8.cpp)  // first: no `using namespace std` of a later file in effect
#include REF_FILE(matching/src/Matcher.cpp)
#include REF_FILE(feature/src/MapPoint.cpp)
#include REF_FILE(map/src/Frame.cpp)
#include REF_FILE(feature/src/PPGGraph.cpp)
#include REF_FILE(sensors/src/GeometricCamera.cpp)
#include REF_FILE(sensors/src/Pinhole.cpp)
#undef protected
#undef private

namespace {

class HarnessCamera : public GeometricCamera {
   public:
    HarnessCamera(const std::vector<float>& p, int w, int h, bool fisheye) : GeometricCamera(p, w, h, 20.f) {
        mnId = 0;
        mnType = fisheye ? CAM_FISHEYE : CAM_PINHOLE;
        InitializeImageBounds();
    }
    Eigen::Vector2d project(const Eigen::Vector3d&) override { return Eigen::Vector2d(); }
    Eigen::Vector2f project(const Eigen::Vector3f&) override { return Eigen::Vector2f(); }
    Eigen::Vector3f unproject(const Eigen::Vector2f&) override { return Eigen::Vector3f(); }
    Eigen::Matrix<double, 2, 3> projectJac(const Eigen::Vector3d&) override { return Eigen::Matrix<double, 2, 3>(); }
    cv::Mat toK() override {  // sensors/src/Pinhole.cpp:68-72
        cv::Mat K = cv::Mat::zeros(3, 3, CV_32F);
        K.at<float>(0, 0) = mvParameters[0];
        K.at<float>(0, 2) = mvParameters[2];
        K.at<float>(1, 1) = mvParameters[1];
        K.at<float>(1, 2) = mvParameters[3];
        K.at<float>(2, 2) = 1.f;
        return K;
    }
    cv::Mat toD() override {  // :74-78
        cv::Mat D(4, 1, CV_32F);
        for (int i = 0; i < 4; i++) D.at<float>(i, 0) = mvParameters[4 + i];
        return D;
    }
    Eigen::Matrix3f toK_() override { return Eigen::Matrix3f(); }
    int imWidth() override { return mnWidth; }
    int imHeight() override { return mnHeight; }
    bool ReconstructWithTwoViews(const std::vector<KeyPointEx>&, const std::vector<KeyPointEx>&, const std::vector<int>&,
                                 SE3f&, std::vector<cv::Point3f>&, std::vector<bool>&) override {
        return false;
    }
    bool epipolarConstrain(const KeyPointEx&, const KeyPointEx&, const Eigen::Matrix3f&, const Eigen::Vector3f&) override {
        return false;
    }
};

MapPoint* make_point(KeyFrame* kf, const float* desc, bool bad, bool in_view, bool observed, float u, float v, float vcos,
                     bool tracked, unsigned long frame_id) {
    MapPoint* mp = new MapPoint(Eigen::Vector3f(0.f, 0.f, 1.f), kf);  // the real constructor
    mp->mbBad = bad;
    mp->mbTrackInView = in_view;
    mp->nObs = observed ? 1 : 0;
    mp->mTrackProjX = u;
    mp->mTrackProjY = v;
    mp->mTrackViewCos = vcos;
    mp->mnTrackedbyFrame = tracked ? frame_id : 0;
    mp->mDescriptor = cv::Mat(1, 256, CV_32F);
    if (desc) memcpy(mp->mDescriptor.data, desc, 1024);
    return mp;
}

}  // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

// Arguments as oracle/ppg_oracle.c::ppgo_extend_map_matches documents them (tracked / kp_mp / kedge_me in and out);
// params8 = fx fy cx cy d0..d3.  The caller must pass edge_ok consistent with MapEdge::isBad (an edge with a bad end
// point is bad) -- the flattening of include/ppg_shim.hpp guarantees that.  Returns nmatches, or -1.
REF_API int ref_extend_map_matches(const float* params8, int width, int height, int fisheye, int P, const float* map_desc,
                                   const unsigned char* candidate, const unsigned char* observed,
                                   const unsigned char* bad, const int* edge_off, const int* edge_other,
                                   const unsigned char* edge_ok, const float* proj_uv, const float* view_cos,
                                   unsigned char* tracked, int n, const float* kx, const float* ky, const float* frame_desc,
                                   int* kp_mp, int n_kedges, const int* kes, const int* kee, const int* conn_off,
                                   const int* conn_idx, int* kedge_me, float th, float ratio) {
    try {
        HarnessCamera cam(std::vector<float>(params8, params8 + 8), width, height, fisheye != 0);
        const unsigned long FID = 7;
        KeyFrame* kf = static_cast<KeyFrame*>(calloc(1, sizeof(KeyFrame)));  // the MapPoint constructor reads two ids of it
        std::vector<MapPoint*> pts(P);
        for (int p = 0; p < P; p++)
            pts[p] = make_point(kf, map_desc + (size_t)p * 256, bad[p] != 0, candidate[p] != 0, observed[p] != 0,
                                proj_uv[2 * p], proj_uv[2 * p + 1], view_cos[p], tracked[p] != 0, FID);
        // a point outside the table: the far end of edges whose theOtherPt(pMP) is nullptr, and the holder of
        // keypoints that are taken by a map point the caller did not list (kp_mp == -2)
        MapPoint* outsideA = make_point(kf, nullptr, false, false, true, 0, 0, 0, false, FID);
        MapPoint* outsideB = make_point(kf, nullptr, false, false, true, 0, 0, 0, false, FID);
        const int E = edge_off[P];
        std::vector<MapEdge*> edges(E);
        std::map<MapEdge*, int> edge_pos;
        for (int p = 0; p < P; p++)
            for (int k = edge_off[p]; k < edge_off[p + 1]; k++) {
                const int q = edge_other[k];
                MapEdge* e = q >= 0 ? new MapEdge(pts[p], pts[q]) : new MapEdge(outsideA, outsideB);  // real constructor
                e->mbValid = edge_ok[k] != 0;
                edges[k] = e;
                edge_pos[e] = k;
            }
        // getEdges() order = the CSR order (the constructor above appended every edge to both end points)
        for (int p = 0; p < P; p++) pts[p]->mvEdges.assign(edges.begin() + edge_off[p], edges.begin() + edge_off[p + 1]);

        Frame F;  // Frame::Frame()
        F.mnId = FID;
        F.N = n;
        F.mpCamera = &cam;
        F.mvKeysUn.resize(n);
        for (int i = 0; i < n; i++) {
            KeyPointEx k(kx[i], ky[i], 1.f);
            k.mPosUn = k.mPos;
            k.mbOut = false;
            for (int c = conn_off[i]; c < conn_off[i + 1]; c++) k.mvConnected.push_back((unsigned int)conn_idx[c]);
            F.mvKeysUn[i] = k;
        }
        F.mvKeys = F.mvKeysUn;
        for (int e = 0; e < n_kedges; e++) F.mvKeyEdges.emplace_back((unsigned int)kes[e], (unsigned int)kee[e]);
        F.mDescriptors = cv::Mat(std::max(n, 1), 256, CV_32F);
        if (n > 0) memcpy(F.mDescriptors.data, frame_desc, (size_t)n * 1024);
        F.mvpMapPoints.assign(n, nullptr);
        for (int i = 0; i < n; i++) F.mvpMapPoints[i] = kp_mp[i] >= 0 ? pts[kp_mp[i]] : (kp_mp[i] == -2 ? outsideA : nullptr);
        F.mvpMapEdges.assign(n_kedges, nullptr);
        for (int e = 0; e < n_kedges; e++) F.mvpMapEdges[e] = kedge_me[e] >= 0 ? edges[kedge_me[e]] : nullptr;
        F.AssignFeaturesToGrid();

        Matcher matcher(&cam, ratio);
        const int nm = matcher.ExtendMapMatches(F, pts, th);

        std::map<MapPoint*, int> row_of;
        for (int p = 0; p < P; p++) row_of[pts[p]] = p;
        for (int i = 0; i < n; i++) {
            MapPoint* m = F.mvpMapPoints[i];
            kp_mp[i] = !m ? -1 : (row_of.count(m) ? row_of[m] : -2);
        }
        for (int e = 0; e < n_kedges; e++) kedge_me[e] = F.mvpMapEdges[e] ? edge_pos[F.mvpMapEdges[e]] : -1;
        for (int p = 0; p < P; p++) tracked[p] = pts[p]->mnTrackedbyFrame == FID ? 1 : 0;
        for (MapEdge* e : edges) delete e;
        for (MapPoint* p : pts) delete p;
        delete outsideA;
        delete outsideB;
        free(kf);
        return nm;
    } catch (const std::exception& e) {
        std::cerr << "ref_extend_map_matches: " << e.what() << std::endl;
        return -1;
    }
}

// The order in which ExtendMapMatches walks the candidates: Matcher.cpp:206-224 filters `!isBad() && mbTrackInView` in
// vpMapPoints order and std::sorts by getEdges().size() descending -- an UNSTABLE sort, so the order of equal degrees is
// whatever this libstdc++'s introsort makes of that initial sequence.  The same call on the same sequence of degrees
// gives the same permutation; the tests use it to hand the oracle (whose documented tie rule is table order) a table
// that is already in the reference's walk order.  Returns the number of candidates.
REF_API int ref_candidate_order(int P, const unsigned char* candidate, const unsigned char* bad, const int* edge_off,
                                int* order) {
    std::vector<std::pair<int, size_t>> c;  // (row, getEdges().size())
    c.reserve(P);
    for (int p = 0; p < P; p++) {
        if (bad[p] || !candidate[p]) continue;
        c.emplace_back(p, (size_t)(edge_off[p + 1] - edge_off[p]));
    }
    std::sort(c.begin(), c.end(),
              [](const std::pair<int, size_t>& a, const std::pair<int, size_t>& b) { return a.second > b.second; });
    for (size_t i = 0; i < c.size(); i++) order[i] = c[i].first;
    return (int)c.size();
}

// The real Matcher::SearchForInitialization (Matcher.cpp:582-651) on two Frames rebuilt from flat arrays.
// prev (n1 x 2) is vbPrevMatched, in and out; matches12 (n1) out.  Returns nmatches.
REF_API int ref_search_for_initialization(const float* params8, int width, int height, int fisheye, int n1,
                                          const float* kx1, const float* ky1, const float* desc1, float* prev, int n2,
                                          const float* kx2, const float* ky2, const float* desc2, int window, float ratio,
                                          int* matches12) {
    HarnessCamera cam(std::vector<float>(params8, params8 + 8), width, height, fisheye != 0);
    auto fill = [&](Frame& F, int n, const float* kx, const float* ky, const float* d) {
        F.N = n;
        F.mpCamera = &cam;
        F.mvKeysUn.resize(n);
        for (int i = 0; i < n; i++) F.mvKeysUn[i] = KeyPointEx(kx[i], ky[i], 1.f);
        F.mvKeys = F.mvKeysUn;
        F.mDescriptors = cv::Mat(std::max(n, 1), 256, CV_32F);
        if (n > 0) memcpy(F.mDescriptors.data, d, (size_t)n * 1024);
        F.AssignFeaturesToGrid();
    };
    Frame F1, F2;
    fill(F1, n1, kx1, ky1, desc1);
    fill(F2, n2, kx2, ky2, desc2);
    std::vector<cv::Point2f> vprev(n1);
    for (int i = 0; i < n1; i++) vprev[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]);
    std::vector<int> m12;
    Matcher matcher(&cam, ratio);
    const int nm = matcher.SearchForInitialization(F1, F2, vprev, m12, window);
    for (int i = 0; i < n1; i++) {
        matches12[i] = m12[i];
        prev[2 * i] = vprev[i].x;
        prev[2 * i + 1] = vprev[i].y;
    }
    return nm;
}

// Frame::GetFeaturesInArea alone (the window query every matcher uses): indices in the reference's visiting order.
REF_API int ref_features_in_area(const float* params8, int width, int height, int fisheye, int n, const float* kx,
                                 const float* ky, float x, float y, float r, int* out) {
    HarnessCamera cam(std::vector<float>(params8, params8 + 8), width, height, fisheye != 0);
    Frame F;
    F.N = n;
    F.mpCamera = &cam;
    F.mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) {
        KeyPointEx k(kx[i], ky[i], 1.f);
        F.mvKeysUn[i] = k;
    }
    F.AssignFeaturesToGrid();
    const std::vector<size_t> v = F.GetFeaturesInArea(x, y, r);
    for (size_t i = 0; i < v.size(); i++) out[i] = (int)v[i];
    return (int)v.size();
}

// DescriptorDistance (feature/src/MapPoint.cpp:22-29) on two 1 x 256 rows.
REF_API float ref_descriptor_distance(const float* a, const float* b) {
    cv::Mat A(1, 256, CV_32F, const_cast<float*>(a)), B(1, 256, CV_32F, const_cast<float*>(b));
    return DescriptorDistance(A, B);
}

#include "keyframe_raw.hpp"

// params8 = fx fy cx cy + 4 distortion coefficients; R1 / t1, R2 / t2 = the world -> camera poses T1w, T2w (row-major R).
// fisheye = 0: the reference's Pinhole camera, 1: its KannalaBrandt8 camera (epipolarConstrain = TriangulateMatches,
// KannalaBrandt8.cpp:167-222, compiled unmodified; its JacobiSVD is the stand-in's, see eigen_standin.hpp).
// match12 (n1) out; F12 (9, row-major), ep (2), R12 (9, row-major), t12 (3) out: what Matcher.cpp:776-788 and
// Pinhole.cpp:101-104 compute from the poses, for the callers that hand the same numbers to the oracle and to the GPU.
// Returns nmatches.
REF_API int ref_search_for_triangulation(const float* params8, int width, int height, int fisheye, const float* R1,
                                         const float* t1, const float* R2, const float* t2, int n1, const float* pos1,
                                         const float* desc1, const int* node1, const unsigned char* mp1, int n2,
                                         const float* pos2, const float* desc2, const int* node2,
                                         const unsigned char* mp2, int* match12, float* F12out, float* epout,
                                         float* R12out, float* t12out) {
    const std::vector<float> prm(params8, params8 + 8);
    GeometricCamera* camp = fisheye ? static_cast<GeometricCamera*>(new KannalaBrandt8(prm, width, height, 20.f))
                                    : static_cast<GeometricCamera*>(new Pinhole(prm, width, height, 20.f));
    GeometricCamera& cam = *camp;
    MapPoint* some = static_cast<MapPoint*>(calloc(1, sizeof(MapPoint)));  // only tested against nullptr
    KeyFrame* k1 = raw_keyframe(n1, pos1, desc1, node1, mp1, some, pose_of(R1, t1));
    KeyFrame* k2 = raw_keyframe(n2, pos2, desc2, node2, mp2, some, pose_of(R2, t2));
    {   // the same expressions as Matcher.cpp:776-788 / Pinhole.cpp:101-104, evaluated by the reference's own classes
        SE3f T1w = k1->GetPose(), T2w = k2->GetPose(), Tw2 = k2->GetPoseInverse();
        Eigen::Vector3f Cw = k1->GetCameraCenter();
        Eigen::Vector3f C2 = T2w * Cw;
        Eigen::Vector2f ep = cam.project(C2);
        SE3f T12 = T1w * Tw2;
        Eigen::Matrix3f R12 = T12.rotationMatrix();
        Eigen::Vector3f t12 = T12.translation();
        Eigen::Matrix3f t12x = SO3f::hat(t12);
        Eigen::Matrix3f K1 = cam.toK_(), K2 = cam.toK_();
        Eigen::Matrix3f F = K1.transpose().inverse() * t12x * R12 * K2.inverse();
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                F12out[3 * i + j] = F(i, j);
                if (R12out) R12out[3 * i + j] = R12(i, j);
            }
        for (int i = 0; i < 3 && t12out; i++) t12out[i] = t12[i];
        epout[0] = ep[0];
        epout[1] = ep[1];
    }
    Matcher matcher(&cam, 0.8f);
    std::vector<std::pair<size_t, size_t>> pairs;
    const int nm = matcher.SearchForTriangulation(k1, k2, pairs, false);
    for (int i = 0; i < n1; i++) match12[i] = -1;
    for (const auto& pr : pairs) match12[pr.first] = (int)pr.second;
    drop_keyframe(k1);
    drop_keyframe(k2);
    free(some);
    delete camp;
    return nm;
}

// KannalaBrandt8::unproject / project / TriangulateMatches alone (sensors/src/KannalaBrandt8.cpp:44-91, :175-222) for the
// unit checks of the oracle's restatement.  Returns TriangulateMatches' value; r1 / r2 (3 each) = the unprojected rays,
// uv1 (2) = project(r1 * 3) (a point on the first ray), x3D (3) = the triangulated point when the value is > 0.
REF_API float ref_kb8_triangulate(const float* params8, int width, int height, const float* pos1, const float* pos2,
                                  const float* R12, const float* t12, float* r1, float* r2, float* uv1, float* x3D) {
    KannalaBrandt8 cam(std::vector<float>(params8, params8 + 8), width, height, 20.f);
    KeyPointEx k1(pos1[0], pos1[1], 1.f), k2(pos2[0], pos2[1], 1.f);
    Eigen::Vector3f a = cam.unproject(k1.mPos), b = cam.unproject(k2.mPos);
    Eigen::Vector2f p = cam.project(Eigen::Vector3f(a[0] * 3.f, a[1] * 3.f, 3.f));
    for (int i = 0; i < 3; i++) {
        r1[i] = a[i];
        r2[i] = b[i];
    }
    uv1[0] = p[0];
    uv1[1] = p[1];
    Eigen::Matrix3f R;
    Eigen::Vector3f t;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) R(i, j) = R12[3 * i + j];
        t[i] = t12[i];
    }
    Eigen::Vector3f X(0.f, 0.f, 0.f);
    const float z = cam.TriangulateMatches(k1, k2, R, t, X);
    for (int i = 0; i < 3; i++) x3D[i] = X[i];
    return z;
}

// Matcher::SearchByBoW(KeyFrame*, Frame&, ...) (Matcher.cpp:393-477) on a raw key frame (state_kf: 0 no map point, 1 good,
// 2 bad) and a Frame.  f2kf[i] = the key-frame feature whose map point frame feature i received, or -1.  -> nmatches.
REF_API int ref_search_by_bow_kf_f(const float* params8, int width, int height, int n_kf, const float* desc_kf,
                                   const int* node_kf, const unsigned char* state_kf, int n_f, const float* desc_f,
                                   const int* node_f, float ratio, int* f2kf) {
    HarnessCamera cam(std::vector<float>(params8, params8 + 8), width, height, false);
    std::vector<MapPoint*> owned;
    KeyFrame* kf = raw_keyframe_with_points(n_kf, desc_kf, node_kf, state_kf, owned);
    Frame F;
    F.N = n_f;
    F.mpCamera = &cam;
    F.mvKeysUn.resize(n_f);
    F.mDescriptors = cv::Mat(std::max(n_f, 1), 256, CV_32F);
    if (n_f > 0) memcpy(F.mDescriptors.data, desc_f, (size_t)n_f * 1024);
    for (int i = 0; i < n_f; i++)
        if (node_f[i] >= 0) F.mFeatVec[(unsigned int)node_f[i]].push_back((unsigned int)i);
    std::vector<MapPoint*> out;
    Matcher matcher(&cam, ratio);
    const int nm = matcher.SearchByBoW(kf, F, out);
    std::map<MapPoint*, int> feat_of;
    for (int i = 0; i < n_kf; i++)
        if (owned[i]) feat_of[owned[i]] = i;
    for (int i = 0; i < n_f; i++) f2kf[i] = out[i] ? feat_of[out[i]] : -1;
    for (MapPoint* m : owned) delete m;
    drop_keyframe(kf);
    return nm;
}

// Matcher::SearchByBoW(KeyFrame*, KeyFrame*, ...) (Matcher.cpp:663-754).  match12[i1] = the feature of KF2 whose map point
// vpMatches12[i1] is, or -1.  -> nmatches.
REF_API int ref_search_by_bow_kf_kf(const float* params8, int width, int height, int n1, const float* desc1,
                                    const int* node1, const unsigned char* state1, int n2, const float* desc2,
                                    const int* node2, const unsigned char* state2, float ratio, int* match12) {
    HarnessCamera cam(std::vector<float>(params8, params8 + 8), width, height, false);
    std::vector<MapPoint*> own1, own2;
    KeyFrame* k1 = raw_keyframe_with_points(n1, desc1, node1, state1, own1);
    KeyFrame* k2 = raw_keyframe_with_points(n2, desc2, node2, state2, own2);
    std::vector<MapPoint*> out;
    Matcher matcher(&cam, ratio);
    const int nm = matcher.SearchByBoW(k1, k2, out);
    std::map<MapPoint*, int> feat_of;
    for (int i = 0; i < n2; i++)
        if (own2[i]) feat_of[own2[i]] = i;
    for (int i = 0; i < n1; i++) match12[i] = out[i] ? feat_of[out[i]] : -1;
    for (MapPoint* m : own1) delete m;
    for (MapPoint* m : own2) delete m;
    drop_keyframe(k1);
    drop_keyframe(k2);
    return nm;
}

// The real Frame::CheckInFrustum (map/src/Frame.cpp:223-260) with the reference's own camera classes -- Pinhole::project
// (sensors/src/Pinhole.cpp:32-38) or KannalaBrandt8::project (sensors/src/KannalaBrandt8.cpp:44-59) -- and
// GeometricCamera::IsInImage on real MapPoint objects (GetWorldPos / GetNormal / GetMin / MaxDistanceInvariance).
// min_dist / max_dist are the invariance bounds themselves (0.5 mfMinDepth, 2 mfMaxDepth: exact in float).
// out4 (m x 4) = mTrackProjX, mTrackProjY, mTrackDepth, mTrackViewCos (0 where not in view: the reference leaves it
// unset); in_view = mbTrackInView; visible = mnVisible after the call (IncreaseVisible, :259).
REF_API void ref_check_in_frustum(const float* params8, int width, int height, int fisheye, const float* Rcw,
                                  const float* tcw, const float* Ow, int m, const float* world_pos, const float* normal,
                                  const float* min_dist, const float* max_dist, float cos_limit, unsigned char* in_view,
                                  float* out4, int* visible) {
    GeometricCamera* cam = fisheye ? static_cast<GeometricCamera*>(new KannalaBrandt8(
                                         std::vector<float>(params8, params8 + 8), width, height, 20.f))
                                   : static_cast<GeometricCamera*>(
                                         new Pinhole(std::vector<float>(params8, params8 + 8), width, height, 20.f));
    Frame F;
    F.mpCamera = cam;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) F.mRcw(i, j) = Rcw[3 * i + j];
        F.mtcw[i] = tcw[i];
        F.mOw[i] = Ow[i];
    }
    KeyFrame* kf = static_cast<KeyFrame*>(calloc(1, sizeof(KeyFrame)));
    for (int j = 0; j < m; j++) {
        MapPoint* mp = new MapPoint(Eigen::Vector3f(world_pos[3 * j], world_pos[3 * j + 1], world_pos[3 * j + 2]), kf);
        mp->mNormalVector = Eigen::Vector3f(normal[3 * j], normal[3 * j + 1], normal[3 * j + 2]);
        mp->mfMinDepth = min_dist[j] * 2.0f;
        mp->mfMaxDepth = max_dist[j] * 0.5f;
        mp->mTrackViewCos = 0.f;
        const int before = mp->mnVisible;
        F.CheckInFrustum(mp, cos_limit);
        in_view[j] = mp->mbTrackInView ? 1 : 0;
        out4[4 * j] = mp->mTrackProjX;
        out4[4 * j + 1] = mp->mTrackProjY;
        out4[4 * j + 2] = mp->mTrackDepth;
        out4[4 * j + 3] = mp->mTrackViewCos;
        visible[j] = mp->mnVisible - before;
        delete mp;
    }
    free(kf);
    delete cam;
}

// The real Matcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th) (Matcher.cpp:31-87; mode 0) and
// Matcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, sAlreadyFound, th, descDist) (:1337-1411; mode 1) on
// objects rebuilt from flat arrays, with the reference's own Pinhole / KannalaBrandt8 camera; mode 2:
// SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (:479-568) (projection_case in keyframe_raw.hpp
// documents the arrays).  Returns nmatches.
REF_API int ref_search_by_projection(const float* params8, int width, int height, int fisheye, int mode, const float* Rcw,
                                     const float* tcw, int n_src, const float* world_pos, const float* mp_desc,
                                     const unsigned char* state, const unsigned char* observed, const float* min_dist,
                                     const float* max_dist, int n, const float* kx, const float* ky, const float* desc,
                                     int* kp_mp, float th, float desc_dist, float* proj_uv, unsigned char* row_valid,
                                     const float* normal, float scale) {
    const std::vector<float> prm(params8, params8 + 8);
    GeometricCamera* cam = fisheye ? static_cast<GeometricCamera*>(new KannalaBrandt8(prm, width, height, 20.f))
                                   : static_cast<GeometricCamera*>(new Pinhole(prm, width, height, 20.f));
    Matcher matcher(cam, 0.9f);
    const int nm = projection_case(matcher, cam, mode, Rcw, tcw, n_src, world_pos, mp_desc, state, observed, min_dist,
                                   max_dist, n, kx, ky, desc, kp_mp, th, desc_dist, proj_uv, row_valid, normal, scale);
    delete cam;
    return nm;
}

// The real MapPoint::ComputeDistinctiveDescriptors (feature/src/MapPoint.cpp:234-302) on a MapPoint whose observations
// are raw key frames, one per row of obs_desc (state: 0 good, 1 the key frame isBad(), 2 index -1).  mObservations is a
// std::map keyed by KeyFrame*: the key frames are carved out of ONE allocation in row order, so the map iterates them in
// that order.  out_desc (256) = mDescriptor after the call (the initial one, all -1, when the function returns early).
REF_API void ref_distinctive_descriptor(int n_obs, const float* obs_desc, const unsigned char* state, int point_bad,
                                        float* out_desc) {
    KeyFrame* kfs = static_cast<KeyFrame*>(calloc((size_t)std::max(n_obs, 1), sizeof(KeyFrame)));
    std::vector<float> pos(2, 0.f);
    for (int i = 0; i < n_obs; i++) {
        KeyFrame* kf = kfs + i;
        new (&kf->mDescriptors) cv::Mat(1, 256, CV_32F);
        memcpy(kf->mDescriptors.data, obs_desc + (size_t)i * 256, 1024);
        new (&kf->mMutexConnections) std::mutex();
        kf->mbBad = state[i] == 1;
    }
    KeyFrame* kf0 = static_cast<KeyFrame*>(calloc(1, sizeof(KeyFrame)));
    MapPoint* mp = new MapPoint(Eigen::Vector3f(0.f, 0.f, 1.f), kf0);
    mp->mbBad = point_bad != 0;
    mp->mDescriptor = cv::Mat(1, 256, CV_32F);
    for (int k = 0; k < 256; k++) mp->mDescriptor.at<float>(0, k) = -1.f;
    for (int i = 0; i < n_obs; i++) mp->mObservations[kfs + i] = state[i] == 2 ? -1 : 0;
    mp->ComputeDistinctiveDescriptors();
    memcpy(out_desc, mp->mDescriptor.data, 1024);
    delete mp;
    for (int i = 0; i < n_obs; i++) (kfs + i)->mDescriptors.~Mat();
    free(kfs);
    free(kf0);
}

// The real Matcher::SearchBySim3 (Matcher.cpp:1149-1335; loop closing) on two raw key frames whose features carry real
// MapPoint objects.  Per key frame k: n_k features (positions, descriptors), state_k (0 no map point, 1 a point, 2 a bad
// point), world_pos_k / mp_desc_k / min_dist_k / max_dist_k of the points, the pose R_k / t_k (world -> camera).
// S12 = (R12, t12, s12).  matches12 (n1) in / out: the KF2 FEATURE whose map point vpMatches12[i1] is, or -1.
// Also out, per feature of either key frame: whether the loop head lets it search (valid_k) and where it projects into
// the other key frame (uv_k), by the same expressions as :1189-1213 / :1253-1277.  The key-frame grids answer through the
// reference's real Frame::GetFeaturesInArea of shadow frames (keyframe_raw.hpp).  Returns nFound.
REF_API int ref_search_by_sim3(const float* params8, int width, int height, int fisheye, const float* R1, const float* t1,
                               const float* R2, const float* t2, const float* R12, const float* t12, float s12, int n1,
                               const float* pos1, const float* desc1, const unsigned char* state1, const float* world1,
                               const float* mpdesc1, const float* mind1, const float* maxd1, int n2, const float* pos2,
                               const float* desc2, const unsigned char* state2, const float* world2,
                               const float* mpdesc2, const float* mind2, const float* maxd2, int* matches12, float th,
                               unsigned char* valid1, float* uv1, unsigned char* valid2, float* uv2) {
    const std::vector<float> prm(params8, params8 + 8);
    GeometricCamera* cam = fisheye ? static_cast<GeometricCamera*>(new KannalaBrandt8(prm, width, height, 20.f))
                                   : static_cast<GeometricCamera*>(new Pinhole(prm, width, height, 20.f));
    KeyFrame* kf0 = static_cast<KeyFrame*>(calloc(1, sizeof(KeyFrame)));
    struct Side {
        KeyFrame* kf;
        Frame shadow;
        std::vector<MapPoint*> pts;
    } S[2];
    const int n[2] = {n1, n2};
    const float* pos[2] = {pos1, pos2};
    const float* desc[2] = {desc1, desc2};
    const unsigned char* state[2] = {state1, state2};
    const float* world[2] = {world1, world2};
    const float* mpdesc[2] = {mpdesc1, mpdesc2};
    const float* mind[2] = {mind1, mind2};
    const float* maxd[2] = {maxd1, maxd2};
    const float* Rk[2] = {R1, R2};
    const float* tk[2] = {t1, t2};
    for (int k = 0; k < 2; k++) {
        std::vector<int> node((size_t)std::max(n[k], 1), -1);
        std::vector<unsigned char> none((size_t)std::max(n[k], 1), 0);
        S[k].kf = raw_keyframe(n[k], pos[k], desc[k], node.data(), none.data(), nullptr, pose_of(Rk[k], tk[k]));
        S[k].kf->mpCamera = cam;
        Frame& F = S[k].shadow;
        F.N = n[k];
        F.mpCamera = cam;
        F.mvKeysUn = S[k].kf->mvKeysUn;
        F.mvKeys = F.mvKeysUn;
        F.AssignFeaturesToGrid();
        kf_shadow[S[k].kf] = &F;
        S[k].pts.assign(n[k], nullptr);
        for (int i = 0; i < n[k]; i++) {
            if (!state[k][i]) continue;
            MapPoint* mp = new MapPoint(Eigen::Vector3f(world[k][3 * i], world[k][3 * i + 1], world[k][3 * i + 2]), kf0);
            mp->mbBad = state[k][i] == 2;
            mp->mfMinDepth = mind[k][i] * 2.0f;
            mp->mfMaxDepth = maxd[k][i] * 0.5f;
            mp->mDescriptor = cv::Mat(1, 256, CV_32F);
            memcpy(mp->mDescriptor.data, mpdesc[k] + (size_t)i * 256, 1024);
            mp->mObservations[S[k].kf] = i;  // GetIndexInKeyFrame (:1171)
            S[k].pts[i] = mp;
        }
        S[k].kf->mvpMapPoints = S[k].pts;
    }
    Eigen::Matrix3f Rm;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Rm(i, j) = R12[3 * i + j];
    const Sim3f S12(Rm, Eigen::Vector3f(t12[0], t12[1], t12[2]), s12);
    std::vector<MapPoint*> vpMatches12(n1, nullptr);
    for (int i = 0; i < n1; i++)
        if (matches12[i] >= 0) vpMatches12[i] = S[1].pts[matches12[i]];
    {   // the loop heads, for the callers that feed the oracle's frozen search core
        SE3f T1w = S[0].kf->GetPose(), T2w = S[1].kf->GetPose();
        Sim3f S21 = S12.inverse();
        std::vector<bool> am1(n1, false), am2(n2, false);
        for (int i = 0; i < n1; i++)
            if (vpMatches12[i]) {
                am1[i] = true;
                int idx2 = vpMatches12[i]->GetIndexInKeyFrame(S[1].kf);
                if (idx2 >= 0 && idx2 < n2) am2[idx2] = true;
            }
        for (int k = 0; k < 2; k++) {
            unsigned char* valid = k == 0 ? valid1 : valid2;
            float* uvo = k == 0 ? uv1 : uv2;
            for (int i = 0; i < n[k]; i++) {
                valid[i] = 0;
                uvo[2 * i] = uvo[2 * i + 1] = -1.f;
                MapPoint* pMP = S[k].pts[i];
                if (!pMP || (k == 0 ? am1[i] : am2[i])) continue;
                if (pMP->isBad()) continue;
                Eigen::Vector3f p3Dw = pMP->GetWorldPos();
                Eigen::Vector3f own = (k == 0 ? T1w : T2w) * p3Dw;
                Eigen::Vector3f other = k == 0 ? S21 * own : S12 * own;
                if (other(2) < 0.0) continue;
                const Eigen::Vector2f uv = cam->project(other);
                if (!cam->IsInImage(uv[0], uv[1])) continue;
                const float dist3D = other.norm();
                if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
                valid[i] = 1;
                uvo[2 * i] = uv[0];
                uvo[2 * i + 1] = uv[1];
            }
        }
    }
    Matcher matcher(cam, 0.75f);
    const int nFound = matcher.SearchBySim3(S[0].kf, S[1].kf, vpMatches12, S12, th);
    std::map<MapPoint*, int> idx2_of;
    for (int i = 0; i < n2; i++)
        if (S[1].pts[i]) idx2_of[S[1].pts[i]] = i;
    for (int i = 0; i < n1; i++) matches12[i] = vpMatches12[i] ? idx2_of[vpMatches12[i]] : -1;
    for (int k = 0; k < 2; k++) {
        kf_shadow.erase(S[k].kf);
        for (MapPoint* m : S[k].pts) delete m;
        drop_keyframe(S[k].kf);
    }
    free(kf0);
    delete cam;
    return nFound;
}
