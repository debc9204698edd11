// Reference-pinning harness, part 3: include/ppg_shim.hpp compiled against the reference's REAL headers (Frame.h,
// MapPoint.h, PPGGraph.h, GeometricCamera.h, Matcher.h from /root/reference) and EXECUTED on real MapPoint / MapEdge /
// Frame objects: the same pointer graph goes once through the reference's own Matcher::ExtendMapMatches (host) and once
// through ppg_shim::Matcher::ExtendMapMatches (flattening -> C ABI -> GPU -> write-back), and what the two leave in
// F.mvpMapPoints / F.mvpMapEdges / mnTrackedbyFrame is handed back for comparison.  Same for SearchForInitialization and
// SearchForTriangulation (two raw KeyFrames, the reference's own Pinhole / KannalaBrandt8 camera).
// Links libppg_b200.so; needs a GPU at run time (tests/test_ref_pin.py, -m gpu).  TEST INFRASTRUCTURE ONLY.
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
#include REF_FILE(sensors/src/KannalaBrandt8.cpp)  // first: before the `using namespace std` of Matcher.cpp
#include REF_FILE(matching/src/Matcher.cpp)
#include REF_FILE(feature/src/MapPoint.cpp)
#include REF_FILE(map/src/Frame.cpp)
#include REF_FILE(feature/src/PPGGraph.cpp)
#include REF_FILE(sensors/src/GeometricCamera.cpp)
#include REF_FILE(sensors/src/Pinhole.cpp)
#undef protected
#undef private

#define PPG_SHIM_NO_REFERENCE_HEADERS  // they are all in already, through Matcher.cpp
#include "ppg_shim.hpp"
#include "keyframe_raw.hpp"

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
    cv::Mat toK() override {
        cv::Mat K = cv::Mat::zeros(3, 3, CV_32F);
        K.at<float>(0, 0) = mvParameters[0];
        K.at<float>(0, 2) = mvParameters[2];
        K.at<float>(1, 1) = mvParameters[1];
        K.at<float>(1, 2) = mvParameters[3];
        K.at<float>(2, 2) = 1.f;
        return K;
    }
    cv::Mat toD() override {
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

// One world: map points, map edges and a frame as objects, built from the flat arrays.
struct World {
    HarnessCamera* cam;
    KeyFrame* kf;
    std::vector<MapPoint*> pts;
    MapPoint *outsideA, *outsideB;
    std::vector<MapEdge*> edges;
    std::map<MapEdge*, int> edge_pos;
    std::map<MapPoint*, int> row_of;
    Frame F;
    static const unsigned long FID = 7;

    World(HarnessCamera* c, int P, const float* map_desc, const unsigned char* candidate, const unsigned char* observed,
          const unsigned char* bad, const int* edge_off, const int* edge_other, const unsigned char* edge_ok,
          const float* proj_uv, const float* view_cos, const unsigned char* tracked, int n, const float* kx,
          const float* ky, const float* frame_desc, const int* kp_mp, int n_kedges, const int* kes, const int* kee,
          const int* conn_off, const int* conn_idx)
        : cam(c) {
        kf = static_cast<KeyFrame*>(calloc(1, sizeof(KeyFrame)));
        auto mk = [&](const float* d, bool b, bool v, bool o, float u, float vv, float vc, bool t) {
            MapPoint* mp = new MapPoint(Eigen::Vector3f(0.f, 0.f, 1.f), kf);
            mp->mbBad = b;
            mp->mbTrackInView = v;
            mp->nObs = o ? 1 : 0;
            mp->mTrackProjX = u;
            mp->mTrackProjY = vv;
            mp->mTrackViewCos = vc;
            mp->mnTrackedbyFrame = t ? FID : 0;
            mp->mDescriptor = cv::Mat(1, 256, CV_32F);
            if (d) memcpy(mp->mDescriptor.data, d, 1024);
            return mp;
        };
        pts.resize(P);
        for (int p = 0; p < P; p++) {
            pts[p] = mk(map_desc + (size_t)p * 256, bad[p] != 0, candidate[p] != 0, observed[p] != 0, proj_uv[2 * p],
                        proj_uv[2 * p + 1], view_cos[p], tracked[p] != 0);
            row_of[pts[p]] = p;
        }
        outsideA = mk(nullptr, false, false, true, 0, 0, 0, false);
        outsideB = mk(nullptr, false, false, true, 0, 0, 0, false);
        edges.resize(edge_off[P]);
        for (int p = 0; p < P; p++)
            for (int k = edge_off[p]; k < edge_off[p + 1]; k++) {
                const int q = edge_other[k];
                MapEdge* e = q >= 0 ? new MapEdge(pts[p], pts[q]) : new MapEdge(outsideA, outsideB);
                e->mbValid = edge_ok[k] != 0;
                edges[k] = e;
                edge_pos[e] = k;
            }
        for (int p = 0; p < P; p++) pts[p]->mvEdges.assign(edges.begin() + edge_off[p], edges.begin() + edge_off[p + 1]);
        F.mnId = FID;
        F.N = n;
        F.mpCamera = cam;
        F.mvKeysUn.resize(n);
        for (int i = 0; i < n; i++) {
            KeyPointEx k(kx[i], ky[i], 1.f);
            k.mPosUn = k.mPos;
            k.mbOut = false;
            for (int cc = conn_off[i]; cc < conn_off[i + 1]; cc++) k.mvConnected.push_back((unsigned int)conn_idx[cc]);
            F.mvKeysUn[i] = k;
        }
        F.mvKeys = F.mvKeysUn;
        for (int e = 0; e < n_kedges; e++) F.mvKeyEdges.emplace_back((unsigned int)kes[e], (unsigned int)kee[e]);
        F.mDescriptors = cv::Mat(std::max(n, 1), 256, CV_32F);
        if (n > 0) memcpy(F.mDescriptors.data, frame_desc, (size_t)n * 1024);
        F.mvpMapPoints.assign(n, nullptr);
        for (int i = 0; i < n; i++) F.mvpMapPoints[i] = kp_mp[i] >= 0 ? pts[kp_mp[i]] : (kp_mp[i] == -2 ? outsideA : nullptr);
        F.mvpMapEdges.assign(n_kedges, nullptr);
        F.AssignFeaturesToGrid();
    }
    void read(int* kp_mp, int* kedge_me, unsigned char* tracked) {
        for (size_t i = 0; i < F.mvpMapPoints.size(); i++) {
            MapPoint* m = F.mvpMapPoints[i];
            kp_mp[i] = !m ? -1 : (row_of.count(m) ? row_of[m] : -2);
        }
        for (size_t e = 0; e < F.mvpMapEdges.size(); e++) kedge_me[e] = F.mvpMapEdges[e] ? edge_pos[F.mvpMapEdges[e]] : -1;
        for (size_t p = 0; p < pts.size(); p++) tracked[p] = pts[p]->mnTrackedbyFrame == FID ? 1 : 0;
    }
    ~World() {
        for (MapEdge* e : edges) delete e;
        for (MapPoint* p : pts) delete p;
        delete outsideA;
        delete outsideB;
        free(kf);
    }
};

}  // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

// Runs BOTH on identical worlds.  vp_order (n_vp rows): the vpMapPoints argument handed to the shim -- the candidates in
// the reference's walk order (the shim's documented tie rule is the order of vpMapPoints), then the rest; the reference
// gets the rows in table order.  Outputs: [0] = reference, [1] = shim.  Returns 0, or -1 with the message on stderr.
REF_API int shim_extend_both(const float* params8, int width, int height, int fisheye, const char* weights, int P,
                             const float* map_desc, const unsigned char* candidate, const unsigned char* observed,
                             const unsigned char* bad, const int* edge_off, const int* edge_other,
                             const unsigned char* edge_ok, const float* proj_uv, const float* view_cos,
                             const unsigned char* tracked, int n, const float* kx, const float* ky,
                             const float* frame_desc, const int* kp_mp, int n_kedges, const int* kes, const int* kee,
                             const int* conn_off, const int* conn_idx, const int* vp_order, int n_vp, float th, float ratio,
                             int* nmatches2, int* kp_mp2, int* kedge_me2, unsigned char* tracked2) {
    try {
        HarnessCamera cam(std::vector<float>(params8, params8 + 8), width, height, fisheye != 0);
        {
            World w(&cam, P, map_desc, candidate, observed, bad, edge_off, edge_other, edge_ok, proj_uv, view_cos, tracked, n,
                    kx, ky, frame_desc, kp_mp, n_kedges, kes, kee, conn_off, conn_idx);
            ::Matcher ref(&cam, ratio);
            nmatches2[0] = ref.ExtendMapMatches(w.F, w.pts, th);
            w.read(kp_mp2, kedge_me2, tracked2);
        }
        {
            World w(&cam, P, map_desc, candidate, observed, bad, edge_off, edge_other, edge_ok, proj_uv, view_cos, tracked, n,
                    kx, ky, frame_desc, kp_mp, n_kedges, kes, kee, conn_off, conn_idx);
            // the ctx as the drop-in extractor class creates it (ppg_shim::PPGExtractor: ppg_create from the camera)
            ppg_shim::PPGExtractor::JUNCTION_MAX_NUM = 500;
            struct WeightsCam : HarnessCamera {
                using HarnessCamera::HarnessCamera;
            };
            std::string dir(weights);
            ppg_shim::PPGExtractor ex(&cam, dir);  // loads <dir>/ppg_weights.bin
            ppg_shim::Matcher m(ex.context(), &cam, ratio);
            std::vector<MapPoint*> vp(n_vp);
            for (int i = 0; i < n_vp; i++) vp[i] = w.pts[vp_order[i]];
            nmatches2[1] = m.ExtendMapMatches(w.F, vp, th);
            w.read(kp_mp2 + std::max(n, 1), kedge_me2 + std::max(n_kedges, 1), tracked2 + std::max(P, 1));
        }
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "shim_extend_both: " << e.what() << std::endl;
        return -1;
    }
}

REF_API int shim_init_both(const float* params8, int width, int height, int fisheye, const char* weights, int n1,
                           const float* kx1, const float* ky1, const float* desc1, const float* prev, int n2,
                           const float* kx2, const float* ky2, const float* desc2, int window, float ratio, int* nmatches2,
                           int* matches12_2, float* prev2) {
    try {
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
        ppg_shim::PPGExtractor ex(&cam, std::string(weights));
        for (int which = 0; which < 2; which++) {
            Frame F1, F2;
            fill(F1, n1, kx1, ky1, desc1);
            fill(F2, n2, kx2, ky2, desc2);
            std::vector<cv::Point2f> vprev(n1);
            for (int i = 0; i < n1; i++) vprev[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]);
            std::vector<int> m12;
            if (which == 0) {
                ::Matcher ref(&cam, ratio);
                nmatches2[0] = ref.SearchForInitialization(F1, F2, vprev, m12, window);
            } else {
                ppg_shim::Matcher m(ex.context(), &cam, ratio);
                nmatches2[1] = m.SearchForInitialization(F1, F2, vprev, m12, window);
            }
            for (int i = 0; i < n1; i++) {
                matches12_2[which * std::max(n1, 1) + i] = m12[i];
                prev2[(which * std::max(n1, 1) + i) * 2] = vprev[i].x;
                prev2[(which * std::max(n1, 1) + i) * 2 + 1] = vprev[i].y;
            }
        }
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "shim_init_both: " << e.what() << std::endl;
        return -1;
    }
}

// Matcher::SearchForTriangulation: the reference's host function and ppg_shim::Matcher's GPU path on the same two key
// frames with the reference's own Pinhole (fisheye = 0) or KannalaBrandt8 (1) camera.  match12_2 = [2][max(n1, 1)]:
// 0 reference, 1 shim.
REF_API int shim_triangulation_both(const float* params8, int width, int height, int fisheye, const char* weights, const float* R1,
                                    const float* t1, const float* R2, const float* t2, int n1, const float* pos1,
                                    const float* desc1, const int* node1, const unsigned char* mp1, int n2,
                                    const float* pos2, const float* desc2, const int* node2, const unsigned char* mp2,
                                    int* nmatches2, int* match12_2) {
    try {
        const std::vector<float> prm(params8, params8 + 8);
        std::unique_ptr<GeometricCamera> camp(fisheye ? static_cast<GeometricCamera*>(new KannalaBrandt8(prm, width, height, 20.f))
                                                      : static_cast<GeometricCamera*>(new Pinhole(prm, width, height, 20.f)));
        GeometricCamera& cam = *camp;
        ppg_shim::PPGExtractor ex(&cam, std::string(weights));
        MapPoint* some = static_cast<MapPoint*>(calloc(1, sizeof(MapPoint)));
        KeyFrame* k1 = raw_keyframe(n1, pos1, desc1, node1, mp1, some, pose_of(R1, t1));
        KeyFrame* k2 = raw_keyframe(n2, pos2, desc2, node2, mp2, some, pose_of(R2, t2));
        for (int which = 0; which < 2; which++) {
            std::vector<std::pair<size_t, size_t>> pairs;
            if (which == 0) {
                ::Matcher ref(&cam, 0.8f);
                nmatches2[0] = ref.SearchForTriangulation(k1, k2, pairs, false);
            } else {
                ppg_shim::Matcher m(ex.context(), &cam, 0.8f);
                nmatches2[1] = m.SearchForTriangulation(k1, k2, pairs, false);
            }
            int* out = match12_2 + which * std::max(n1, 1);
            for (int i = 0; i < n1; i++) out[i] = -1;
            size_t last = 0;
            for (size_t k = 0; k < pairs.size(); k++) {
                if (k > 0 && pairs[k].first <= last) return -2;  // vMatchedPairs is in ascending order of the first index
                last = pairs[k].first;
                out[pairs[k].first] = (int)pairs[k].second;
            }
        }
        drop_keyframe(k1);
        drop_keyframe(k2);
        free(some);
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "shim_triangulation_both: " << e.what() << std::endl;
        return -1;
    }
}

// Matcher::SearchByBoW, both overloads: the reference's host functions and ppg_shim::Matcher's GPU paths on the same raw
// key frames (state: 0 no map point, 1 good, 2 bad) / Frame.  out2 = [2][max(n, 1)]: 0 reference, 1 shim;
// kf_kf = 0: f2kf of the frame features (n = n2), kf_kf = 1: match12 of KF1's features (n = n1).
REF_API int shim_bow_both(const float* params8, int width, int height, const char* weights, int kf_kf, int n1,
                          const float* desc1, const int* node1, const unsigned char* state1, int n2, const float* desc2,
                          const int* node2, const unsigned char* state2, float ratio, int* nmatches2, int* out2) {
    try {
        Pinhole cam(std::vector<float>(params8, params8 + 8), width, height, 20.f);
        ppg_shim::PPGExtractor ex(&cam, std::string(weights));
        std::vector<MapPoint*> own1, own2;
        KeyFrame* k1 = raw_keyframe_with_points(n1, desc1, node1, state1, own1);
        std::map<MapPoint*, int> feat1, feat2;
        for (int i = 0; i < n1; i++)
            if (own1[i]) feat1[own1[i]] = i;
        const int n_out = std::max(kf_kf ? n1 : n2, 1);
        if (kf_kf) {
            KeyFrame* k2 = raw_keyframe_with_points(n2, desc2, node2, state2, own2);
            for (int i = 0; i < n2; i++)
                if (own2[i]) feat2[own2[i]] = i;
            for (int which = 0; which < 2; which++) {
                std::vector<MapPoint*> out;
                if (which == 0) {
                    ::Matcher ref(&cam, ratio);
                    nmatches2[0] = ref.SearchByBoW(k1, k2, out);
                } else {
                    ppg_shim::Matcher m(ex.context(), &cam, ratio);
                    nmatches2[1] = m.SearchByBoW(k1, k2, out);
                }
                for (int i = 0; i < n1; i++) out2[which * n_out + i] = out[i] ? feat2[out[i]] : -1;
            }
            for (MapPoint* m : own2) delete m;
            drop_keyframe(k2);
        } else {
            for (int which = 0; which < 2; which++) {
                Frame F;
                F.N = n2;
                F.mpCamera = &cam;
                F.mvKeysUn.resize(n2);
                F.mDescriptors = cv::Mat(std::max(n2, 1), 256, CV_32F);
                if (n2 > 0) memcpy(F.mDescriptors.data, desc2, (size_t)n2 * 1024);
                for (int i = 0; i < n2; i++)
                    if (node2[i] >= 0) F.mFeatVec[(unsigned int)node2[i]].push_back((unsigned int)i);
                std::vector<MapPoint*> out;
                if (which == 0) {
                    ::Matcher ref(&cam, ratio);
                    nmatches2[0] = ref.SearchByBoW(k1, F, out);
                } else {
                    ppg_shim::Matcher m(ex.context(), &cam, ratio);
                    nmatches2[1] = m.SearchByBoW(k1, F, out);
                }
                for (int i = 0; i < n2; i++) out2[which * n_out + i] = out[i] ? feat1[out[i]] : -1;
            }
        }
        for (MapPoint* m : own1) delete m;
        drop_keyframe(k1);
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "shim_bow_both: " << e.what() << std::endl;
        return -1;
    }
}

// Matcher::SearchByProjection (mode 0: CurrentFrame / LastFrame, mode 1: CurrentFrame / KeyFrame / sAlreadyFound): the
// reference's host function and ppg_shim::Matcher's GPU path on identical objects.  kp_mp_2 = [2][max(n, 1)]: 0
// reference, 1 shim (coding of projection_case in keyframe_raw.hpp).
REF_API int shim_projection_both(const float* params8, int width, int height, int fisheye, const char* weights, int mode,
                                 const float* Rcw, const float* tcw, int n_src, const float* world_pos,
                                 const float* mp_desc, const unsigned char* state, const unsigned char* observed,
                                 const float* min_dist, const float* max_dist, int n, const float* kx, const float* ky,
                                 const float* desc, const int* kp_mp, float th, float desc_dist, int* nmatches2,
                                 int* kp_mp_2, const float* normal, float scale) {
    try {
        const std::vector<float> prm(params8, params8 + 8);
        std::unique_ptr<GeometricCamera> camp(fisheye ? static_cast<GeometricCamera*>(new KannalaBrandt8(prm, width, height, 20.f))
                                                      : static_cast<GeometricCamera*>(new Pinhole(prm, width, height, 20.f)));
        for (int which = 0; which < 2; which++) {
            int* out = kp_mp_2 + which * std::max(n, 1);
            for (int i = 0; i < n; i++) out[i] = kp_mp[i];
            if (which == 0) {
                ::Matcher ref(camp.get(), 0.9f);
                nmatches2[0] = projection_case(ref, camp.get(), mode, Rcw, tcw, n_src, world_pos, mp_desc, state, observed,
                                               min_dist, max_dist, n, kx, ky, desc, out, th, desc_dist, nullptr, nullptr, normal, scale);
            } else {
                ppg_shim::PPGExtractor ex(camp.get(), std::string(weights));  // needs the GPU: after the host leg
                ppg_shim::Matcher m(ex.context(), camp.get(), 0.9f);
                nmatches2[1] = projection_case(m, camp.get(), mode, Rcw, tcw, n_src, world_pos, mp_desc, state, observed,
                                               min_dist, max_dist, n, kx, ky, desc, out, th, desc_dist, nullptr, nullptr, normal, scale);
            }
        }
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "shim_projection_both: " << e.what() << std::endl;
        return -1;
    }
}
