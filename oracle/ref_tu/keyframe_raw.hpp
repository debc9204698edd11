// Shared by matcher_tu.cpp and shim_tu.cpp (included AFTER the reference's sources).  TEST INFRASTRUCTURE ONLY.
#pragma once
// ------------------------------------------------------------------------------------------------
// Matcher::SearchForTriangulation (Matcher.cpp:767-885) with the reference's own Pinhole camera (project,
// epipolarConstrain: sensors/src/Pinhole.cpp:33-39, 98-114).  feature/src/KeyFrame.cpp is not compiled (it pulls in the
// whole map): the four accessors the function calls are restated below as they stand at KeyFrame.cpp:56-72, :291-295,
// and the two KeyFrame objects are raw storage with exactly the members the function reads constructed in place.
SE3f KeyFrame::GetPose() {
    std::unique_lock<std::mutex> lock(mMutexPose);
    return mTcw;
}
SE3f KeyFrame::GetPoseInverse() {
    std::unique_lock<std::mutex> lock(mMutexPose);
    return mTwc;
}
Eigen::Vector3f KeyFrame::GetCameraCenter() {
    std::unique_lock<std::mutex> lock(mMutexPose);
    return mTwc.translation();
}
std::vector<MapPoint*> KeyFrame::GetMapPointMatches() {  // KeyFrame.cpp:285-289
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    return mvpMapPoints;
}
// KeyFrame::GetFeaturesInArea (KeyFrame.cpp:479-521) is the twin of Frame::GetFeaturesInArea over the grid the key frame
// copied from its frame (KeyFrame.cpp: mGrid(F.mGrid)); KeyFrame.cpp is not compiled, so a raw key frame answers through
// the reference's REAL Frame::GetFeaturesInArea of a shadow Frame with the same keypoints (kf_shadow, filled by the
// harness).
static std::map<const KeyFrame*, Frame*> kf_shadow;
std::vector<size_t> KeyFrame::GetFeaturesInArea(const float& x, const float& y, const float& r) const {
    return kf_shadow.at(this)->GetFeaturesInArea(x, y, r);
}
bool KeyFrame::isBad() {  // KeyFrame.cpp:456-460
    std::unique_lock<std::mutex> lock(mMutexConnections);
    return mbBad;
}
MapPoint* KeyFrame::GetMapPoint(const size_t& idx) {
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    return mvpMapPoints[idx];
}
// Pinhole's constructor creates one and its vtable refers to Reconstruct (ReconstructWithTwoViews); never used here
TwoViewReconstruction::TwoViewReconstruction(const Eigen::Matrix3f&, float, int) {}
bool TwoViewReconstruction::Reconstruct(const std::vector<KeyPointEx>&, const std::vector<KeyPointEx>&,
                                        const std::vector<int>&, SE3f&, std::vector<cv::Point3f>&, std::vector<bool>&) {
    return false;
}

namespace {
SE3f pose_of(const float* R, const float* t) {
    Eigen::Matrix3f Rm;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Rm(i, j) = R[3 * i + j];
    return SE3f(Rm, Eigen::Vector3f(t[0], t[1], t[2]));
}
KeyFrame* raw_keyframe(int n, const float* pos, const float* desc, const int* node, const unsigned char* has_mp,
                       MapPoint* some_point, const SE3f& Tcw) {
    KeyFrame* kf = static_cast<KeyFrame*>(calloc(1, sizeof(KeyFrame)));
    new (&kf->mFeatVec) DBoW3::FeatureVector();
    new (&kf->mvKeysUn) std::vector<KeyPointEx>();
    new (&kf->mDescriptors) cv::Mat();
    new (&kf->mvpMapPoints) std::vector<MapPoint*>();
    new (&kf->mTcw) SE3f(Tcw);
    new (&kf->mTwc) SE3f(Tcw.inverse());  // KeyFrame::SetPose, KeyFrame.cpp:38-40
    kf->N = n;
    kf->mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) kf->mvKeysUn[i] = KeyPointEx(pos[2 * i], pos[2 * i + 1], 1.f);
    kf->mDescriptors = cv::Mat(std::max(n, 1), 256, CV_32F);
    if (n > 0) memcpy(kf->mDescriptors.data, desc, (size_t)n * 1024);
    kf->mvpMapPoints.assign(n, nullptr);
    for (int i = 0; i < n; i++) {
        if (has_mp[i]) kf->mvpMapPoints[i] = some_point;
        if (node[i] >= 0) kf->mFeatVec[(unsigned int)node[i]].push_back((unsigned int)i);  // DBoW3 fills it in feature order
    }
    return kf;
}
// Key frame for the SearchByBoW functions: state[i] = 0 no map point, 1 a good one, 2 a bad one (MapPoint::isBad());
// the MapPoint objects come from the real constructor and are returned in `owned` (feature i -> owned[i] or nullptr).
KeyFrame* raw_keyframe_with_points(int n, const float* desc, const int* node, const unsigned char* state,
                                   std::vector<MapPoint*>& owned) {
    std::vector<float> pos(2 * (size_t)std::max(n, 1), 0.f);
    std::vector<unsigned char> none((size_t)std::max(n, 1), 0);
    KeyFrame* kf = raw_keyframe(n, pos.data(), desc, node, none.data(), nullptr, SE3f());
    owned.assign(n, nullptr);
    for (int i = 0; i < n; i++) {
        if (!state[i]) continue;
        MapPoint* mp = new MapPoint(Eigen::Vector3f(0.f, 0.f, 1.f), kf);  // the real constructor (reads two ids of kf)
        mp->mbBad = state[i] == 2;
        owned[i] = mp;
        kf->mvpMapPoints[i] = mp;
    }
    return kf;
}
void drop_keyframe(KeyFrame* kf) {
    kf->mFeatVec.~map();
    kf->mvKeysUn.~vector();
    kf->mDescriptors.~Mat();
    kf->mvpMapPoints.~vector();
    free(kf);
}
// One Matcher::SearchByProjection case on objects rebuilt from flat arrays; MatcherT = ::Matcher or ppg_shim::Matcher.
//   mode 0: SearchByProjection(CurrentFrame, LastFrame, th) (Matcher.cpp:31-87);
//   mode 1: SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, descDist) (:1337-1411);
//   mode 2: SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (:479-568; loop closing): the "current"
//     side is a raw KEY FRAME (keypoints, descriptors, the camera, a shadow Frame for its grid), kp_mp is vpMatched, the
//     source features with a point are vpPoints (state 2 = bad), Scw = (Rcw, tcw * scale, scale), desc_dist is
//     ratioHamming; normal = GetNormal() of the source points (the viewing-angle test of :522-525).
//   source side (LastFrame / pKF), n_src features: state 0 no map point, 1 a map point, 2 an outlier (mode 0) / a bad
//     point (mode 1), 3 in sAlreadyFound (mode 1); world_pos, mp_desc (GetDescriptor), observed (Observations() > 0),
//     min_dist / max_dist (Get*DistanceInvariance, mode 1).
//   CurrentFrame: pose Rcw / tcw, n keypoints, descriptors, kp_mp in / out: -1 none, k >= 0 the map point of source
//     feature k, -2 / -3 a map point outside the source with / without observations.
// Optional out, per source feature: whether the reference's tests let it search (row_valid) and its projection (proj_uv),
// by the same expressions as :43-56 / :1355-1371 -- what a caller hands to the C ABI.
template <class MatcherT>
int projection_case(MatcherT& matcher, GeometricCamera* cam, int mode, const float* Rcw, const float* tcw, int n_src,
                    const float* world_pos, const float* mp_desc, const unsigned char* state,
                    const unsigned char* observed, const float* min_dist, const float* max_dist, int n, const float* kx,
                    const float* ky, const float* desc, int* kp_mp, float th, float desc_dist, float* proj_uv,
                    unsigned char* row_valid, const float* normal = nullptr, float scale = 1.f) {
    KeyFrame* kf0 = static_cast<KeyFrame*>(calloc(1, sizeof(KeyFrame)));
    auto point = [&](const float* P, const float* d, bool obs, bool bad, float mn, float mx) {
        MapPoint* mp = new MapPoint(Eigen::Vector3f(P[0], P[1], P[2]), kf0);
        mp->nObs = obs ? 1 : 0;
        mp->mbBad = bad;
        mp->mfMinDepth = mn * 2.0f;
        mp->mfMaxDepth = mx * 0.5f;
        mp->mDescriptor = cv::Mat(1, 256, CV_32F);
        if (d) memcpy(mp->mDescriptor.data, d, 1024);
        return mp;
    };
    std::vector<MapPoint*> src(n_src, nullptr);
    for (int i = 0; i < n_src; i++)
        if (state[i])
            src[i] = point(world_pos + 3 * i, mp_desc + (size_t)i * 256, observed[i] != 0, mode >= 1 && state[i] == 2,
                           min_dist ? min_dist[i] : 0.f, max_dist ? max_dist[i] : 1e30f);
    if (normal)
        for (int i = 0; i < n_src; i++)
            if (src[i]) src[i]->mNormalVector = Eigen::Vector3f(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
    const float zero3[3] = {0, 0, 1};
    MapPoint* outside_obs = point(zero3, nullptr, true, false, 0.f, 1e30f);
    MapPoint* outside_unobs = point(zero3, nullptr, false, false, 0.f, 1e30f);

    Frame F;  // CurrentFrame
    F.N = n;
    F.mpCamera = cam;
    F.mTcw = pose_of(Rcw, tcw);
    F.mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) F.mvKeysUn[i] = KeyPointEx(kx[i], ky[i], 1.f);
    F.mvKeys = F.mvKeysUn;
    F.mDescriptors = cv::Mat(std::max(n, 1), 256, CV_32F);
    if (n > 0) memcpy(F.mDescriptors.data, desc, (size_t)n * 1024);
    F.mvpMapPoints.assign(n, nullptr);
    for (int i = 0; i < n; i++)
        F.mvpMapPoints[i] = kp_mp[i] >= 0 ? src[kp_mp[i]]
                                          : (kp_mp[i] == -2 ? outside_obs : (kp_mp[i] == -3 ? outside_unobs : nullptr));
    F.AssignFeaturesToGrid();

    if (row_valid && proj_uv) {  // the caller's side of the split: which source features search, and where
        SE3f Tcw = F.GetPose();
        if (mode == 2) {  // :482: the pose comes out of the similarity
            Eigen::Matrix3f Rm;
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) Rm(i, j) = Rcw[3 * i + j];
            Sim3f Scw(Rm, Eigen::Vector3f(tcw[0] * scale, tcw[1] * scale, tcw[2] * scale), scale);
            Tcw = SE3f(Scw.rotationMatrix(), Scw.translation() / Scw.scale());
        }
        Eigen::Vector3f Ow = Tcw.inverse().translation();
        for (int i = 0; i < n_src; i++) {
            row_valid[i] = 0;
            proj_uv[2 * i] = proj_uv[2 * i + 1] = -1.f;
            MapPoint* pMP = src[i];
            if (!pMP) continue;
            bool found = false;  // mode 2: spAlreadyFound = the points vpMatched holds on entry (:485-486)
            for (int k = 0; k < n && mode == 2; k++) found = found || kp_mp[k] == i;
            if (mode == 0 ? state[i] == 2 : (mode == 1 ? (pMP->isBad() || state[i] == 3) : (pMP->isBad() || found))) continue;
            Eigen::Vector3f x3Dw = pMP->GetWorldPos();
            Eigen::Vector3f x3Dc = Tcw * x3Dw;
            if (mode == 0) {
                const float invzc = 1.0 / x3Dc(2);
                if (invzc < 0) continue;
            }
            if (mode == 2 && x3Dc(2) < 0.0) continue;  // :499-500
            Eigen::Vector2f uv = cam->project(x3Dc);
            if (!cam->IsInImage(uv(0), uv(1))) continue;
            if (mode >= 1) {
                Eigen::Vector3f PO = x3Dw - Ow;
                float dist3D = PO.norm();
                if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
                if (mode == 2) {
                    Eigen::Vector3f Pn = pMP->GetNormal();
                    if (PO.dot(Pn) < 0.5 * dist3D) continue;  // :522-525
                }
            }
            row_valid[i] = 1;
            proj_uv[2 * i] = uv(0);
            proj_uv[2 * i + 1] = uv(1);
        }
    }

    int nm = 0;
    if (mode == 0) {
        Frame L;  // LastFrame
        L.N = n_src;
        L.mpCamera = cam;
        L.mvpMapPoints = src;
        L.mvbOutlier.assign(n_src, false);
        for (int i = 0; i < n_src; i++) L.mvbOutlier[i] = state[i] == 2;
        nm = matcher.SearchByProjection(F, L, th);
    } else if (mode == 2) {
        std::vector<float> pos(2 * (size_t)std::max(n, 1));
        std::vector<int> node((size_t)std::max(n, 1), -1);
        std::vector<unsigned char> none((size_t)std::max(n, 1), 0);
        for (int i = 0; i < n; i++) {
            pos[2 * i] = kx[i];
            pos[2 * i + 1] = ky[i];
        }
        KeyFrame* kf = raw_keyframe(n, pos.data(), desc, node.data(), none.data(), nullptr, SE3f());
        kf->mpCamera = cam;
        kf_shadow[kf] = &F;  // F carries the same keypoints and the grid (AssignFeaturesToGrid above)
        std::vector<MapPoint*> vpPoints, vpMatched(F.mvpMapPoints);
        for (int i = 0; i < n_src; i++)
            if (src[i]) vpPoints.push_back(src[i]);
        Eigen::Matrix3f Rm;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) Rm(i, j) = Rcw[3 * i + j];
        Sim3f Scw(Rm, Eigen::Vector3f(tcw[0] * scale, tcw[1] * scale, tcw[2] * scale), scale);
        nm = matcher.SearchByProjection(kf, Scw, vpPoints, vpMatched, (int)th, desc_dist);
        F.mvpMapPoints = vpMatched;
        kf_shadow.erase(kf);
        drop_keyframe(kf);
    } else {
        std::vector<float> pos(2 * (size_t)std::max(n_src, 1), 0.f), dsc(256 * (size_t)std::max(n_src, 1), 0.f);
        std::vector<int> node((size_t)std::max(n_src, 1), -1);
        std::vector<unsigned char> none((size_t)std::max(n_src, 1), 0);
        KeyFrame* kf = raw_keyframe(n_src, pos.data(), dsc.data(), node.data(), none.data(), nullptr, SE3f());
        kf->mvpMapPoints = src;
        std::set<MapPoint*> found;
        for (int i = 0; i < n_src; i++)
            if (state[i] == 3) found.insert(src[i]);
        nm = matcher.SearchByProjection(F, kf, found, th, desc_dist);
        drop_keyframe(kf);
    }
    std::map<MapPoint*, int> idx_of;
    for (int i = 0; i < n_src; i++)
        if (src[i]) idx_of[src[i]] = i;
    for (int i = 0; i < n; i++) {
        MapPoint* m = F.mvpMapPoints[i];
        kp_mp[i] = !m ? -1 : (m == outside_obs ? -2 : (m == outside_unobs ? -3 : idx_of[m]));
    }
    for (MapPoint* m : src) delete m;
    delete outside_obs;
    delete outside_unobs;
    free(kf0);
    return nm;
}
}  // namespace
