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
}  // namespace

