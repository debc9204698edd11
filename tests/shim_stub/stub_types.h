// Minimal stand-ins for the reference / OpenCV / Eigen types the shim touches, ONLY so that
// tests/test_shim_compiles.py can syntax-check include/ppg_shim.hpp in a container without those libraries.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <utility>
#include <vector>
#define CV_32FC1 5
namespace cv {
struct Mat {
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    size_t step = 0;
    Mat() {}
    Mat(int r, int c, int) : rows(r), cols(c) {}
    int channels() const { return 1; }
    template <typename T> T& at(int, int = 0) { static T t; return t; }
    template <typename T> T* ptr(int = 0) { return nullptr; }
    template <typename T> const T* ptr(int = 0) const { return nullptr; }
};
struct Point2f {
    float x = 0, y = 0;
    Point2f() {}
    Point2f(float a, float b) : x(a), y(b) {}
};
}  // namespace cv
struct Vec2f {
    float v[2];
    float& operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    struct Init { Vec2f* p; int k; Init operator,(float x) { p->v[k] = x; return Init{p, k + 1}; } };
    Init operator<<(float x) { v[0] = x; return Init{this, 1}; }
};
// Eigen / SE3.h look-alikes for the pose-only geometry of ppg_shim::search_for_triangulation (syntax check only)
namespace Eigen {
struct Vector2f {
    float v[2];
    float& operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    float operator()(int i) const { return v[i]; }
};
struct Vector3f {
    float v[3];
    float operator[](int i) const { return v[i]; }
    float operator()(int i) const { return v[i]; }
    Vector3f operator-(const Vector3f&) const { return *this; }
    Vector3f operator/(float) const { return *this; }
    float dot(const Vector3f&) const { return 0.f; }
    float norm() const { return 0.f; }
};
struct Matrix3f {
    float m[9];
    float operator()(int r, int c) const { return m[3 * r + c]; }
    Matrix3f transpose() const { return *this; }
    Matrix3f inverse() const { return *this; }
    Matrix3f operator*(const Matrix3f&) const { return *this; }
};
}  // namespace Eigen
struct SE3f {
    SE3f() {}
    SE3f(const Eigen::Matrix3f&, const Eigen::Vector3f&) {}
    Eigen::Matrix3f rotationMatrix() const { return Eigen::Matrix3f(); }
    Eigen::Vector3f translation() const { return Eigen::Vector3f(); }
    SE3f operator*(const SE3f&) const { return *this; }
    SE3f inverse() const { return *this; }
    Eigen::Vector3f operator*(const Eigen::Vector3f& p) const { return p; }
};
struct SO3f {
    static Eigen::Matrix3f hat(const Eigen::Vector3f&) { return Eigen::Matrix3f(); }
};
struct KeyPointEx {
    KeyPointEx() {}
    KeyPointEx(float x, float y, float sc) : mfScore(sc), mbOut(true) { mPos.v[0] = x; mPos.v[1] = y; }
    Vec2f mPos, mPosUn;
    float mfScore;
    std::vector<unsigned int> mvConnected;
    std::vector<std::pair<unsigned int, unsigned int>> mvColine;
    bool mbOut;
};
struct KeyEdge {
    KeyEdge() {}
    KeyEdge(const unsigned int& a, const unsigned int& b) : startIdx(a), endIdx(b) {}
    unsigned int startIdx, endIdx;
    bool isBad;
    float lscore, length;
};
struct GeometricCamera {
    static const unsigned int CAM_PINHOLE = 0, CAM_FISHEYE = 1;
    virtual cv::Mat toK() = 0;
    virtual cv::Mat toD() = 0;
    virtual int imWidth() = 0;
    virtual int imHeight() = 0;
    virtual Eigen::Vector2f project(const Eigen::Vector3f&) = 0;
    virtual Eigen::Matrix3f toK_() = 0;
    bool IsInImage(const float&, const float&) const { return true; }
    unsigned int mnType;
    std::vector<float> mvParameters;  // sensors/include/GeometricCamera.h:87
};
struct MapPoint;
struct MapEdge {
    MapPoint *mpMPs, *mpMPe;
    bool mbBad, mbValid;
    MapPoint* theOtherPt(MapPoint* p) { return p == mpMPs ? mpMPe : (p == mpMPe ? mpMPs : nullptr); }
    bool isBad() { return mbBad; }
};
typedef Eigen::Vector3f Vec3f;
struct Mat3f {
    float m[9];
    float operator()(int r, int c) const { return m[3 * r + c]; }
};
struct MapPoint {
    float mTrackProjX, mTrackProjY, mTrackViewCos, mTrackDepth;
    Vec3f GetWorldPos() { return Vec3f(); }
    Vec3f GetNormal() { return Vec3f(); }
    float GetMinDistanceInvariance() { return 0.f; }
    float GetMaxDistanceInvariance() { return 0.f; }
    void IncreaseVisible(int = 1) {}
    bool mbTrackInView;
    long unsigned int mnTrackedbyFrame;
    int Observations() { return 0; }
    bool isBad() { return false; }
    std::vector<MapEdge*> getEdges() { return {}; }
    cv::Mat GetDescriptor() { return cv::Mat(); }
};
struct KeyFrame {
    GeometricCamera* mpCamera = nullptr;
    std::map<unsigned int, std::vector<unsigned int>> mFeatVec;
    cv::Mat mDescriptors;
    int N = 0;
    std::vector<KeyPointEx> mvKeysUn;
    std::vector<MapPoint*> GetMapPointMatches() { return {}; }
    MapPoint* GetMapPoint(const size_t&) { return nullptr; }
    SE3f GetPose() { return SE3f(); }
    SE3f GetPoseInverse() { return SE3f(); }
    Eigen::Vector3f GetCameraCenter() { return Eigen::Vector3f(); }
};
struct Frame {
    int N = 0;
    GeometricCamera* mpCamera = nullptr;
    std::vector<bool> mvbOutlier;
    SE3f GetPose() const { return SE3f(); }
    Mat3f mRcw;
    Vec3f mtcw, mOw;
    std::map<unsigned int, double> mBowVec;                      // DBoW3::BowVector
    std::map<unsigned int, std::vector<unsigned int>> mFeatVec;  // DBoW3::FeatureVector
    long unsigned int mnId;
    std::vector<KeyPointEx> mvKeysUn;
    std::vector<KeyEdge> mvKeyEdges;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<MapEdge*> mvpMapEdges;
    cv::Mat mDescriptors;
};

#include <set>
struct Sim3f {
    Eigen::Matrix3f rotationMatrix() const { return Eigen::Matrix3f(); }
    Eigen::Vector3f translation() const { return Eigen::Vector3f(); }
    float scale() const { return 1.f; }
};
class Matcher {  // matching/include/Matcher.h:20-64, signatures only
public:
    Matcher(GeometricCamera* pCam, float nnratio = 0.6) : mpCamera(pCam), mfNNratio(nnratio) {}
    int SearchByProjection(Frame&, const std::vector<MapPoint*>&, const float = 3) { return 0; }
    int SearchByProjection(Frame&, const Frame&, const float) { return 0; }
    int SearchByProjection(Frame&, KeyFrame*, const std::set<MapPoint*>&, const float, const float) { return 0; }
    int SearchByProjection(KeyFrame*, Sim3f&, const std::vector<MapPoint*>&, std::vector<MapPoint*>&, int, float = 1.0) { return 0; }
    int SearchByBoW(KeyFrame*, Frame&, std::vector<MapPoint*>&) { return 0; }
    int SearchByBoW(KeyFrame*, KeyFrame*, std::vector<MapPoint*>&) { return 0; }
    int SearchForInitialization(Frame&, Frame&, std::vector<cv::Point2f>&, std::vector<int>&, int = 10) { return 0; }
    int SearchForTriangulation(KeyFrame*, KeyFrame*, std::vector<std::pair<size_t, size_t>>&, const bool = false) { return 0; }
    int SearchBySim3(KeyFrame*, KeyFrame*, std::vector<MapPoint*>&, const Sim3f&, const float) { return 0; }
    int Fuse(KeyFrame*, const std::vector<MapPoint*>&, const float = 3.0) { return 0; }
    int Fuse(KeyFrame*, Sim3f&, const std::vector<MapPoint*>&, float, std::vector<MapPoint*>&) { return 0; }
    int ExtendMapMatches(Frame&, const std::vector<MapPoint*>&, const float) { return 0; }
    static constexpr float TH_LOW = 0.7f, TH_HIGH = 0.8f;
    GeometricCamera* mpCamera;
    float mfNNratio;
};
