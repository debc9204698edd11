#define PPG_SHIM_NO_REFERENCE_HEADERS
#include "stub_types.h"
#include "ppg_shim.hpp"
int shim_instantiate(GeometricCamera* cam, Frame& F, std::vector<MapPoint*>& mps) {
    ppg_shim::PPGExtractor ex(cam, "net");
    std::vector<KeyPointEx> a, b;
    std::vector<KeyEdge> e;
    cv::Mat d;
    ex.run(cv::Mat(), a, b, e, d);
    ppg_shim::upload_map_descriptors(ex.context(), mps);
    std::vector<float> uv(2 * mps.size());
    std::vector<uint8_t> free_mask(F.mvKeysUn.size(), 1);
    int n = (int)ppg_shim::search_window(ex.context(), F, mps, uv, free_mask, 15.f, 0.8f).accept.size();        // :31-87
    n += (int)ppg_shim::search_window(ex.context(), F, mps, uv, free_mask, 3.f, 0.7f, 5.99).accept.size();     // Fuse
    ppg_shim::upload_map_geometry(ex.context(), mps);
    ppg_shim::check_in_frustum(ex.context(), F, mps, 0.5f, 10.f, 0.8f);  // Frame.cpp:223-260
    ppg_shim::compute_bow(ex.context(), F);  // Frame.cpp:331-340
    KeyFrame kf;
    std::vector<MapPoint*> bowm;
    n += ppg_shim::search_by_bow(ex.context(), &kf, F, bowm, 0.8f, 0.7f);  // Matcher.cpp:393-477
    n += ppg_shim::extend_map_matches(ex.context(), F, mps, 10.f, 0.8f);  // Matcher.cpp:203-381
    // class Matcher: all twelve signatures of matching/include/Matcher.h:20-64 resolve on the shim class
    ppg_shim::Matcher m(ex.context(), cam, 0.8f);
    std::vector<cv::Point2f> prev(F.mvKeysUn.size());
    std::vector<int> m12;
    std::vector<std::pair<size_t, size_t>> pairs;
    std::set<MapPoint*> found;
    Sim3f S;
    n += m.ExtendMapMatches(F, mps, 10.f) + m.SearchByBoW(&kf, F, bowm) + m.SearchForInitialization(F, F, prev, m12, 50);
    n += m.SearchByProjection(F, mps) + m.SearchByProjection(F, F, 15.f) + m.SearchByProjection(F, &kf, found, 3.f, 0.7f) +
         m.SearchByProjection(&kf, S, mps, bowm, 10);
    n += m.SearchByBoW(&kf, &kf, bowm) + m.SearchForTriangulation(&kf, &kf, pairs) + m.SearchBySim3(&kf, &kf, bowm, S, 7.5f);
    n += m.Fuse(&kf, mps) + m.Fuse(&kf, S, mps, 3.f, bowm);
    n += (m.mfNNratio > 0.f && ppg_shim::Matcher::TH_LOW < ppg_shim::Matcher::TH_HIGH && m.mpCamera == cam) ? 1 : 0;
    return n + (int)ppg_shim::search_local_points(ex.context(), F, mps, 10.f, 0.8f).accept.size();
}
