// Reference-pinning harness, part 1: the reference's OWN extractor source compiled as it lies under /root/reference
// (feature/src/PPGExtractor.cpp, feature/src/PPGGraph.cpp, sensors/src/GeometricCamera.cpp; REF_ROOT is given by
// oracle/ref_build.py) against the real LibTorch (CPU) and the stand-ins of oracle/ref_standins for OpenCV / Eigen /
// DBoW3.  TEST INFRASTRUCTURE ONLY: it produces the vectors that pin oracle/ppg_oracle.c (tests/test_ref_pin.py) and the
// committed fixtures tests/golden/ref_l1_*.npz; nothing in ppg_slam_b200/ links or loads it.
//
// The source is not modified.  Three preprocessor shims let it run without a GPU and let the harness reach the
// tensors between the stages:
//   kCUDA -> kCPU            the hard-coded `torch::Device(torch::kCUDA, 0)` (PPGExtractor.cpp:41)
//   cuda  -> cpu             `torch::cuda::synchronize()` (:125, a no-op here) and `allPoints.cuda()` (:533)
//   private -> public        junctions / heatmap / descriptors / junc_pred / heatmap_score / eigenHeat are private members
// All standard / torch headers are included BEFORE the shims, so the shims only touch the reference's own text.
#include <algorithm>
#include <atomic>
#include <cassert>
#include <chrono>
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

#include <torch/nn/functional.h>
#include <torch/script.h>
#include <torch/torch.h>

#include "cv_eigen_standin.hpp"
#include "DBoW3/DBoW3.h"

namespace torch {
namespace cpu {
inline void synchronize(int64_t = -1) {}
}  // namespace cpu
}  // namespace torch

#define private public
#define protected public
#define kCUDA kCPU
#define cuda cpu
#include REF_FILE(feature/src/PPGExtractor.cpp)
#include REF_FILE(feature/src/PPGGraph.cpp)
#include REF_FILE(sensors/src/GeometricCamera.cpp)
#undef cuda
#undef kCUDA
#undef protected
#undef private

namespace {

// GeometricCamera is abstract; toK / toD / imWidth / imHeight as sensors/src/Pinhole.cpp:68-87 and
// KannalaBrandt8.cpp:136-146 define them, the projection functions (not used by the extractor) as stubs.
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

struct Harness {
    HarnessCamera* cam = nullptr;
    PPGExtractor* ex = nullptr;
};

}  // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

// params: fx fy cx cy d0 d1 d2 d3 (System.cpp:45-70 order); model_dir holds the four TorchScript files.
REF_API void* ref_extractor_create(const float* params8, int width, int height, int fisheye, const char* model_dir,
                                   int threads) {
    try {
        if (threads > 0) torch::set_num_threads(threads);
        Harness* h = new Harness();
        h->cam = new HarnessCamera(std::vector<float>(params8, params8 + 8), width, height, fisheye != 0);
        h->ex = new PPGExtractor(h->cam, std::string(model_dir));
        return h;
    } catch (const std::exception& e) {
        std::cerr << "ref_extractor_create: " << e.what() << std::endl;
        return nullptr;
    }
}

REF_API void ref_extractor_destroy(void* hp) {
    Harness* h = static_cast<Harness*>(hp);
    if (!h) return;
    delete h->ex;
    delete h->cam;
    delete h;
}

// GeometricCamera::InitializeImageBounds (sensors/src/GeometricCamera.cpp:26-61): minX minY maxX maxY, then the two
// inverse cell sizes.
REF_API void ref_image_bounds(void* hp, int* mm4, float* inv2) {
    Harness* h = static_cast<Harness*>(hp);
    mm4[0] = h->cam->mnMinX;
    mm4[1] = h->cam->mnMinY;
    mm4[2] = h->cam->mnMaxX;
    mm4[3] = h->cam->mnMaxY;
    inv2[0] = h->cam->mfGridElementWidthInv;
    inv2[1] = h->cam->mfGridElementHeightInv;
}

// One frame through PPGExtractor::inference + detectKeyPoint + detectLines + genPointDescriptor, i.e. run() without its
// final copies (PPGExtractor.cpp:118-147).  Dense maps the stages consumed are handed out so that the oracle can be fed
// EXACTLY what the reference saw:
//   prob  H x W   softmax + pixel_shuffle junction map (junc_pred, :161-162)
//   heat_raw H x W  softmax(heatmap)[:,1] before refineHeatMap (:242)
//   heat_ref H x W  after refineHeatMap, before remap (heatmap_score after :243-256)
//   heat_final H x W  eigenHeat (:259-263)
//   desc 256 x Hc x Wc  raw dense descriptors
// Returns the number of keypoints, or -1.  If prob_in / heat_in / desc_in are given they REPLACE the network outputs
// (logits are not needed: the tensors are substituted after the softmax by running the stages on prepared members).
REF_API int ref_extract(void* hp, const unsigned char* gray, float* prob, float* heat_raw, float* heat_ref,
                        float* heat_final, float* desc) {
    Harness* h = static_cast<Harness*>(hp);
    PPGExtractor& ex = *h->ex;
    try {
        const int H = ex.mnImHeight, W = ex.mnImWidth;
        cv::Mat img(H, W, CV_8UC1);
        memcpy(img.data, gray, (size_t)H * W);
        ex.inference(img);
        if (heat_raw) {
            torch::Tensor s = torch::softmax(ex.heatmap, 1).select(1, 1)[0].contiguous().cpu();
            memcpy(heat_raw, s.data_ptr<float>(), (size_t)H * W * 4);
        }
        if (desc) {
            torch::Tensor d = ex.descriptors[0].contiguous().cpu();
            memcpy(desc, d.data_ptr<float>(), (size_t)d.numel() * 4);
        }
        ex.detectKeyPoint();
        if (prob) {
            torch::Tensor p = ex.junc_pred.contiguous();
            memcpy(prob, p.data_ptr<float>(), (size_t)H * W * 4);
        }
        ex.detectLines();
        if (!ex.mvKeyPoints.empty()) {
            if (heat_ref) {
                torch::Tensor r = ex.heatmap_score.contiguous();
                memcpy(heat_ref, r.data_ptr<float>(), (size_t)H * W * 4);
            }
            if (heat_final)
                for (int y = 0; y < H; y++)
                    for (int x = 0; x < W; x++) heat_final[(size_t)y * W + x] = ex.eigenHeat(y, x);
        }
        ex.genPointDescriptor();
        return (int)ex.mvKeyPoints.size();
    } catch (const std::exception& e) {
        std::cerr << "ref_extract: " << e.what() << std::endl;
        return -1;
    }
}

REF_API int ref_counts(void* hp, int* n_edges, int* n_conn, int* n_coline) {
    PPGExtractor& ex = *static_cast<Harness*>(hp)->ex;
    int nc = 0, nl = 0;
    for (const KeyPointEx& k : ex.mvKeyPoints) {
        nc += (int)k.mvConnected.size();
        nl += (int)k.mvColine.size();
    }
    *n_edges = (int)ex.mvKeyEdges.size();
    *n_conn = nc;
    *n_coline = nl;
    return (int)ex.mvKeyPoints.size();
}

// Keypoint fields (KeyPointEx, sensors/include/GeometricCamera.h:22-37) and the graph as flat arrays.
REF_API void ref_fetch(void* hp, float* pos_xy, float* posun_xy, float* score, unsigned char* out, int* edge_se,
                       float* edge_lscore, int* conn_off, int* conn_idx, int* col_off, int* col_pairs, float* normdesc) {
    PPGExtractor& ex = *static_cast<Harness*>(hp)->ex;
    const int n = (int)ex.mvKeyPoints.size();
    int co = 0, lo = 0;
    for (int i = 0; i < n; i++) {
        const KeyPointEx& k = ex.mvKeyPoints[i];
        pos_xy[2 * i] = k.mPos[0];
        pos_xy[2 * i + 1] = k.mPos[1];
        posun_xy[2 * i] = k.mPosUn[0];
        posun_xy[2 * i + 1] = k.mPosUn[1];
        score[i] = k.mfScore;
        out[i] = k.mbOut ? 1 : 0;
        conn_off[i] = co;
        for (unsigned int e : k.mvConnected) conn_idx[co++] = (int)e;
        col_off[i] = lo;
        for (const auto& pr : k.mvColine) {
            col_pairs[2 * lo] = (int)pr.first;
            col_pairs[2 * lo + 1] = (int)pr.second;
            lo++;
        }
    }
    conn_off[n] = co;
    col_off[n] = lo;
    for (size_t e = 0; e < ex.mvKeyEdges.size(); e++) {
        edge_se[2 * e] = (int)ex.mvKeyEdges[e].startIdx;
        edge_se[2 * e + 1] = (int)ex.mvKeyEdges[e].endIdx;
        edge_lscore[e] = ex.mvKeyEdges[e].lscore;
    }
    if (normdesc && n > 0) {
        torch::Tensor d = ex.normDesc.contiguous();
        memcpy(normdesc, d.data_ptr<float>(), (size_t)n * 256 * 4);
    }
}
