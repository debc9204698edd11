// opencv2/core/eigen.hpp stand-in: cv::cv2eigen for a CV_32F matrix.
#pragma once
#include "cv_standin.hpp"
#include "eigen_standin.hpp"
namespace cv {
template <typename S, int R, int C, int O, int MR, int MC>
inline void cv2eigen(const Mat& src, Eigen::Matrix<S, R, C, O, MR, MC>& dst) {
    dst.resize(src.rows, src.cols);
    for (int i = 0; i < src.rows; i++)
        for (int j = 0; j < src.cols; j++) dst(i, j) = (S)src.at<float>(i, j);
}
}  // namespace cv
