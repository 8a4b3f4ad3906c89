// TEST INFRASTRUCTURE ONLY -- see ../core/core.hpp.
#pragma once
#include <opencv2/core/core.hpp>
namespace cv {
inline Mat imread(const std::string& /*path*/, int /*flags*/) {
  Mat& installed = cvshim::prior_image();
  if (installed.rows > 0) return installed;
  Mat m(480, 640, CV_16UC1);
  for (int r = 0; r < m.rows; ++r)
    for (int c = 0; c < m.cols; ++c) m.at<unsigned short>(r, c) = 10000;
  return m;
}
}  // namespace cv
