// TEST INFRASTRUCTURE ONLY -- minimal stand-in for the OpenCV types the reference engine
// touches while reading the per-pixel prior image
// (/root/reference/src/3rdparty/super4pcs/src/super4pcs/algorithms/match4pcsBase.cc:317-338).
// The harness (oracle/ref_harness.cc) installs the image through cvshim::prior_image();
// cv::imread ignores the path and returns that image (or a 480x640 all-10000 image, i.e.
// prior 1.0 everywhere, if none was installed).  at<>() clamps, because the reference
// indexes the image without bounds checks (:335-338).
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_16UC1 2
#define CV_32FC1 5

namespace cv {
class Mat {
 public:
  int rows = 0, cols = 0;
  int elem = 1;
  std::shared_ptr<std::vector<unsigned char>> buf;
  Mat() {}
  Mat(int r, int c, int type) : rows(r), cols(c), elem(type == CV_16UC1 ? 2 : 4) {
    buf = std::make_shared<std::vector<unsigned char>>(size_t(r) * c * elem, 0);
  }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  template <typename T>
  T& at(int r, int c) {
    if (r < 0) r = 0;
    if (c < 0) c = 0;
    if (r >= rows) r = rows - 1;
    if (c >= cols) c = cols - 1;
    return reinterpret_cast<T*>(buf->data())[size_t(r) * cols + c];
  }
};
}  // namespace cv

namespace cvshim {
inline cv::Mat& prior_image() {
  static cv::Mat img;
  return img;
}
}  // namespace cvshim
