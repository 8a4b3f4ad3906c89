// ABI stand-ins for the three Eigen types in the signature of getProbableTransformsSuper4PCS
// (S4/super4pcs_test.cc:39-43; caller-side declaration PPE/src/hypothesis_generation/
// ObjectPoseCandidateSet.cpp:5-9).  Eigen is not installed in this image and the reference's vendored
// copy may not be copied, so the drop-in library is built against these layout- and
// mangling-compatible declarations:
//   Eigen::Isometry3d = Eigen::Transform<double,3,Eigen::Isometry(=1),0>: one column-major 4x4 double
//                       matrix, 128 bytes, 16-byte aligned (EIGEN_MAX_STATIC_ALIGN_BYTES = 16, i.e. a
//                       build without -mavx, as the reference's catkin build is)
//   Eigen::Matrix3f   = Eigen::Matrix<float,3,3,0,3,3>: nine column-major floats, 4-byte aligned
// Both have user-provided copy constructors like the real classes, so they are passed the same way
// (by invisible reference) under the Itanium C++ ABI.  Building with -DPGP_USE_REAL_EIGEN and
// -I<path to Eigen> swaps in the real headers; tests/test_adaptor.py checks that both variants
// produce the identical mangled symbol.
#pragma once
#ifdef PGP_USE_REAL_EIGEN
#include <Eigen/Core>
#include <Eigen/Geometry>
#else
#include <cstring>
namespace Eigen {
template <typename Scalar, int Dim, int Mode, int Options> class Transform;
template <typename Scalar, int Rows, int Cols, int Options, int MaxRows, int MaxCols> class Matrix;

template <> class alignas(16) Transform<double, 3, 1, 0> {
 public:
  Transform() { setIdentity(); }
  Transform(const Transform& o) { std::memcpy(m_, o.m_, sizeof(m_)); }
  Transform& operator=(const Transform& o) { std::memcpy(m_, o.m_, sizeof(m_)); return *this; }
  void setIdentity() { for (int i = 0; i < 16; ++i) m_[i] = (i % 5 == 0) ? 1.0 : 0.0; }
  double& operator()(int r, int c) { return m_[4 * c + r]; }          // column-major
  double operator()(int r, int c) const { return m_[4 * c + r]; }
 private:
  double m_[16];
};
template <> class Matrix<float, 3, 3, 0, 3, 3> {
 public:
  Matrix() { for (float& v : m_) v = 0.f; }
  Matrix(const Matrix& o) { std::memcpy(m_, o.m_, sizeof(m_)); }
  Matrix& operator=(const Matrix& o) { std::memcpy(m_, o.m_, sizeof(m_)); return *this; }
  float& operator()(int r, int c) { return m_[3 * c + r]; }
  float operator()(int r, int c) const { return m_[3 * c + r]; }
 private:
  float m_[9];
};
typedef Transform<double, 3, 1, 0> Isometry3d;
typedef Matrix<float, 3, 3, 0, 3, 3> Matrix3f;
}  // namespace Eigen
static_assert(sizeof(Eigen::Isometry3d) == 128 && alignof(Eigen::Isometry3d) == 16, "Isometry3d layout");
static_assert(sizeof(Eigen::Matrix3f) == 36 && alignof(Eigen::Matrix3f) == 4, "Matrix3f layout");
#endif
