// TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or executed from the product path.
//
// C-ABI harness around the UNMODIFIED reference engine, compiled in place from
// /root/reference by oracle/Makefile into oracle/_ref/libs4ref.so.  It subclasses
// match_4pcs::MatchSuper4PCS to reach the protected members
// (S4/algorithms/match4pcsBase.h:132-274) and exposes, function by function, the pieces of
// the PCS -> LCP hot path that the CUDA kernels replace:
//
//   ref_verify_batch           Match4PCSBase::Verify            match4pcsBase.cc:1699-1731
//   ref_weighted_verify_batch  Match4PCSBase::WeightedVerify    match4pcsBase.cc:1733-1766
//   ref_rigid_from_quad        ComputeRigidTransformFromCongruentPair :1411-1488 (+ :1504-1614)
//   ref_extract_pairs          MatchSuper4PCS::ExtractPairs     super4pcs.cc:193-236
//   ref_find_quads             MatchSuper4PCS::FindCongruentQuadrilaterals super4pcs.cc:78-187
//   ref_select_quadrilateral   Match4PCSBase::SelectQuadrilateral :507-580  (operMode 0 bases)
//   ref_perform_n_steps        Match4PCSBase::Perform_N_steps   :1823-1927 (needs the patched TU)
//
// (S4 = /root/reference/src/3rdparty/super4pcs/src/super4pcs.)
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <map>
#include <streambuf>
#include <string>
#include <thread>
#include <vector>

#include "algorithms/super4pcs.h"

extern "C" unsigned pgp_oracle_stocs_seed;
unsigned pgp_oracle_stocs_seed = 1;

using match_4pcs::Point3D;
using match_4pcs::Quadrilateral;
typedef std::map<std::vector<int>, std::vector<std::pair<int, int>>> PPFMapT;

namespace {

// The engine prints to std::cout from init(); silence it for the duration of a call.
struct CoutSilencer {
  std::streambuf* old;
  struct NullBuf : std::streambuf {
    int overflow(int c) override { return c; }
  } nb;
  CoutSilencer() { old = std::cout.rdbuf(&nb); }
  ~CoutSilencer() { std::cout.rdbuf(old); }
};

struct Oracle : match_4pcs::MatchSuper4PCS {
  using Base = Super4PCS::Match4PCSBase;
  explicit Oracle(const match_4pcs::Match4PCSOptions& o) : match_4pcs::MatchSuper4PCS(o) {}
  using Base::allTransforms;
  using Base::base_3D_;
  using Base::baseSet;
  using Base::best_LCP_;
  using Base::best_lcp_index;
  using Base::centroid_P_;
  using Base::centroid_Q_;
  using Base::operMode;
  using Base::orig_probabilities_;
  using Base::Perform_N_steps;
  using Base::registered_indices;
  using Base::sampled_P_3D_;
  using Base::sampled_Q_3D_;
  using Base::validation_Q_3D;
  using Base::Verify;
  using Base::WeightedVerify;
  using Base::base_selection_time;
  using Base::congruent_set_extraction;
  using Base::congruent_set_verification;
  using Base::P_diameter_;
  using Base::max_base_diameter_;
  using Base::ExtractCongruentSet;

  std::vector<Point3D> P, Q, V;
  PPFMapT ppf;
};

std::vector<Point3D> make_cloud(const float* xyz, const float* nrm, int n) {
  std::vector<Point3D> out;
  out.reserve(n);
  for (int i = 0; i < n; ++i) {
    Point3D p(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    if (nrm) {
      Point3D::VectorType nn(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
      // set_normal() normalises (S4/shared4pcs.h:85-87); a zero normal would give NaN, the
      // reader zeroes those instead (S4/utils/geometry.h:56-82) -- mirror that.
      if (nn.squaredNorm() < 0.01f) {
        // leave the default zero normal
      } else {
        p.set_normal(nn);
      }
    }
    out.push_back(p);
  }
  return out;
}

Eigen::Matrix<float, 4, 4> mat_from_3x4(const float* T) {
  Eigen::Matrix<float, 4, 4> M;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) M(r, c) = T[4 * r + c];
  M(3, 0) = 0.f; M(3, 1) = 0.f; M(3, 2) = 0.f; M(3, 3) = 1.f;
  return M;
}

}  // namespace

extern "C" {

// Builds one separately-initialised matcher (the kd-tree keeps its traversal stack in a
// member, S4/accelerators/kdtree.h:311, so one instance per thread).  K9 = row-major 3x3
// intrinsics (nullable => identity-like pin-hole that maps everything to pixel (0,0)).
// prior_img (nullable) = rows x cols uint16, value/10000 = prior.
void* ref_create(const float* P_xyz, const float* P_nrm, int nP,
                 const float* Q_xyz, const float* Q_nrm, int nQ,
                 const float* V_xyz, const float* V_nrm, int nV,
                 double delta, const float* K9,
                 const uint16_t* prior_img, int rows, int cols,
                 unsigned srand_seed) {
  CoutSilencer quiet;
  match_4pcs::Match4PCSOptions opt;
  // same settings as S4/super4pcs_test.cc:91-99 with the file-scope defaults :20-37
  opt.overlap_estimation = 0.5;
  opt.sample_size = 400;
  opt.max_normal_difference = -1;
  opt.max_color_distance = -1;
  opt.max_time_seconds = 2;
  opt.delta = delta;
  Oracle* o = new Oracle(opt);
  o->P = make_cloud(P_xyz, P_nrm, nP);
  o->Q = make_cloud(Q_xyz, Q_nrm, nQ);
  o->V = make_cloud(V_xyz, V_nrm, nV);

  cv::Mat& img = cvshim::prior_image();
  if (prior_img && rows > 0 && cols > 0) {
    img = cv::Mat(rows, cols, CV_16UC1);
    std::memcpy(img.buf->data(), prior_img, size_t(rows) * cols * 2);
  } else {
    img = cv::Mat();
  }
  Eigen::Matrix3f K;
  if (K9) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) K(r, c) = K9[3 * r + c];
  } else {
    K << 1, 0, 0, 0, 1, 0, 0, 0, 1;
  }
  srand(srand_seed);
  o->init(o->P, o->Q, o->V, "unused.png", K, "obj", o->ppf, 0);
  return o;
}

void ref_destroy(void* h) { delete static_cast<Oracle*>(h); }

void ref_get_centroids(void* h, float* cP, float* cQ) {
  Oracle* o = static_cast<Oracle*>(h);
  for (int i = 0; i < 3; ++i) { cP[i] = o->centroid_P_[i]; cQ[i] = o->centroid_Q_[i]; }
}

void ref_get_priors(void* h, float* out) {
  Oracle* o = static_cast<Oracle*>(h);
  std::copy(o->orig_probabilities_.begin(), o->orig_probabilities_.end(), out);
}

float ref_get_diameter(void* h) { return static_cast<Oracle*>(h)->P_diameter_; }

// centred clouds as the engine holds them after init() (S4/.../match4pcsBase.cc:242-268)
void ref_get_centred(void* h, int which, float* xyz, float* nrm) {
  Oracle* o = static_cast<Oracle*>(h);
  const std::vector<Point3D>& c = which == 0 ? o->sampled_P_3D_ : which == 1 ? o->sampled_Q_3D_ : o->validation_Q_3D;
  for (size_t i = 0; i < c.size(); ++i)
    for (int k = 0; k < 3; ++k) {
      if (xyz) xyz[3 * i + k] = c[i].pos()[k];
      if (nrm) nrm[3 * i + k] = c[i].normal()[k];
    }
}

// Verify with best_LCP_ forced to 0 before every call (no early termination): full counts.
// T = n row-major 3x4 centred-frame transforms.  counts[i] = round(fraction * |Qval|).
void ref_verify_batch(void* h, const float* T, int64_t n, uint32_t* counts) {
  Oracle* o = static_cast<Oracle*>(h);
  const float nv = float(o->validation_Q_3D.size());
  for (int64_t i = 0; i < n; ++i) {
    o->best_LCP_ = 0.f;
    float f = o->Verify(mat_from_3x4(T + 12 * i));
    counts[i] = uint32_t(f * nv + 0.5f);
  }
  o->best_LCP_ = 0.f;
}

// Verify exactly as Perform_N_steps drives it (running best => early termination active).
// Returns the raw fractions; used only to show that early exit never changes the arg-max.
void ref_verify_running_best(void* h, const float* T, int64_t n, float* frac, int64_t* best_index) {
  Oracle* o = static_cast<Oracle*>(h);
  o->best_LCP_ = 0.f;
  int64_t best = -1;
  for (int64_t i = 0; i < n; ++i) {
    float f = o->Verify(mat_from_3x4(T + 12 * i));
    frac[i] = f;
    if (f > o->best_LCP_) { o->best_LCP_ = f; best = i; }
  }
  *best_index = best;
  o->best_LCP_ = 0.f;
}

// WeightedVerify: scores[i] = sum(prior[resId] over gated in-range points)/|Qval|,
// nreg[i] = number of registered indices.  If reg_out != NULL, the registered scene indices
// of hypothesis `reg_of` are written there (capacity |Qval|).
void ref_weighted_verify_batch(void* h, const float* T, int64_t n, float* scores, int32_t* nreg,
                               int64_t reg_of, int32_t* reg_out) {
  Oracle* o = static_cast<Oracle*>(h);
  std::vector<int> reg;
  for (int64_t i = 0; i < n; ++i) {
    reg.clear();
    scores[i] = o->WeightedVerify(mat_from_3x4(T + 12 * i), reg);
    if (nreg) nreg[i] = int32_t(reg.size());
    if (reg_out && i == reg_of) std::copy(reg.begin(), reg.end(), reg_out);
  }
}

// Multi-thread CPU baseline: `nthreads` separately-initialised matchers (handles[]), static
// interleaved partition of the hypotheses, Verify with full counts.  Returns wall seconds.
double ref_verify_batch_mt(void** handles, int nthreads, const float* T, int64_t n, uint32_t* counts) {
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) {
    th.emplace_back([=]() {
      Oracle* o = static_cast<Oracle*>(handles[t]);
      const float nv = float(o->validation_Q_3D.size());
      for (int64_t i = t; i < n; i += nthreads) {
        o->best_LCP_ = 0.f;
        float f = o->Verify(mat_from_3x4(T + 12 * i));
        counts[i] = uint32_t(f * nv + 0.5f);
      }
      o->best_LCP_ = 0.f;
    });
  }
  for (auto& x : th) x.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ComputeRigidTransformFromCongruentPair for one (base, quad).  Outputs: T16 = centred 4x4
// (column-major, as Eigen stores it), pose16 = un-centred Isometry3d matrix (column-major
// doubles).  Returns 1 if a transform was pushed, 0 if rejected.
int ref_rigid_from_quad(void* h, const int* base, const int* quad, float* T16, double* pose16) {
  Oracle* o = static_cast<Oracle*>(h);
  std::vector<std::pair<Eigen::Isometry3d, float>> poses;
  size_t before = o->allTransforms.size();
  Quadrilateral q(quad[0], quad[1], quad[2], quad[3]);
  o->ComputeRigidTransformFromCongruentPair(base[0], base[1], base[2], base[3], q, poses);
  if (o->allTransforms.size() == before) return 0;
  Eigen::Matrix<float, 4, 4> M = o->allTransforms.back();
  o->allTransforms.pop_back();
  std::memcpy(T16, M.data(), sizeof(float) * 16);
  std::memcpy(pose16, poses.back().first.matrix().data(), sizeof(double) * 16);
  return 1;
}

// ExtractPairs over the centred search model Q.  Returns the number of ordered pairs; writes
// at most cap of them as (first, second) int pairs.
int64_t ref_extract_pairs(void* h, float pair_distance, float eps, int32_t* pairs, int64_t cap) {
  Oracle* o = static_cast<Oracle*>(h);
  std::vector<std::pair<int, int>> out;
  // base_point1/2 only matter when max_angle > 0 (S4/pairCreationFunctor.h:238-248); use 0,1.
  o->base_3D_[0] = o->sampled_P_3D_[0];
  o->base_3D_[1] = o->sampled_P_3D_[std::min<size_t>(1, o->sampled_P_3D_.size() - 1)];
  o->ExtractPairs(pair_distance, 0.f, eps, 0, 1, &out, std::vector<int>(4, 0));
  int64_t n = int64_t(out.size());
  for (int64_t i = 0; i < std::min(n, cap); ++i) { pairs[2 * i] = out[i].first; pairs[2 * i + 1] = out[i].second; }
  return n;
}

// FindCongruentQuadrilaterals for base scene ids b[4] (sets base_3D_, which the function
// reads for the base angle, S4/algorithms/super4pcs.cc:109-111).
int64_t ref_find_quads(void* h, const int* b, float inv1, float inv2, float eps,
                       const int32_t* pairs1, int64_t n1, const int32_t* pairs2, int64_t n2,
                       int32_t* quads, int64_t cap) {
  Oracle* o = static_cast<Oracle*>(h);
  for (int k = 0; k < 4; ++k) o->base_3D_[k] = o->sampled_P_3D_[b[k]];
  std::vector<std::pair<int, int>> A(n1), B(n2);
  for (int64_t i = 0; i < n1; ++i) A[i] = std::make_pair(pairs1[2 * i], pairs1[2 * i + 1]);
  for (int64_t i = 0; i < n2; ++i) B[i] = std::make_pair(pairs2[2 * i], pairs2[2 * i + 1]);
  std::vector<Quadrilateral> out;
  o->FindCongruentQuadrilaterals(inv1, inv2, eps, eps, A, B, &out);
  int64_t n = int64_t(out.size());
  for (int64_t i = 0; i < std::min(n, cap); ++i)
    for (int k = 0; k < 4; ++k) quads[4 * i + k] = out[i][k];
  return n;
}

// SelectQuadrilateral (operMode 0 base selection; deterministic given srand()).
int ref_select_quadrilateral(void* h, unsigned seed, int* b, float* inv) {
  Oracle* o = static_cast<Oracle*>(h);
  srand(seed);
  float i1 = 0, i2 = 0;
  bool ok = o->SelectQuadrilateral(i1, i2, b[0], b[1], b[2], b[3]);
  inv[0] = i1; inv[1] = i2;
  return ok ? 1 : 0;
}

// Full pipeline init -> Perform_N_steps in the given operMode (0 = Super4PCS pairs + Verify).
// Requires the TU built from the one-line-patched copy of match4pcsBase.cc (oracle/Makefile).
// Outputs the improving chain (poses as column-major double 4x4 + scores), the centred
// transforms that were verified, and the stage timers.
int ref_perform_n_steps(void* h, int mode, unsigned seed, double* chain_pose16, float* chain_score,
                        int cap, int64_t* n_transforms, float* best_lcp, int* best_index, float* stage_s) {
  CoutSilencer quiet;
  Oracle* o = static_cast<Oracle*>(h);
  o->operMode = mode;
  srand(seed);
  pgp_oracle_stocs_seed = seed * 7919u + 12345u;      // operMode 1: the engine seeds of this run's successive base draws (see Makefile)
  std::vector<std::pair<Eigen::Isometry3d, float>> allPose;
  o->Perform_N_steps(&o->Q, allPose, "/nonexistent/", "obj");
  int n = int(allPose.size());
  for (int i = 0; i < std::min(n, cap); ++i) {
    std::memcpy(chain_pose16 + 16 * i, allPose[i].first.matrix().data(), sizeof(double) * 16);
    chain_score[i] = allPose[i].second;
  }
  *n_transforms = int64_t(o->allTransforms.size());
  *best_lcp = o->best_LCP_;
  *best_index = o->best_lcp_index;
  stage_s[0] = o->base_selection_time;
  stage_s[1] = o->congruent_set_extraction;
  stage_s[2] = o->congruent_set_verification;
  return n;
}

// Centred transforms the last ref_perform_n_steps verified (row-major 3x4 each).
void ref_get_transforms(void* h, float* T, int64_t cap) {
  Oracle* o = static_cast<Oracle*>(h);
  int64_t n = std::min<int64_t>(cap, int64_t(o->allTransforms.size()));
  for (int64_t i = 0; i < n; ++i)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) T[12 * i + 4 * r + c] = o->allTransforms[i](r, c);
}

int64_t ref_get_registered(void* h, int32_t* out, int64_t cap) {
  Oracle* o = static_cast<Oracle*>(h);
  int64_t n = int64_t(o->registered_indices.size());
  for (int64_t i = 0; i < std::min(n, cap); ++i) out[i] = o->registered_indices[i];
  return n;
}

int ref_get_bases(void* h, int32_t* ids, float* inv, int cap) {
  Oracle* o = static_cast<Oracle*>(h);
  int n = int(o->baseSet.size());
  for (int i = 0; i < std::min(n, cap); ++i) {
    for (int k = 0; k < 4; ++k) ids[4 * i + k] = o->baseSet[i]->baseIds_[k];
    inv[2 * i] = o->baseSet[i]->invariant1_;
    inv[2 * i + 1] = o->baseSet[i]->invariant2_;
  }
  return n;
}

// ---- operMode 1 (StoCS + PPF map) pieces.  The map object lives in the Oracle (init() stored a pointer to it, :343).
void ref_set_ppf_map(void* h, const int32_t* keys4, const int64_t* offsets, const int32_t* pairs, int64_t n_keys) {
  Oracle* o = static_cast<Oracle*>(h);
  o->ppf.clear();
  for (int64_t k = 0; k < n_keys; ++k) {
    std::vector<int> key(keys4 + 4 * k, keys4 + 4 * k + 4);
    std::vector<std::pair<int, int>> v;
    for (int64_t e = offsets[k]; e < offsets[k + 1]; ++e) v.push_back(std::make_pair(pairs[2 * e], pairs[2 * e + 1]));
    o->ppf.insert(std::make_pair(key, v));
  }
}

// Match4PCSBase::computePPF (:582-598) of n index pairs of the scene cloud P
void ref_compute_ppf(void* h, const int32_t* pairs, int64_t n, int32_t* keys4) {
  Oracle* o = static_cast<Oracle*>(h);
  for (int64_t t = 0; t < n; ++t) {
    int i = pairs[2 * t], j = pairs[2 * t + 1];
    std::vector<int> k;
    o->computePPF(i, j, k);
    for (int c = 0; c < 4; ++c) keys4[4 * t + c] = k[c];
  }
}

// SelectQuadrilateralStoCS (:600-792) with the engine seed pinned (oracle/Makefile, second patch)
int ref_select_stocs(void* h, unsigned engine_seed, int* b, float* inv) {
  Oracle* o = static_cast<Oracle*>(h);
  pgp_oracle_stocs_seed = engine_seed;
  float i1 = 0, i2 = 0, prob = 0;
  bool ok = o->SelectQuadrilateralStoCS(i1, i2, b[0], b[1], b[2], b[3], prob);
  inv[0] = i1; inv[1] = i2;
  return ok ? 1 : 0;
}

// ExtractCongruentSet (:1929-2039) in operMode 1: pair lists from the PPF map, then FindCongruentQuadrilaterals
int64_t ref_congruent_set_mode1(void* h, const int* b, float inv1, float inv2, int32_t* quads, int64_t cap) {
  Oracle* o = static_cast<Oracle*>(h);
  o->operMode = 1;
  std::vector<int> ids(b, b + 4);
  Super4PCS::BaseGraph g(ids, inv1, inv2, 1.0f);
  o->ExtractCongruentSet(&g);
  int64_t n = int64_t(g.congruent_quads.size());
  for (int64_t i = 0; i < std::min(n, cap); ++i) {
    quads[4 * i] = g.congruent_quads[i].vertices[0]; quads[4 * i + 1] = g.congruent_quads[i].vertices[1];
    quads[4 * i + 2] = g.congruent_quads[i].vertices[2]; quads[4 * i + 3] = g.congruent_quads[i].vertices[3];
  }
  return n;
}

// ExtractCongruentSet (:1929-2039) in operMode 2: six pair extractions + FindCongruentQuadrilateralsV4PCS (:978-1044)
int64_t ref_congruent_set_mode2(void* h, const int* b, int32_t* quads, int64_t cap) {
  Oracle* o = static_cast<Oracle*>(h);
  o->operMode = 2;
  std::vector<int> ids(b, b + 4);
  Super4PCS::BaseGraph g(ids, 0.f, 0.f, 1.0f);
  o->ExtractCongruentSet(&g);
  int64_t n = int64_t(g.congruent_quads.size());
  for (int64_t i = 0; i < std::min(n, cap); ++i) {
    quads[4 * i] = g.congruent_quads[i].vertices[0]; quads[4 * i + 1] = g.congruent_quads[i].vertices[1];
    quads[4 * i + 2] = g.congruent_quads[i].vertices[2]; quads[4 * i + 3] = g.congruent_quads[i].vertices[3];
  }
  return n;
}

// SelectTetrahedronBase (:466-503) after srand(seed): widest random triangle + the most voluminous of 100 random 4th points
int ref_select_tetrahedron(void* h, unsigned seed, int* b) {
  Oracle* o = static_cast<Oracle*>(h);
  srand(seed);
  Oracle::Scalar i1 = 0, i2 = 0;
  return o->SelectTetrahedronBase(i1, i2, b[0], b[1], b[2], b[3]) ? 1 : 0;
}

}  // extern "C"
