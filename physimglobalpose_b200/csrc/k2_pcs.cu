// K2 / K2b -- congruent-set (4PCS / Super4PCS) hypothesis generation on the device.
//
//   base selection   Match4PCSBase::SelectQuadrilateral / SelectRandomTriangle / TryQuadrilateral
//                    S4/algorithms/match4pcsBase.cc:507-580, 377-410, 415-464, distSegmentToSegment :81-148
//   pair extraction  MatchSuper4PCS::ExtractPairs  S4/algorithms/super4pcs.cc:193-236; the accepting filter is
//                    PairCreationFunctor::process  S4/pairCreationFunctor.h:167-253 (= brute force of 4pcs.cc:109-192)
//   quad join        MatchSuper4PCS::FindCongruentQuadrilaterals  S4/algorithms/super4pcs.cc:78-187 with
//                    IndexedNormalSet<Point,3,7>  S4/accelerators/normalset.h:71-151, normalset.hpp:114-214
//   transforms       ComputeRigidTransformFromCongruentPair :1411-1488 + ComputeRigidTransformation :1504-1614
//   driver           Perform_N_steps :1831-1877 (bases -> quads -> transforms; the verification loop is K3/K4)
//
// The hierarchical sphere rasterisation of the reference's pair extraction is only a candidate
// generator for the exact filter, so an all-pairs sweep (2k points -> 2M distance evaluations)
// yields the same pair SET.  The quad join reproduces the reference's quantisation (power-of-two
// position grid in unit-cube coordinates, 7^3 direction grid, rasterised cone) so that the quad set
// is the reference's, not merely a superset.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "pgp_internal.cuh"

namespace {

constexpr int NG = 7;                        // direction grid cells per axis (normalset: _ngSize)
__host__ __device__ inline uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

// ------------------------------------------------------------------------------- pair extraction
// thread i owns the pairs (j, i), j < i.  FILL = false: count;  true: write (j,i),(i,j) at the scanned offset.
template <bool FILL>
__global__ void __launch_bounds__(256) k2_pairs(const float4* __restrict__ Q, int nq, float dist, float eps, uint32_t* __restrict__ cnt,
                                                int2* __restrict__ out, long long cap) {
  __shared__ float4 tile[256];
  const int i = blockIdx.x * 256 + threadIdx.x;
  const float4 qi = i < nq ? Q[i] : make_float4(0, 0, 0, 0);
  const double d = (double)dist, e = (double)eps;      // pair_distance / pair_distance_epsilon are doubles (pairCreationFunctor.h:38-39)
  uint32_t n = 0;
  long long w = FILL && i < nq ? 2ll * cnt[i] : 0;
  const int jmax = min(nq, (int)(blockIdx.x + 1) * 256);
  for (int j0 = 0; j0 < jmax; j0 += 256) {
    __syncthreads();
    if (j0 + (int)threadIdx.x < nq) tile[threadIdx.x] = Q[j0 + threadIdx.x];
    __syncthreads();
    const int m = i < nq ? min(256, i - j0) : 0;        // only j < i
    for (int t = 0; t < m; ++t) {
      const float4 p = tile[t];
      const float dx = __fsub_rn(qi.x, p.x), dy = __fsub_rn(qi.y, p.y), dz = __fsub_rn(qi.z, p.z);
      const float dd = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dz, dz))));   // (q - p).norm()
      if (fabs((double)dd - d) > e) continue;           // std::abs(distance - pair_distance) > pair_distance_epsilon
      if (FILL) {
        if (w + 1 < cap) { out[w] = make_int2(j0 + t, i); out[w + 1] = make_int2(i, j0 + t); }
        w += 2;
      }
      ++n;
    }
  }
  if (!FILL && i < nq) cnt[i] = n;
}

// ------------------------------------------------------------------------------- quad join
struct JoinParams {
  const float4* Qn;      // model points in the unit cube (worldToUnit, pairCreationFunctor.h:76-80)
  const float4* Q;       // centred model points
  const int2* A; long long n1;
  const int2* B; long long n2;
  float inv1, inv2, cos_alpha, thr2;
  float cell;            // 1 / egSize
  int eg;                // position grid cells per axis (power of two)
  uint32_t n_buckets;    // power of two
};

__device__ __forceinline__ int pos_cell(const JoinParams& p, float x, float y, float z) {
  // coordinatesPos = p / _epsilon ; index = int(x) + int(y) g + int(z) g^2   (accelerators/utils.h:141-148)
  const int cx = (int)__fdiv_rn(x, p.cell), cy = (int)__fdiv_rn(y, p.cell), cz = (int)__fdiv_rn(z, p.cell);
  return cx + (cy + cz * p.eg) * p.eg;
}
__device__ __forceinline__ int dir_cell(float x, float y, float z) {
  // coordinatesNormal = (n/2 + 1/2) / _nepsilon, _nepsilon = 1/7 + 0.00001   (normalset.h:96,108-112)
  const float ne = 1.0f / 7.0f + 0.00001f;
  const int cx = (int)__fdiv_rn(__fadd_rn(__fdiv_rn(x, 2.0f), 0.5f), ne);
  const int cy = (int)__fdiv_rn(__fadd_rn(__fdiv_rn(y, 2.0f), 0.5f), ne);
  const int cz = (int)__fdiv_rn(__fadd_rn(__fdiv_rn(z, 2.0f), 0.5f), ne);
  return cx + (cy + cz * NG) * NG;
}
__device__ __forceinline__ void normalize3(float& x, float& y, float& z) {
  const float s = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  const float l = __fsqrt_rn(s);
  x = __fdiv_rn(x, l); y = __fdiv_rn(y, l); z = __fdiv_rn(z, l);
}

// A side: bucket + direction cell of every pair of the first edge
__global__ void k2_join_keys(JoinParams p, uint32_t* __restrict__ bucket_of, uint32_t* __restrict__ key_of, uint32_t* __restrict__ counts) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n1) return;
  const int2 pr = p.A[k];
  const float4 a = p.Qn[pr.x], b = p.Qn[pr.y];
  float dx = __fsub_rn(b.x, a.x), dy = __fsub_rn(b.y, a.y), dz = __fsub_rn(b.z, a.z);
  const float ex = __fadd_rn(a.x, __fmul_rn(p.inv1, dx)), ey = __fadd_rn(a.y, __fmul_rn(p.inv1, dy)), ez = __fadd_rn(a.z, __fmul_rn(p.inv1, dz));
  normalize3(dx, dy, dz);
  const int pc = pos_cell(p, ex, ey, ez), dc = dir_cell(dx, dy, dz);
  const bool ok = pc >= 0 && dc >= 0 && dc < NG * NG * NG && ex >= 0.f && ey >= 0.f && ez >= 0.f && ex < 1.f && ey < 1.f && ez < 1.f;
  const uint32_t bkt = ok ? ((uint32_t)pc & (p.n_buckets - 1)) : 0xffffffffu;
  bucket_of[k] = bkt;
  key_of[k] = ((uint32_t)pc << 9) | (uint32_t)(dc & 511);     // eg <= 128 -> pc < 2^21
  if (ok) atomicAdd(counts + bkt, 1u);
}
__global__ void k2_join_scatter(long long n1, const uint32_t* __restrict__ bucket_of, uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n1 || bucket_of[k] == 0xffffffffu) return;
  sorted[atomicAdd(cursor + bucket_of[k], 1u)] = (uint32_t)k;
}

// B side: one thread per pair of the second edge; rasterise the cone of directions at angle alpha
// around the pair's direction (normalset.hpp:160-214) and collect the A pairs in the same position
// cell whose direction cell is coloured.
template <bool FILL>
__global__ void __launch_bounds__(128) k2_join_query(JoinParams p, const uint32_t* __restrict__ bucket_start, const uint32_t* __restrict__ sorted,
                                                     const uint32_t* __restrict__ key_of, uint32_t* __restrict__ cnt, int4* __restrict__ out, long long cap) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n2) return;
  const int2 pr = p.B[i];
  const float4 a = p.Qn[pr.x], b = p.Qn[pr.y];
  float dx = __fsub_rn(b.x, a.x), dy = __fsub_rn(b.y, a.y), dz = __fsub_rn(b.z, a.z);
  const float fx = __fadd_rn(a.x, __fmul_rn(p.inv2, dx)), fy = __fadd_rn(a.y, __fmul_rn(p.inv2, dy)), fz = __fadd_rn(a.z, __fmul_rn(p.inv2, dz));
  uint32_t n = 0;
  long long w = FILL ? (long long)cnt[i] : 0;
  if (fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < 1.f && fy < 1.f && fz < 1.f) {
    const int pc = pos_cell(p, fx, fy, fz);
    const uint32_t bkt = (uint32_t)pc & (p.n_buckets - 1);
    const uint32_t s = bucket_start[bkt], e = bucket_start[bkt + 1];
    if (e > s) {
      normalize3(dx, dy, dz);                      // queryn
      // world-space query point for the final check (super4pcs.cc:135-139,160-170)
      const float4 wa = p.Q[pr.x], wb = p.Q[pr.y];
      const float qx = __fadd_rn(wa.x, __fmul_rn(p.inv2, __fsub_rn(wb.x, wa.x)));
      const float qy = __fadd_rn(wa.y, __fmul_rn(p.inv2, __fsub_rn(wb.y, wa.y)));
      const float qz = __fadd_rn(wa.z, __fmul_rn(p.inv2, __fsub_rn(wb.z, wa.z)));
      // coloured direction cells: 343 bits
      uint32_t col[11];
#pragma unroll
      for (int t = 0; t < 11; ++t) col[t] = 0;
      const float alpha = acosf(p.cos_alpha);
      const float perimeter = 2.0f * 3.14159265358979323846f * atanf(alpha);
      const unsigned nb = 2u * (unsigned)ceilf(perimeter * (float)NG / 2.0f);
      const float step = 2.0f * 3.14159265358979323846f / (float)nb;
      const float sa = sinf(alpha);
      // q = FromTwoVectors((0,0,1), n):  c = n.z; axis = z x n = (-n.y, n.x, 0); s = sqrt(2(1+c)); vec = axis/s; w = s/2
      const float c = dz;
      float vx, vy, vz, qw;
      if (c < -1.0f + 1e-6f) { vx = 1.f; vy = 0.f; vz = 0.f; qw = 0.f; }   // antiparallel: rotation by pi about x (Eigen picks an SVD axis)
      else {
        const float sq = sqrtf((1.0f + c) * 2.0f), invs = 1.0f / sq;
        vx = -dy * invs; vy = dx * invs; vz = 0.f; qw = sq * 0.5f;
      }
      for (unsigned t = 0; t < nb; ++t) {
        const float th = (float)t * step;
        const float sx = sa * cosf(th), sy = sa * sinf(th), sz = p.cos_alpha;
        // q * v = v + w * uv + vec x uv, uv = 2 vec x v
        const float ux = 2.0f * (vy * sz - vz * sy), uy = 2.0f * (vz * sx - vx * sz), uz = 2.0f * (vx * sy - vy * sx);
        float rx = sx + qw * ux + (vy * uz - vz * uy);
        float ry = sy + qw * uy + (vz * ux - vx * uz);
        float rz = sz + qw * uz + (vx * uy - vy * ux);
        normalize3(rx, ry, rz);
        const int id = dir_cell(rx, ry, rz);
        if (id >= 0 && id < NG * NG * NG) col[id >> 5] |= 1u << (id & 31);
      }
      for (uint32_t t = s; t < e; ++t) {
        const uint32_t k = sorted[t];
        const uint32_t key = key_of[k];
        if ((int)(key >> 9) != pc) continue;                       // bucket collision
        const uint32_t dc = key & 511u;
        if (!((col[dc >> 5] >> (dc & 31)) & 1u)) continue;
        const int2 ap = p.A[k];
        const float4 pa = p.Q[ap.x], pb = p.Q[ap.y];
        // invPoint = pp1 + (pp2 - pp1) * invariant1 ; squaredNorm <= distance_threshold2 (sic: un-squared threshold)
        const float ix = __fadd_rn(pa.x, __fmul_rn(__fsub_rn(pb.x, pa.x), p.inv1));
        const float iy = __fadd_rn(pa.y, __fmul_rn(__fsub_rn(pb.y, pa.y), p.inv1));
        const float iz = __fadd_rn(pa.z, __fmul_rn(__fsub_rn(pb.z, pa.z), p.inv1));
        const float ddx = __fsub_rn(qx, ix), ddy = __fsub_rn(qy, iy), ddz = __fsub_rn(qz, iz);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)), __fmul_rn(ddz, ddz));
        if (!(d2 <= p.thr2)) continue;
        if (FILL) { if (w < cap) out[w] = make_int4(ap.x, ap.y, pr.x, pr.y); ++w; }
        ++n;
      }
    }
  }
  if (FILL && n > 1) {
    // the bucket order comes from atomics: order this query's quads by their first pair (unique key) so the output is deterministic
    const long long s0 = (long long)cnt[i], e0 = min(cap, s0 + (long long)n);
    for (long long a2 = s0 + 1; a2 < e0; ++a2) {
      const int4 v = out[a2];
      long long b2 = a2;
      while (b2 > s0 && (out[b2 - 1].x > v.x || (out[b2 - 1].x == v.x && out[b2 - 1].y > v.y))) { out[b2] = out[b2 - 1]; --b2; }
      out[b2] = v;
    }
  }
  if (!FILL) cnt[i] = n;
}

// ------------------------------------------------------------------------------- rigid transforms
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ bool normalize(V3& a) {
  const float s = dot(a, a);
  if (s == 0.f) return false;
  const float l = sqrtf(s);
  a.x /= l; a.y /= l; a.z /= l;
  return true;
}
__device__ __forceinline__ bool frame(V3 p0, V3 p1, V3 p2, V3 f[3]) {
  f[0] = sub(p1, p0);
  if (!normalize(f[0])) return false;
  const V3 e = sub(p2, p0);
  const float d = dot(e, f[0]);
  f[1] = {e.x - d * f[0].x, e.y - d * f[0].y, e.z - d * f[0].z};
  if (!normalize(f[1])) return false;
  f[2] = cross(f[0], f[1]);
  return true;
}

// One thread per quad.  Degenerate bases / quads -- for which the reference returns `true` with an
// UNINITIALISED matrix (match4pcsBase.cc:1533-1544) -- are rejected (ok = 0), as are non-orthogonal
// results (:1563, written R*R as in the reference).
__device__ bool rigid_from_quad(const float4* __restrict__ P, const float4* __restrict__ Q, const int* b, int4 quad, float* T) {
  const float4 b0 = P[b[0]], b1 = P[b[1]], b2 = P[b[2]];
  const float4 q0 = Q[quad.x], q1 = Q[quad.y], q2 = Q[quad.z];
  const V3 c1 = {(b0.x + b1.x + b2.x) / 3.0f, (b0.y + b1.y + b2.y) / 3.0f, (b0.z + b1.z + b2.z) / 3.0f};   // :1428
  const V3 c2 = {(q0.x + q1.x + q2.x) / 3.0f, (q0.y + q1.y + q2.y) / 3.0f, (q0.z + q1.z + q2.z) / 3.0f};   // :1452-1454
  V3 fp[3], fq[3];
  if (!frame({b0.x, b0.y, b0.z}, {b1.x, b1.y, b1.z}, {b2.x, b2.y, b2.z}, fp)) return false;
  if (!frame({q0.x, q0.y, q0.z}, {q1.x, q1.y, q1.z}, {q2.x, q2.y, q2.z}, fq)) return false;
  float R[3][3];
  const float fpv[3][3] = {{fp[0].x, fp[0].y, fp[0].z}, {fp[1].x, fp[1].y, fp[1].z}, {fp[2].x, fp[2].y, fp[2].z}};
  const float fqv[3][3] = {{fq[0].x, fq[0].y, fq[0].z}, {fq[1].x, fq[1].y, fq[1].z}, {fq[2].x, fq[2].y, fq[2].z}};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) R[i][j] = fpv[0][i] * fqv[0][j] + fpv[1][i] * fqv[1][j] + fpv[2][i] * fqv[2][j];   // rotate_p^T rotate_q :1560
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float s = R[i][0] * R[0][i] + R[i][1] * R[1][i] + R[i][2] * R[2][i];
    if (s - 1.0f > 1e-6f) return false;
    if (!isfinite(s)) return false;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    T[4 * i + 0] = R[i][0]; T[4 * i + 1] = R[i][1]; T[4 * i + 2] = R[i][2];
    const float cc = i == 0 ? c1.x : i == 1 ? c1.y : c1.z;
    T[4 * i + 3] = cc + (R[i][0] * (-c2.x) + R[i][1] * (-c2.y) + R[i][2] * (-c2.z));   // Tr(c1) R Tr(-c2) :1601-1610
  }
  return true;
}

__global__ void k2_rigid(const float4* __restrict__ P_unsorted, const float4* __restrict__ Q, const int* __restrict__ base4, const int4* __restrict__ quads,
                         long long n, float* __restrict__ T, uint8_t* __restrict__ ok) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int b[4] = {base4[0], base4[1], base4[2], base4[3]};
  float t[12];
  const bool good = rigid_from_quad(P_unsorted, Q, b, quads[i], t);
#pragma unroll
  for (int c = 0; c < 12; ++c) T[12 * i + c] = good ? t[c] : 0.f;
  ok[i] = good ? 1 : 0;
}

// ------------------------------------------------------------------------------- base selection
// One CTA per base.  Same procedure as SelectQuadrilateral in operMode 0, with a counter-based RNG
// instead of rand(): a random first point, the widest of `trials` random triangles whose two edges
// stay below max_base_diameter, the most coplanar fourth point that is not too close to the three,
// then the pairing with the smallest segment-to-segment distance and its two invariants (in double,
// as distSegmentToSegment is instantiated with Scalar = double, :428-435).
struct BaseOut { int id[4]; float inv1, inv2; int ok; float d1, d2, cos_alpha; };

__device__ double seg_seg(const double* p1, const double* p2, const double* q1, const double* q2, double& inv1, double& inv2) {
  const double kSmall = 0.0001;
  double u[3], v[3], w[3];
  for (int k = 0; k < 3; ++k) { u[k] = p2[k] - p1[k]; v[k] = q2[k] - q1[k]; w[k] = p1[k] - q1[k]; }
  const double a = u[0] * u[0] + u[1] * u[1] + u[2] * u[2], b = u[0] * v[0] + u[1] * v[1] + u[2] * v[2], c = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  const double d = u[0] * w[0] + u[1] * w[1] + u[2] * w[2], e = v[0] * w[0] + v[1] * w[1] + v[2] * w[2];
  const double f = a * c - b * b;
  double s1 = 0.0, s2 = f, t1 = 0.0, t2 = f;
  if (f < kSmall) { s1 = 0.0; s2 = 1.0; t1 = e; t2 = c; }
  else {
    s1 = (b * e - c * d); t1 = (a * e - b * d);
    if (s1 < 0.0) { s1 = 0.0; t1 = e; t2 = c; }
    else if (s1 > s2) { s1 = s2; t1 = e + b; t2 = c; }
  }
  if (t1 < 0.0) {
    t1 = 0.0;
    if (-d < 0.0) s1 = 0.0; else if (-d > a) s1 = s2; else { s1 = -d; s2 = a; }
  } else if (t1 > t2) {
    t1 = t2;
    if ((-d + b) < 0.0) s1 = 0; else if ((-d + b) > a) s1 = s2; else { s1 = (-d + b); s2 = a; }
  }
  inv1 = (fabs(s1) < kSmall ? 0.0 : s1 / s2);
  inv2 = (fabs(t1) < kSmall ? 0.0 : t1 / t2);
  double r[3];
  for (int k = 0; k < 3; ++k) r[k] = w[k] + inv1 * u[k] - inv2 * v[k];
  return sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
}

__global__ void __launch_bounds__(256) k2_select_bases(const float4* __restrict__ P, int n, float max_diam, int trials, uint64_t seed, BaseOut* __restrict__ out) {
  __shared__ float s_val[256];
  __shared__ int s_idx[256];
  __shared__ int s_tri[3];
  const int base = blockIdx.x, tid = threadIdx.x;
  BaseOut o{};
  for (int attempt = 0; attempt < 16; ++attempt) {
    const uint64_t s0 = mix64(seed ^ mix64(((uint64_t)base << 8) | (uint64_t)attempt));
    const int first = (int)(mix64(s0) % (uint64_t)n);
    const float4 p0 = P[first];
    const float sqmax = max_diam * max_diam;
    // widest admissible triangle (first maximum in trial order wins, like the serial loop's strict '>')
    float best = 0.f; int best_t = 0x7fffffff;
    for (int t = tid; t < trials; t += 256) {
      const uint64_t r = mix64(s0 + 2 * (uint64_t)t + 1), r2 = mix64(s0 + 2 * (uint64_t)t + 2);
      const float4 a = P[(int)(r % (uint64_t)n)], b = P[(int)(r2 % (uint64_t)n)];
      const V3 u = {a.x - p0.x, a.y - p0.y, a.z - p0.z}, w = {b.x - p0.x, b.y - p0.y, b.z - p0.z};
      const V3 cr = cross(u, w);
      const float wide = sqrtf(dot(cr, cr));
      if (wide > best && dot(u, u) < sqmax && dot(w, w) < sqmax) { best = wide; best_t = t; }
    }
    s_val[tid] = best; s_idx[tid] = best_t;
    __syncthreads();
    for (int o2 = 128; o2 > 0; o2 >>= 1) {
      if (tid < o2) {
        const float v = s_val[tid + o2]; const int ix = s_idx[tid + o2];
        if (v > s_val[tid] || (v == s_val[tid] && ix < s_idx[tid])) { s_val[tid] = v; s_idx[tid] = ix; }
      }
      __syncthreads();
    }
    const int bt = s_idx[0];
    const bool have_tri = s_val[0] > 0.f && bt != 0x7fffffff;
    __syncthreads();
    if (!have_tri) continue;
    const int i1 = first, i2 = (int)(mix64(s0 + 2 * (uint64_t)bt + 1) % (uint64_t)n), i3 = (int)(mix64(s0 + 2 * (uint64_t)bt + 2) % (uint64_t)n);
    const float4 p1 = P[i1], p2 = P[i2], p3 = P[i3];
    // plane through the three points: A x + B y + C z = 1 (:527-546, evaluated in double, stored in float)
    const double x1 = p1.x, y1 = p1.y, z1 = p1.z, x2 = p2.x, y2 = p2.y, z2 = p2.z, x3 = p3.x, y3 = p3.y, z3 = p3.z;
    const float denom = (float)(-x3 * y2 * z1 + x2 * y3 * z1 + x3 * y1 * z2 - x1 * y3 * z2 - x2 * y1 * z3 + x1 * y2 * z3);
    if (denom == 0.f) continue;
    const float A = (float)((-y2 * z1 + y3 * z1 + y1 * z2 - y3 * z2 - y1 * z3 + y2 * z3) / denom);
    const float B = (float)((x2 * z1 - x3 * z1 - x1 * z2 + x3 * z2 + x1 * z3 - x2 * z3) / denom);
    const float C = (float)((-x2 * y1 + x3 * y1 + x1 * y2 - x3 * y2 - x1 * y3 + x2 * y3) / denom);
    const float too_small = (max_diam * 0.1f) * (max_diam * 0.1f);
    float bd = 3.4e38f; int bi = 0x7fffffff;
    for (int i = tid; i < n; i += 256) {
      const float4 q = P[i];
      const V3 qv = {q.x, q.y, q.z};
      const V3 d1 = sub(qv, {p1.x, p1.y, p1.z}), d2 = sub(qv, {p2.x, p2.y, p2.z}), d3 = sub(qv, {p3.x, p3.y, p3.z});
      if (dot(d1, d1) >= too_small && dot(d2, d2) >= too_small && dot(d3, d3) >= too_small) {
        const float dist = fabsf((float)((double)(A * q.x + B * q.y + C * q.z) - 1.0));
        if (dist < bd) { bd = dist; bi = i; }
      }
    }
    s_val[tid] = bd; s_idx[tid] = bi;
    __syncthreads();
    for (int o2 = 128; o2 > 0; o2 >>= 1) {
      if (tid < o2) {
        const float v = s_val[tid + o2]; const int ix = s_idx[tid + o2];
        if (v < s_val[tid] || (v == s_val[tid] && ix < s_idx[tid])) { s_val[tid] = v; s_idx[tid] = ix; }
      }
      __syncthreads();
    }
    const int i4 = s_idx[0];
    __syncthreads();
    if (i4 == 0x7fffffff) continue;
    if (tid == 0) {
      // TryQuadrilateral: all ordered (i,j) with the remaining two in ascending order (:419-446)
      const int ids[4] = {i1, i2, i3, i4};
      double pt[4][3];
      for (int k = 0; k < 4; ++k) { const float4 q = P[ids[k]]; pt[k][0] = q.x; pt[k][1] = q.y; pt[k][2] = q.z; }
      float min_d = 3.4e38f; int bb[4] = {-1, -1, -1, -1}; float inv1 = 0.f, inv2 = 0.f;
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
          if (i == j) continue;
          int k = 0; while (k == i || k == j) k++;
          int l = 0; while (l == i || l == j || l == k) l++;
          double li1, li2;
          const float sd = (float)seg_seg(pt[i], pt[j], pt[k], pt[l], li1, li2);
          if (sd < min_d) { min_d = sd; bb[0] = i; bb[1] = j; bb[2] = k; bb[3] = l; inv1 = (float)li1; inv2 = (float)li2; }
        }
      if (bb[0] >= 0) {
        for (int k = 0; k < 4; ++k) o.id[k] = ids[bb[k]];
        o.inv1 = inv1; o.inv2 = inv2; o.ok = 1;
        const float4 a0 = P[o.id[0]], a1 = P[o.id[1]], a2 = P[o.id[2]], a3 = P[o.id[3]];
        V3 e1 = {a1.x - a0.x, a1.y - a0.y, a1.z - a0.z}, e2 = {a3.x - a2.x, a3.y - a2.y, a3.z - a2.z};
        o.d1 = sqrtf(dot(e1, e1)); o.d2 = sqrtf(dot(e2, e2));     // distance1 / distance6 (:1951-1952)
        normalize(e1); normalize(e2);
        o.cos_alpha = dot(e1, e2);                                // super4pcs.cc:109-111
      }
      s_tri[0] = o.ok;
    }
    __syncthreads();
    if (s_tri[0]) break;
  }
  if (tid == 0) out[base] = o;
}

// keep[i] = 1 for the `keep_n` quads of a base with the smallest hash (a deterministic random subset;
// the reference draws rand() % size until it has 100 distinct indices, :1866-1869)
__global__ void k2_mark_subset(long long n, uint64_t seed, unsigned long long thr, uint8_t* __restrict__ keep) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = (mix64(seed ^ (uint64_t)i) >> 11) <= thr ? 1 : 0;
}

struct Scratch {
  DevBuf cnt, pairs1, pairs2, quads, bucket_of, key_of, bucket_start, sorted, T, ok, base, qn;
};
Scratch g_scratch[16];   // per device

int scan_u32(pgp_ctx* ctx, uint32_t* data, int64_t n, uint64_t* total) {
  PGP_CUDA(ctx, ctx->scene.scratch.reserve((size_t)((n + 1) / 2048 + 4096) * 4));
  int rc = pgp_scan_exclusive_u32(ctx, data, n + 1, ctx->scene.scratch.as<uint32_t>());
  if (rc) return rc;
  uint32_t t = 0;
  PGP_CUDA(ctx, cudaMemcpyAsync(&t, data + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
  PGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *total = t;
  return PGP_OK;
}

// all ordered pairs into `out` (device, grown as needed); returns the number of ORDERED pairs
int extract_pairs_dev(pgp_ctx* ctx, const Model& m, float dist, float eps, DevBuf& out, int64_t* n_pairs) {
  Scratch& sc = g_scratch[ctx->device & 15];
  const int nq = m.nq, B = (nq + 255) / 256;
  PGP_CUDA(ctx, sc.cnt.reserve((size_t)(nq + 1) * 4));
  uint32_t* cnt = sc.cnt.as<uint32_t>();
  PGP_CUDA(ctx, cudaMemsetAsync(cnt, 0, (size_t)(nq + 1) * 4, ctx->stream));
  k2_pairs<false><<<B, 256, 0, ctx->stream>>>(m.search.as<float4>(), nq, dist, eps, cnt, nullptr, 0);
  ctx->launches++;
  uint64_t total = 0;
  int rc = scan_u32(ctx, cnt, nq, &total);
  if (rc) return rc;
  *n_pairs = (int64_t)total * 2;
  PGP_CUDA(ctx, out.reserve((size_t)std::max<int64_t>(*n_pairs, 1) * 8));
  if (total) {
    k2_pairs<true><<<B, 256, 0, ctx->stream>>>(m.search.as<float4>(), nq, dist, eps, cnt, out.as<int2>(), *n_pairs);
    ctx->launches++;
  }
  PGP_CUDA(ctx, cudaGetLastError());
  return PGP_OK;
}

int find_quads_dev(pgp_ctx* ctx, const Model& m, float cos_alpha, float inv1, float inv2, float eps, const int2* A, int64_t n1, const int2* B,
                   int64_t n2, DevBuf& out, int64_t* n_quads) {
  Scratch& sc = g_scratch[ctx->device & 15];
  *n_quads = 0;
  if (n1 <= 0 || n2 <= 0) return PGP_OK;
  JoinParams p{};
  p.Qn = m.search_unit.as<float4>(); p.Q = m.search.as<float4>();
  p.A = A; p.n1 = n1; p.B = B; p.n2 = n2;
  p.inv1 = inv1; p.inv2 = inv2; p.cos_alpha = cos_alpha; p.thr2 = eps;
  // IndexedNormalSet(eps_n): gridDepth = int(-log2(eps_n)); egSize = 2^gridDepth; _epsilon = 1/egSize  (normalset.h:117-123)
  const float eps_n = eps / m.unit_ratio;
  int depth = (int)(-std::log2(eps_n));
  if (depth < 0) depth = 0;
  if (depth > 7) return pgp_fail(ctx, PGP_E_TOO_LARGE, "quad join: model diameter / delta too large (grid depth %d > 7)", depth);
  p.eg = 1 << depth;
  p.cell = 1.0f / (float)p.eg;
  p.n_buckets = 1u << std::min(3 * depth, 20);
  PGP_CUDA(ctx, sc.bucket_of.reserve((size_t)n1 * 4));
  PGP_CUDA(ctx, sc.key_of.reserve((size_t)n1 * 4));
  PGP_CUDA(ctx, sc.sorted.reserve((size_t)n1 * 4));
  PGP_CUDA(ctx, sc.bucket_start.reserve(((size_t)p.n_buckets + 1) * 8));
  uint32_t* bs = sc.bucket_start.as<uint32_t>();
  uint32_t* cursor = bs + p.n_buckets + 1;
  PGP_CUDA(ctx, cudaMemsetAsync(bs, 0, ((size_t)p.n_buckets + 1) * 4, ctx->stream));
  k2_join_keys<<<(unsigned)((n1 + 255) / 256), 256, 0, ctx->stream>>>(p, sc.bucket_of.as<uint32_t>(), sc.key_of.as<uint32_t>(), bs);
  ctx->launches++;
  uint64_t tot = 0;
  int rc = scan_u32(ctx, bs, p.n_buckets, &tot);
  if (rc) return rc;
  PGP_CUDA(ctx, cudaMemcpyAsync(cursor, bs, (size_t)p.n_buckets * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  k2_join_scatter<<<(unsigned)((n1 + 255) / 256), 256, 0, ctx->stream>>>(n1, sc.bucket_of.as<uint32_t>(), cursor, sc.sorted.as<uint32_t>());
  ctx->launches++;
  PGP_CUDA(ctx, sc.cnt.reserve((size_t)(n2 + 1) * 4));
  uint32_t* cnt = sc.cnt.as<uint32_t>();
  PGP_CUDA(ctx, cudaMemsetAsync(cnt, 0, (size_t)(n2 + 1) * 4, ctx->stream));
  k2_join_query<false><<<(unsigned)((n2 + 127) / 128), 128, 0, ctx->stream>>>(p, bs, sc.sorted.as<uint32_t>(), sc.key_of.as<uint32_t>(), cnt, nullptr, 0);
  ctx->launches++;
  uint64_t nquads = 0;
  rc = scan_u32(ctx, cnt, n2, &nquads);
  if (rc) return rc;
  *n_quads = (int64_t)nquads;
  PGP_CUDA(ctx, out.reserve((size_t)std::max<uint64_t>(nquads, 1) * 16));
  if (nquads) {
    k2_join_query<true><<<(unsigned)((n2 + 127) / 128), 128, 0, ctx->stream>>>(p, bs, sc.sorted.as<uint32_t>(), sc.key_of.as<uint32_t>(), cnt, out.as<int4>(),
                                                                              (long long)nquads);
    ctx->launches++;
  }
  PGP_CUDA(ctx, cudaGetLastError());
  return PGP_OK;
}

}  // namespace

int k2_extract_pairs(pgp_ctx* ctx, const Model& m, float dist, float eps, int32_t* pairs_host, int64_t cap, int64_t* n_pairs) {
  Scratch& sc = g_scratch[ctx->device & 15];
  int rc = extract_pairs_dev(ctx, m, dist, eps, sc.pairs1, n_pairs);
  if (rc) return rc;
  const int64_t n = std::min(cap, *n_pairs);
  if (pairs_host && n > 0) {
    PGP_CUDA(ctx, cudaMemcpyAsync(pairs_host, sc.pairs1.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return PGP_OK;
}

int k2_find_quads(pgp_ctx* ctx, const Model& m, const int32_t* base4, float inv1, float inv2, float eps, const int32_t* p1, int64_t n1,
                  const int32_t* p2, int64_t n2, int32_t* quads_host, int64_t cap, int64_t* n_quads) {
  Scratch& sc = g_scratch[ctx->device & 15];
  const Scene& s = ctx->scene;
  for (int k = 0; k < 4; ++k)
    if (base4[k] < 0 || base4[k] >= s.n) return pgp_fail(ctx, PGP_E_INVALID, "base id out of range");
  // angle between the base's two edges (super4pcs.cc:109-111) from the centred scene points
  float4 bp[4];
  for (int k = 0; k < 4; ++k) PGP_CUDA(ctx, cudaMemcpyAsync(&bp[k], s.unsorted.as<float4>() + base4[k], 16, cudaMemcpyDeviceToHost, ctx->stream));
  PGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  auto unit = [](float4 a, float4 b, float o[3]) {
    o[0] = b.x - a.x; o[1] = b.y - a.y; o[2] = b.z - a.z;
    volatile float s2 = o[0] * o[0]; s2 = s2 + o[1] * o[1]; s2 = s2 + o[2] * o[2];
    float l = sqrtf(s2);
    o[0] /= l; o[1] /= l; o[2] /= l;
  };
  float u[3], v[3];
  unit(bp[0], bp[1], u); unit(bp[2], bp[3], v);
  volatile float ca = u[0] * v[0]; ca = ca + u[1] * v[1]; ca = ca + u[2] * v[2];
  *n_quads = 0;
  if (n1 <= 0 || n2 <= 0) return PGP_OK;
  PGP_CUDA(ctx, sc.pairs1.reserve((size_t)n1 * 8));
  PGP_CUDA(ctx, sc.pairs2.reserve((size_t)n2 * 8));
  PGP_CUDA(ctx, cudaMemcpyAsync(sc.pairs1.p, p1, (size_t)n1 * 8, cudaMemcpyHostToDevice, ctx->stream));
  PGP_CUDA(ctx, cudaMemcpyAsync(sc.pairs2.p, p2, (size_t)n2 * 8, cudaMemcpyHostToDevice, ctx->stream));
  int rc = find_quads_dev(ctx, m, ca, inv1, inv2, eps, sc.pairs1.as<int2>(), n1, sc.pairs2.as<int2>(), n2, sc.quads, n_quads);
  if (rc) return rc;
  const int64_t n = std::min(cap, *n_quads);
  if (quads_host && n > 0) {
    PGP_CUDA(ctx, cudaMemcpyAsync(quads_host, sc.quads.p, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    PGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return PGP_OK;
}

int k2_rigid_from_quads(pgp_ctx* ctx, const Model& m, const int32_t* base4, const int32_t* quads_host, int64_t n, float* T_host, uint8_t* ok_host) {
  Scratch& sc = g_scratch[ctx->device & 15];
  if (n == 0) return PGP_OK;
  for (int k = 0; k < 4; ++k)
    if (base4[k] < 0 || base4[k] >= ctx->scene.n) return pgp_fail(ctx, PGP_E_INVALID, "base id out of range");
  PGP_CUDA(ctx, sc.quads.reserve((size_t)n * 16));
  PGP_CUDA(ctx, sc.T.reserve((size_t)n * 48));
  PGP_CUDA(ctx, sc.ok.reserve((size_t)n));
  PGP_CUDA(ctx, sc.base.reserve(64));
  PGP_CUDA(ctx, cudaMemcpyAsync(sc.quads.p, quads_host, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
  PGP_CUDA(ctx, cudaMemcpyAsync(sc.base.p, base4, 16, cudaMemcpyHostToDevice, ctx->stream));
  k2_rigid<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->scene.unsorted.as<float4>(), m.search.as<float4>(), sc.base.as<int>(), sc.quads.as<int4>(),
                                                                n, sc.T.as<float>(), sc.ok.as<uint8_t>());
  ctx->launches++;
  PGP_CUDA(ctx, cudaMemcpyAsync(T_host, sc.T.p, (size_t)n * 48, cudaMemcpyDeviceToHost, ctx->stream));
  PGP_CUDA(ctx, cudaMemcpyAsync(ok_host, sc.ok.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  PGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  PGP_CUDA(ctx, cudaGetLastError());
  return PGP_OK;
}

// order-preserving compaction of the accepted transforms of one base behind the ones already generated
__global__ void k2_flags(const uint8_t* __restrict__ ok, const uint8_t* __restrict__ keep, long long n, uint32_t* __restrict__ flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (ok[i] && (!keep || keep[i])) ? 1u : 0u;
}
__global__ void k2_append(const float* __restrict__ T, const uint8_t* __restrict__ ok, const uint8_t* __restrict__ keep, const uint32_t* __restrict__ off,
                          long long n, float* __restrict__ dst, long long room) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !ok[i] || (keep && !keep[i]) || (long long)off[i] >= room) return;
#pragma unroll
  for (int c = 0; c < 12; ++c) dst[12 * (long long)off[i] + c] = T[12 * i + c];
}

int k2_generate(pgp_ctx* ctx, Model& m, const pgp_pcs_opts* o, uint64_t seed, int64_t max_hyp, int64_t* n_hyp) {
  Scratch& sc = g_scratch[ctx->device & 15];
  const Scene& s = ctx->scene;
  *n_hyp = 0;
  m.n_gen = 0;
  const int nb = std::max(1, o->n_bases);
  float max_diam = o->max_base_diameter;
  if (!(max_diam > 0.f)) max_diam = m.search_diameter;   // P_diameter_ estimate of init() (:274-283): here the exact bbox diagonal bound
  PGP_CUDA(ctx, sc.base.reserve((size_t)nb * sizeof(BaseOut) + 64));
  PGP_CUDA(ctx, m.gen_T.reserve((size_t)max_hyp * 48));
  int64_t cur = 0;
  BaseOut* d_bases = reinterpret_cast<BaseOut*>(sc.base.as<char>() + 64);
  k2_select_bases<<<nb, 256, 0, ctx->stream>>>(s.unsorted.as<float4>(), s.n, max_diam, std::max(1, o->base_trials), seed, d_bases);
  ctx->launches++;
  std::vector<BaseOut> bases(nb);
  PGP_CUDA(ctx, cudaMemcpyAsync(bases.data(), d_bases, (size_t)nb * sizeof(BaseOut), cudaMemcpyDeviceToHost, ctx->stream));
  PGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  PGP_CUDA(ctx, cudaGetLastError());
  const float eps = s.delta;          // distance_factor * options_.delta, distance_factor = 1 (match4pcsBase.h:99)
  for (int b = 0; b < nb; ++b) {
    const BaseOut& bo = bases[b];
    if (!bo.ok) continue;
    int64_t n1 = 0, n2 = 0, nq = 0;
    int rc = extract_pairs_dev(ctx, m, bo.d1, eps, sc.pairs1, &n1);
    if (rc) return rc;
    if (n1 == 0) continue;
    rc = extract_pairs_dev(ctx, m, bo.d2, eps, sc.pairs2, &n2);
    if (rc) return rc;
    if (n2 == 0) continue;
    rc = find_quads_dev(ctx, m, bo.cos_alpha, bo.inv1, bo.inv2, eps, sc.pairs1.as<int2>(), n1, sc.pairs2.as<int2>(), n2, sc.quads, &nq);
    if (rc) return rc;
    if (nq == 0) continue;
    PGP_CUDA(ctx, sc.T.reserve((size_t)nq * 48));
    PGP_CUDA(ctx, sc.ok.reserve((size_t)nq * 2));
    PGP_CUDA(ctx, cudaMemcpyAsync(sc.base.p, bo.id, 16, cudaMemcpyHostToDevice, ctx->stream));
    k2_rigid<<<(unsigned)((nq + 127) / 128), 128, 0, ctx->stream>>>(s.unsorted.as<float4>(), m.search.as<float4>(), sc.base.as<int>(), sc.quads.as<int4>(), nq,
                                                                   sc.T.as<float>(), sc.ok.as<uint8_t>());
    ctx->launches++;
    uint8_t* keep = nullptr;
    if (o->max_quads_per_base > 0 && nq > o->max_quads_per_base) {
      keep = sc.ok.as<uint8_t>() + nq;
      // oversample by 25 % and cut at exactly max_quads_per_base through the scan offsets below
      const double frac = std::min(1.0, 1.25 * (double)o->max_quads_per_base / (double)nq);
      const unsigned long long thr = (unsigned long long)(frac * 9007199254740992.0);   // 2^53
      k2_mark_subset<<<(unsigned)((nq + 255) / 256), 256, 0, ctx->stream>>>(nq, mix64(seed ^ (0xABCDull + (uint64_t)b)), thr, keep);
      ctx->launches++;
    }
    PGP_CUDA(ctx, sc.cnt.reserve((size_t)(nq + 1) * 4));
    uint32_t* flag = sc.cnt.as<uint32_t>();
    PGP_CUDA(ctx, cudaMemsetAsync(flag + nq, 0, 4, ctx->stream));
    k2_flags<<<(unsigned)((nq + 255) / 256), 256, 0, ctx->stream>>>(sc.ok.as<uint8_t>(), keep, nq, flag);
    ctx->launches++;
    uint64_t added = 0;
    rc = scan_u32(ctx, flag, nq, &added);
    if (rc) return rc;
    int64_t room = max_hyp - cur;
    if (o->max_quads_per_base > 0) room = std::min<int64_t>(room, o->max_quads_per_base);
    k2_append<<<(unsigned)((nq + 255) / 256), 256, 0, ctx->stream>>>(sc.T.as<float>(), sc.ok.as<uint8_t>(), keep, flag, nq,
                                                                    m.gen_T.as<float>() + 12 * cur, room);
    ctx->launches++;
    cur += std::min<int64_t>(room, (int64_t)added);
    if (cur >= max_hyp) break;
  }
  PGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  PGP_CUDA(ctx, cudaGetLastError());
  m.n_gen = std::min<int64_t>((int64_t)cur, max_hyp);
  *n_hyp = m.n_gen;
  return PGP_OK;
}
