// K1 -- scene voxel grid build on the device.
//
// Replaces Match4PCSBase::initKdTree (S4/algorithms/match4pcsBase.cc:1046-1056) and
// KdTree::finalize/createTree/split (S4/accelerators/kdtree.h:355-370,560-641,522-538): the
// kd-tree is only ever asked "closest scene point within delta" (kdtree.h:394-459), which a
// direct-addressed grid with cell edge >= delta answers from the 27 cells around the query.
//
// Layout in HBM (all L2-resident at the benchmark sizes):
//   pts        n x float4, sorted by cell (x fastest, then y, then z); w = original index
//   aux        n x float4, same order: unit normal, prior
//   cell_start (n_cells + 1) x u32 prefix sums -> the 3 x-neighbour cells of a row are ONE
//              contiguous range, so a 27-cell probe is 9 ranges
//   bitmap     1 bit per cell: "some scene point in the 27 cells around this cell" (the cull)
#include <math.h>
#include <string.h>

#include "pgp_internal.cuh"

namespace {

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
inline float ord2f(int i) { i ^= ((i >> 31) & 0x7fffffff); float f; memcpy(&f, &i, 4); return f; }

// centre the raw cloud on the (host-computed, sequential-fp32) centroid and reduce the AABB.
// sampled_P_3D_[i].pos() -= centroid_P_   match4pcsBase.cc:253-255
__global__ void k1_centre_bounds(const float* __restrict__ raw, int n, float cx, float cy, float cz,
                                 float4* __restrict__ out, int* __restrict__ bounds) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, mx[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  if (i < n) {
    float x = __fsub_rn(raw[3 * i], cx), y = __fsub_rn(raw[3 * i + 1], cy), z = __fsub_rn(raw[3 * i + 2], cz);
    out[i] = make_float4(x, y, z, __int_as_float(i));
    mn[0] = mx[0] = f2ord(x); mn[1] = mx[1] = f2ord(y); mn[2] = mx[2] = f2ord(z);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    mn[k] = __reduce_min_sync(0xffffffffu, mn[k]);
    mx[k] = __reduce_max_sync(0xffffffffu, mx[k]);
  }
  if ((threadIdx.x & 31) == 0 && mn[0] != 0x7fffffff) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { atomicMin(bounds + k, mn[k]); atomicMax(bounds + 3 + k, mx[k]); }
  }
}

__device__ __forceinline__ uint32_t cell_index(const GridParams& g, float x, float y, float z) {
  int cx = (int)cell_coord(x, g.lo[0], g.inv_h);
  int cy = (int)cell_coord(y, g.lo[1], g.inv_h);
  int cz = (int)cell_coord(z, g.lo[2], g.inv_h);
  cx = min(max(cx, 2), g.dim[0] - 3); cy = min(max(cy, 2), g.dim[1] - 3); cz = min(max(cz, 2), g.dim[2] - 3);
  return (uint32_t)((cz * g.dim[1] + cy) * g.dim[0] + cx);
}

__global__ void k1_count(const float4* __restrict__ pts, int n, GridParams g, uint32_t* __restrict__ cell_of,
                         uint32_t* __restrict__ counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  uint32_t c = cell_index(g, p.x, p.y, p.z);
  cell_of[i] = c;
  atomicAdd(counts + c, 1u);
}

// ---- exclusive scan of u32 (hand-written, 3 levels cover 2^29 cells) -------------------------
constexpr int SCAN_T = 256, SCAN_ITEMS = 8, SCAN_BLOCK = SCAN_T * SCAN_ITEMS;

__global__ void k1_scan_block(uint32_t* __restrict__ data, int64_t n, uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t warp_tot[SCAN_T / 32];
  int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = (base + k < n) ? data[base + k] : 0u; sum += v[k]; }
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
  __syncthreads();
  uint32_t woff = 0, total = 0;
#pragma unroll
  for (int w = 0; w < SCAN_T / 32; ++w) { uint32_t t = warp_tot[w]; if (w < (threadIdx.x >> 5)) woff += t; total += t; }
  uint32_t run = woff + incl - sum;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { if (base + k < n) data[base + k] = run; run += v[k]; }
  if (threadIdx.x == 0 && block_sums) block_sums[blockIdx.x] = total;
}
__global__ void k1_scan_add(uint32_t* __restrict__ data, int64_t n, const uint32_t* __restrict__ block_offs) {
  int64_t i = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  uint32_t off = block_offs[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k, i += SCAN_T) if (i < n) data[i] += off;
}

}  // namespace

// in-place exclusive scan; scratch must hold ceil(n/2048) + ceil(n/2048^2) + 8 words
int pgp_scan_exclusive_u32(pgp_ctx* ctx, uint32_t* data, int64_t n, uint32_t* scratch) {
  int64_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  k1_scan_block<<<(unsigned)nb, SCAN_T, 0, ctx->stream>>>(data, n, nb > 1 ? scratch : nullptr);
  ctx->launches++;
  if (nb > 1) {
    int rc = pgp_scan_exclusive_u32(ctx, scratch, nb, scratch + nb);
    if (rc) return rc;
    k1_scan_add<<<(unsigned)nb, SCAN_T, 0, ctx->stream>>>(data, n, scratch);
    ctx->launches++;
  }
  PGP_CUDA(ctx, cudaGetLastError());
  return PGP_OK;
}

namespace {

__global__ void k1_scatter(const float4* __restrict__ in, int n, const uint32_t* __restrict__ cell_of,
                           uint32_t* __restrict__ cursor, float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t pos = atomicAdd(cursor + cell_of[i], 1u);
  out[pos] = in[i];
}

// Deterministic order inside a cell (ascending original index): the scatter above is atomic-ordered.
__global__ void k1_sort_cells(float4* __restrict__ pts, const uint32_t* __restrict__ cell_start, int64_t n_cells,
                              unsigned long long* __restrict__ n_occupied) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool occ = false;
  if (c < n_cells) {
    uint32_t s = cell_start[c], e = cell_start[c + 1];
    occ = e > s;
    for (uint32_t i = s + 1; i < e; ++i) {
      float4 v = pts[i];
      int key = __float_as_int(v.w);
      uint32_t j = i;
      while (j > s && __float_as_int(pts[j - 1].w) > key) { pts[j] = pts[j - 1]; --j; }
      pts[j] = v;
    }
  }
  unsigned b = __ballot_sync(0xffffffffu, occ);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_occupied, (unsigned long long)__popc(b));
}

// 1 bit per cell: any scene point in the 27 cells around it.  One warp writes one 32-bit word.
__global__ void k1_dilate(const uint32_t* __restrict__ cell_start, GridParams g, uint32_t* __restrict__ bitmap, int64_t n_words) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool any = false;
  if (c < g.n_cells) {
    int cx = (int)(c % g.dim[0]);
    int cy = (int)((c / g.dim[0]) % g.dim[1]);
    int cz = (int)(c / ((int64_t)g.dim[0] * g.dim[1]));
    if (cx >= 1 && cx < g.dim[0] - 1 && cy >= 1 && cy < g.dim[1] - 1 && cz >= 1 && cz < g.dim[2] - 1) {
#pragma unroll
      for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          int64_t row = ((int64_t)(cz + dz) * g.dim[1] + (cy + dy)) * g.dim[0] + cx;
          any |= cell_start[row + 2] > cell_start[row - 1];
        }
    }
  }
  unsigned b = __ballot_sync(0xffffffffu, any);
  if ((threadIdx.x & 31) == 0 && (c >> 5) < n_words) bitmap[c >> 5] = b;
}

// aux[pos] = (unit normal, prior) of the point sorted to pos.  Normalisation as
// Point3D::set_normal (S4/shared4pcs.h:85-87: n / sqrt(n.n)) with tiny normals zeroed
// (S4/utils/geometry.h:56-82).
__global__ void k1_build_aux(const float4* __restrict__ pts, int n, const float* __restrict__ nrm_raw,
                             const float* __restrict__ prior, float4* __restrict__ aux, float4* __restrict__ aux_orig) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int o = __float_as_int(pts[i].w);
  float nx = 0.f, ny = 0.f, nz = 0.f;
  if (nrm_raw) {
    float x = nrm_raw[3 * o], y = nrm_raw[3 * o + 1], z = nrm_raw[3 * o + 2];
    float s = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    if (!(s < 0.01f)) {
      float len = __fsqrt_rn(s);
      nx = __fdiv_rn(x, len); ny = __fdiv_rn(y, len); nz = __fdiv_rn(z, len);
    }
  }
  const float4 a = make_float4(nx, ny, nz, prior[o]);
  aux[i] = a;
  aux_orig[o] = a;
}

// prior of scene point i = img[row][col]/10000 at the pin-hole projection of the UN-centred point
// (match4pcsBase.cc:317-340).  b_ii.pos() += centroid_P_ re-adds the centroid in fp32 (so the
// projected point is fl(fl(raw - c) + c), not raw); camIntrinsic * Vector3f evaluates
// (K_r0 x + K_r1 y) + K_r2 z; int col = u/w truncates toward zero.
__global__ void k1_prior_project(const float* __restrict__ raw, int n, float cx, float cy, float cz,
                                 const uint16_t* __restrict__ img, int rows, int cols,
                                 float k0, float k1, float k2, float k3, float k4, float k5, float k6, float k7, float k8,
                                 float* __restrict__ prior, int* __restrict__ non_binary) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = __fadd_rn(__fsub_rn(raw[3 * i], cx), cx);
  float y = __fadd_rn(__fsub_rn(raw[3 * i + 1], cy), cy);
  float z = __fadd_rn(__fsub_rn(raw[3 * i + 2], cz), cz);
  float u = __fadd_rn(__fadd_rn(__fmul_rn(k0, x), __fmul_rn(k1, y)), __fmul_rn(k2, z));
  float v = __fadd_rn(__fadd_rn(__fmul_rn(k3, x), __fmul_rn(k4, y)), __fmul_rn(k5, z));
  float w = __fadd_rn(__fadd_rn(__fmul_rn(k6, x), __fmul_rn(k7, y)), __fmul_rn(k8, z));
  float fc = __fdiv_rn(u, w), fr = __fdiv_rn(v, w);
  // float -> int conversion of an out-of-range / NaN value is undefined in C++ (x86 gives INT_MIN);
  // both end up clamped into the image here.
  int col = (fc >= 0.f && fc < (float)cols) ? (int)fc : (fc >= (float)cols ? cols - 1 : 0);
  int row = (fr >= 0.f && fr < (float)rows) ? (int)fr : (fr >= (float)rows ? rows - 1 : 0);
  float p = __fdiv_rn((float)img[(size_t)row * cols + col], 10000.f);
  prior[i] = p;
  if (p != 0.f && p != 1.f) atomicOr(non_binary, 1);
}

__global__ void k1_check_binary(const float* __restrict__ prior, int n, int* __restrict__ non_binary) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { float p = prior[i]; if (p != 0.f && p != 1.f) atomicOr(non_binary, 1); }
}

__global__ void k1_fill(float* __restrict__ p, int n, float v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

// Expects scene.{n, delta, cP, has_nrm} set and xyz_raw / nrm_raw / prior uploaded.
// Leaves pts / aux / cell_start / bitmap built and scene.g filled.
int k1_build_grid(pgp_ctx* ctx) {
  Scene& s = ctx->scene;
  const int n = s.n;
  const int T = 256, B = (n + T - 1) / T;
  cudaStream_t st = ctx->stream;

  PGP_CUDA(ctx, s.pts.reserve((size_t)n * 16));
  PGP_CUDA(ctx, s.aux.reserve((size_t)n * 16));
  PGP_CUDA(ctx, s.aux_orig.reserve((size_t)n * 16));
  PGP_CUDA(ctx, s.unsorted.reserve((size_t)n * 16));
  PGP_CUDA(ctx, s.cell_of.reserve((size_t)n * 4));
  PGP_CUDA(ctx, ctx->work.reserve(4096));

  // 1. centre + AABB
  int* bounds = ctx->work.as<int>();
  int init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
  PGP_CUDA(ctx, cudaMemcpyAsync(bounds, init, sizeof(init), cudaMemcpyHostToDevice, st));
  k1_centre_bounds<<<B, T, 0, st>>>(s.xyz_raw.as<float>(), n, s.cP[0], s.cP[1], s.cP[2], s.unsorted.as<float4>(), bounds);
  ctx->launches++;
  int hb[6];
  PGP_CUDA(ctx, cudaMemcpyAsync(hb, bounds, sizeof(hb), cudaMemcpyDeviceToHost, st));
  PGP_CUDA(ctx, cudaStreamSynchronize(st));
  float mn[3], mx[3];
  for (int k = 0; k < 3; ++k) { mn[k] = ord2f(hb[k]); mx[k] = ord2f(hb[3 + k]); }
  for (int k = 0; k < 3; ++k)
    if (!(mn[k] <= mx[k]) || !std::isfinite(mn[k]) || !std::isfinite(mx[k]))
      return pgp_fail(ctx, PGP_E_INVALID, "scene cloud has non-finite coordinates");

  // 2. grid parameters.  h = delta (1 + 2^-8): see DESIGN.md "27-cell completeness".
  GridParams& g = s.g;
  memset(&g, 0, sizeof(g));
  const float df = s.delta;                    // const Scalar epsilon = options_.delta (double -> float)
  g.r2 = df * df;                              // sq_eps, match4pcsBase.cc:1710
  g.h = df * (1.0f + 1.0f / 256.0f);
  g.inv_h = 1.0f / g.h;
  double cells = 1;
  for (int k = 0; k < 3; ++k) {
    g.lo[k] = mn[k] - 2.5f * g.h;              // points start in cell 2: two empty apron cells
    double ext = ((double)mx[k] - (double)g.lo[k]) / (double)g.h;
    if (ext > 4000.0) return pgp_fail(ctx, PGP_E_TOO_LARGE, "scene extent / delta = %.0f cells on axis %d (limit 4000)", ext, k);
    g.dim[k] = (int)ext + 4;                   // + high-side apron (3 cells: one spare for fp32 rounding)
    cells *= g.dim[k];
  }
  if (cells > (double)PGP_MAX_CELLS) return pgp_fail(ctx, PGP_E_TOO_LARGE, "grid needs %.3g cells (limit %lld)", cells, (long long)PGP_MAX_CELLS);
  g.n_cells = (int64_t)g.dim[0] * g.dim[1] * g.dim[2];
  const int64_t nc = g.n_cells;
  const int64_t n_words = ((nc + 31) / 32 + 3) & ~3ll;   // multiple of 16 bytes: bulk-copy granule

  PGP_CUDA(ctx, s.cell_start.reserve((size_t)(nc + 1) * 4));
  PGP_CUDA(ctx, s.cursor.reserve((size_t)(nc + 1) * 4));
  PGP_CUDA(ctx, s.bitmap.reserve((size_t)n_words * 4));
  PGP_CUDA(ctx, s.scratch.reserve((size_t)((nc + 1) / SCAN_BLOCK + (nc + 1) / SCAN_BLOCK / SCAN_BLOCK + 64) * 4 + 64));
  uint32_t* cs = s.cell_start.as<uint32_t>();

  // 3. histogram -> exclusive scan -> scatter -> per-cell order
  PGP_CUDA(ctx, cudaMemsetAsync(cs, 0, (size_t)(nc + 1) * 4, st));
  k1_count<<<B, T, 0, st>>>(s.unsorted.as<float4>(), n, g, s.cell_of.as<uint32_t>(), cs);
  ctx->launches++;
  int rc = pgp_scan_exclusive_u32(ctx, cs, nc + 1, s.scratch.as<uint32_t>());
  if (rc) return rc;
  PGP_CUDA(ctx, cudaMemcpyAsync(s.cursor.p, cs, (size_t)(nc + 1) * 4, cudaMemcpyDeviceToDevice, st));
  k1_scatter<<<B, T, 0, st>>>(s.unsorted.as<float4>(), n, s.cell_of.as<uint32_t>(), s.cursor.as<uint32_t>(), s.pts.as<float4>());
  ctx->launches++;
  unsigned long long* d_occ = reinterpret_cast<unsigned long long*>(ctx->work.as<char>() + 64);
  PGP_CUDA(ctx, cudaMemsetAsync(d_occ, 0, 8, st));
  k1_sort_cells<<<(unsigned)((nc + T - 1) / T), T, 0, st>>>(s.pts.as<float4>(), cs, nc, d_occ);
  ctx->launches++;
  // 4. dilated occupancy bitmap
  PGP_CUDA(ctx, cudaMemsetAsync(s.bitmap.p, 0, (size_t)n_words * 4, st));
  k1_dilate<<<(unsigned)((nc + T - 1) / T), T, 0, st>>>(cs, g, s.bitmap.as<uint32_t>(), n_words);
  ctx->launches++;
  s.bitmap_words = n_words;
  // 5. normals + priors in sorted order
  k1_build_aux<<<B, T, 0, st>>>(s.pts.as<float4>(), n, s.has_nrm ? s.nrm_raw.as<float>() : nullptr, s.prior.as<float>(), s.aux.as<float4>(), s.aux_orig.as<float4>());
  ctx->launches++;
  unsigned long long occ = 0;
  PGP_CUDA(ctx, cudaMemcpyAsync(&occ, d_occ, 8, cudaMemcpyDeviceToHost, st));
  PGP_CUDA(ctx, cudaStreamSynchronize(st));
  PGP_CUDA(ctx, cudaGetLastError());
  s.n_occupied = (int64_t)occ;
  rc = k1_build_fine(ctx);
  if (rc) return rc;
  s.ready = true;
  return PGP_OK;
}

int k1_fill_priors(pgp_ctx* ctx, float v) {
  Scene& s = ctx->scene;
  PGP_CUDA(ctx, s.prior.reserve((size_t)s.n * 4));
  k1_fill<<<(s.n + 255) / 256, 256, 0, ctx->stream>>>(s.prior.as<float>(), s.n, v);
  ctx->launches++;
  s.priors_binary = (v == 0.f || v == 1.f);
  return PGP_OK;
}

// refreshes aux.w after the priors changed and re-derives priors_binary
int k1_refresh_sorted_priors(pgp_ctx* ctx) {
  Scene& s = ctx->scene;
  int* flag = ctx->work.as<int>() + 32;
  PGP_CUDA(ctx, cudaMemsetAsync(flag, 0, 4, ctx->stream));
  k1_check_binary<<<(s.n + 255) / 256, 256, 0, ctx->stream>>>(s.prior.as<float>(), s.n, flag);
  k1_build_aux<<<(s.n + 255) / 256, 256, 0, ctx->stream>>>(s.pts.as<float4>(), s.n, s.has_nrm ? s.nrm_raw.as<float>() : nullptr,
                                                          s.prior.as<float>(), s.aux.as<float4>(), s.aux_orig.as<float4>());
  ctx->launches += 2;
  int h = 0;
  PGP_CUDA(ctx, cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
  PGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  s.priors_binary = (h == 0);
  s.wlists_ready = false; s.wlists_tried = false;      // K1c's cone records depend on the priors: rebuilt by the next WEIGHTED call
  return PGP_OK;
}

int k1_project_priors(pgp_ctx* ctx, const uint16_t* img_dev, int rows, int cols, const float* K) {
  Scene& s = ctx->scene;
  k1_prior_project<<<(s.n + 255) / 256, 256, 0, ctx->stream>>>(s.xyz_raw.as<float>(), s.n, s.cP[0], s.cP[1], s.cP[2], img_dev, rows, cols,
                                                              K[0], K[1], K[2], K[3], K[4], K[5], K[6], K[7], K[8],
                                                              s.prior.as<float>(), ctx->work.as<int>() + 33);
  ctx->launches++;
  PGP_CUDA(ctx, cudaGetLastError());
  return k1_refresh_sorted_priors(ctx);
}
