// K6 -- explained-point removal before the per-node trimmed ICP of the MCTS search.
//
// UCTState::performTrICP (PPE/src/hypothesis_verification/mcts/UCTState.cpp:121-204): before an object's pose is refined
// against its segment, every segment point within pointRemovalThreshold (8 mm, UCTSearch.cpp) of the objects ALREADY PLACED
// is dropped (:149-174).  The reference builds the "explained" cloud by transforming the CURRENT object's model cloud with
// each placed object's pose (sic, :150-155: pcl::transformPointCloud in float) and runs one FLANN radius search per segment
// point.  Here: the same transformed cloud (fp32, row-wise left to right like pcl::transformPointCloud), then an exact
// all-pairs sweep over shared-memory tiles -- segment x explained is at most a few 10^8 distance tests, and it keeps the
// reference's arithmetic (squared distance summed left to right, strict d2 < r^2 as FLANN's radius result set).
// RESTATED, PARITY UNPINNED: PCL / FLANN are not vendored in the reference tree; the CPU restatement the tests
// check this against is lo_remove_explained in the test infrastructure.
#include "pgp_internal.cuh"

namespace {

__global__ void k6_transform(const float4* __restrict__ model, int nv, const float* __restrict__ T12, int n_placed, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv * n_placed) return;
  const float* T = T12 + 12 * (i / nv);
  const float4 p = model[i % nv];
  out[i] = make_float4(xf_row(T[0], T[1], T[2], T[3], p.x, p.y, p.z), xf_row(T[4], T[5], T[6], T[7], p.x, p.y, p.z),
                       xf_row(T[8], T[9], T[10], T[11], p.x, p.y, p.z), 0.f);
}

constexpr int K6_T = 256;
__global__ void __launch_bounds__(K6_T) k6_mark(const float4* __restrict__ seg, int ns, const float4* __restrict__ expl, int ne, float r2,
                                                uint8_t* __restrict__ flag) {
  __shared__ float4 tile[K6_T];
  const int i = blockIdx.x * K6_T + threadIdx.x;
  const float4 s = i < ns ? seg[i] : make_float4(0, 0, 0, 0);
  bool hit = false;
  for (int j0 = 0; j0 < ne; j0 += K6_T) {
    __syncthreads();
    if (j0 + (int)threadIdx.x < ne) tile[threadIdx.x] = expl[j0 + threadIdx.x];
    __syncthreads();
    if (__syncthreads_and(hit || i >= ns)) break;          // the whole CTA is done
    const int m = min(K6_T, ne - j0);
    if (!hit && i < ns) {
      for (int t = 0; t < m; ++t) {
        const float4 q = tile[t];
        const float dx = __fsub_rn(s.x, q.x), dy = __fsub_rn(s.y, q.y), dz = __fsub_rn(s.z, q.z);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));   // L2_Simple: left to right
        if (d2 < r2) { hit = true; break; }
      }
    }
  }
  if (i < ns) flag[i] = hit ? 1 : 0;
}

// owned by the context, released by k6_release (pgp_destroy)
struct K6Scratch { DevBuf seg, expl, T, flag; };
K6Scratch& k6_scratch_of(pgp_ctx* ctx) {
  if (!ctx->k6_scratch) ctx->k6_scratch = new K6Scratch();
  return *static_cast<K6Scratch*>(ctx->k6_scratch);
}

}  // namespace

// flags_host[i] = 1 when segment point i is explained by a placed object.  Returns the number of UNEXPLAINED points.
int k6_remove_explained(pgp_ctx* ctx, Model& m, const float* seg_xyz_host, int ns, const double* placed16_host, int n_placed, float threshold,
                        uint8_t* flags_host, int* n_unexplained) {
  K6Scratch& sc = k6_scratch_of(ctx);
  *n_unexplained = ns;
  if (ns <= 0) return PGP_OK;
  if (n_placed <= 0) { memset(flags_host, 0, (size_t)ns); return PGP_OK; }
  cudaStream_t st = ctx->stream;
  std::vector<float> seg4((size_t)ns * 4, 0.f), T((size_t)n_placed * 12);
  for (int i = 0; i < ns; ++i) for (int c = 0; c < 3; ++c) seg4[4 * (size_t)i + c] = seg_xyz_host[3 * i + c];
  for (int k = 0; k < n_placed; ++k)                       // utilities::convertToMatrix: Isometry3d -> Matrix4f
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) T[12 * (size_t)k + 4 * r + c] = (float)placed16_host[16 * k + 4 * r + c];
  const int ne = m.nv * n_placed;
  PGP_CUDA(ctx, sc.seg.reserve((size_t)ns * 16));
  PGP_CUDA(ctx, sc.expl.reserve((size_t)ne * 16));
  PGP_CUDA(ctx, sc.T.reserve((size_t)n_placed * 48));
  PGP_CUDA(ctx, sc.flag.reserve((size_t)ns));
  PGP_CUDA(ctx, cudaMemcpyAsync(sc.seg.p, seg4.data(), (size_t)ns * 16, cudaMemcpyHostToDevice, st));
  PGP_CUDA(ctx, cudaMemcpyAsync(sc.T.p, T.data(), (size_t)n_placed * 48, cudaMemcpyHostToDevice, st));
  k6_transform<<<(ne + 255) / 256, 256, 0, st>>>(m.val_raw.as<float4>(), m.nv, sc.T.as<float>(), n_placed, sc.expl.as<float4>());
  k6_mark<<<(ns + K6_T - 1) / K6_T, K6_T, 0, st>>>(sc.seg.as<float4>(), ns, sc.expl.as<float4>(), ne, threshold * threshold, sc.flag.as<uint8_t>());
  ctx->launches += 2;
  PGP_CUDA(ctx, cudaGetLastError());
  PGP_CUDA(ctx, cudaMemcpyAsync(flags_host, sc.flag.p, (size_t)ns, cudaMemcpyDeviceToHost, st));
  PGP_CUDA(ctx, cudaStreamSynchronize(st));
  int kept = 0;
  for (int i = 0; i < ns; ++i) kept += flags_host[i] ? 0 : 1;
  *n_unexplained = kept;
  return PGP_OK;
}

void k6_release(pgp_ctx* ctx) {
  if (!ctx->k6_scratch) return;
  K6Scratch* sc = static_cast<K6Scratch*>(ctx->k6_scratch);
  for (DevBuf* b : {&sc->seg, &sc->expl, &sc->T, &sc->flag}) b->release();
  delete sc;
  ctx->k6_scratch = nullptr;
}
