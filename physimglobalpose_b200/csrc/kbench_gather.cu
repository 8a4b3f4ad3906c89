// Random 32-byte-sector gather micro-benchmark: the roofline denominator SURVEY.md 8(d) asks for next to K3.
//
// K3's phase 1 is one divergent 4-byte gather per query into the L2-resident label table: each lane's load moves one 32-byte
// sector L2 -> L1, and (almost) no two lanes of a request share a 128-byte line.  This kernel issues exactly that traffic pattern
// with nothing else in the way -- every thread reads pseudo-random, 32-byte-aligned words of a buffer, eight independent loads in
// flight per thread, 2048 threads per SM -- and reports sectors x 32 B / time.  With a footprint below the L2 capacity it measures
// the L2 -> SM random-sector throughput; with a footprint far above it, the HBM random-sector throughput.  No reference counterpart.
#include "pgp_internal.cuh"

namespace {

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

constexpr int GU = 8;     // loads in flight per thread

__global__ void __launch_bounds__(256) kb_sector_gather(const uint32_t* __restrict__ buf, uint32_t sector_mask, int rounds, uint32_t* __restrict__ sink) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  uint32_t ctr = tid * 0x9e3779b9u;
  for (int r = 0; r < rounds; ++r) {
    uint32_t v[GU];
#pragma unroll
    for (int u = 0; u < GU; ++u) {
      const uint32_t s = hash32(ctr + (uint32_t)(r * GU + u) * 0x85ebca6bu) & sector_mask;
      v[u] = __ldg(buf + (size_t)s * 8u);              // one 4-byte word of 32-byte sector s
    }
#pragma unroll
    for (int u = 0; u < GU; ++u) acc ^= v[u];
  }
  if (acc == 0x12345678u) sink[0] = acc;               // keeps the loads alive
}

}  // namespace

extern "C" int pgp_bench_sector_gather(pgp_ctx* ctx, int64_t footprint_bytes, int loads_per_thread, float* gbps) {
  if (!ctx || !gbps) return PGP_E_INVALID;
  cudaSetDevice(ctx->device);
  if (footprint_bytes < 4096 || loads_per_thread < GU) return pgp_fail(ctx, PGP_E_INVALID, "pgp_bench_sector_gather: footprint >= 4096 bytes, loads >= %d", GU);
  int64_t sectors = 1;
  while (sectors * 2 * 32 <= footprint_bytes) sectors *= 2;          // power of two, <= footprint
  if (sectors > (1ll << 31)) sectors = 1ll << 31;
  DevBuf buf, sink;
  PGP_CUDA(ctx, buf.reserve((size_t)sectors * 32));
  PGP_CUDA(ctx, sink.reserve(64));
  cudaStream_t st = ctx->stream;
  PGP_CUDA(ctx, cudaMemsetAsync(buf.p, 0, (size_t)sectors * 32, st));
  const int rounds = loads_per_thread / GU;
  const int grid = ctx->sm_count * 8;
  cudaEvent_t a, b;
  PGP_CUDA(ctx, cudaEventCreate(&a));
  PGP_CUDA(ctx, cudaEventCreate(&b));
  float best_ms = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {                                 // rep 0 warms the L2
    PGP_CUDA(ctx, cudaEventRecord(a, st));
    kb_sector_gather<<<grid, 256, 0, st>>>(buf.as<uint32_t>(), (uint32_t)(sectors - 1), rounds, sink.as<uint32_t>());
    PGP_CUDA(ctx, cudaEventRecord(b, st));
    PGP_CUDA(ctx, cudaEventSynchronize(b));
    float ms = 0.f;
    PGP_CUDA(ctx, cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best_ms) best_ms = ms;
    ctx->launches++;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  PGP_CUDA(ctx, cudaGetLastError());
  buf.release(); sink.release();
  const double loads = (double)grid * 256.0 * (double)rounds * GU;
  *gbps = (float)(loads * 32.0 / ((double)best_ms * 1e-3) / 1e9);
  return PGP_OK;
}
