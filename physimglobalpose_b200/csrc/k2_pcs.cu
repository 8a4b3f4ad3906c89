#include "pgp_internal.cuh"
int k2_extract_pairs(pgp_ctx* ctx, const Model&, float, float, int32_t*, int64_t, int64_t*) { return pgp_fail(ctx, PGP_E_INVALID, "not built yet"); }
int k2_find_quads(pgp_ctx* ctx, const Model&, const int32_t*, float, float, float, const int32_t*, int64_t, const int32_t*, int64_t, int32_t*, int64_t, int64_t*) { return pgp_fail(ctx, PGP_E_INVALID, "not built yet"); }
int k2_rigid_from_quads(pgp_ctx* ctx, const Model&, const int32_t*, const int32_t*, int64_t, float*, uint8_t*) { return pgp_fail(ctx, PGP_E_INVALID, "not built yet"); }
int k2_generate(pgp_ctx* ctx, Model&, const pgp_pcs_opts*, uint64_t, int64_t, int64_t*) { return pgp_fail(ctx, PGP_E_INVALID, "not built yet"); }
