#include "pgp_internal.cuh"
int k5_tricp(pgp_ctx* ctx, Model&, const float*, int, double*, int, float, float, int, int*, float*) { return pgp_fail(ctx, PGP_E_INVALID, "not built yet"); }
