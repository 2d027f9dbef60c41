// Blocked right-looking modular LU for one large matrix and many primes (placeholder until the
// panel / trailing-update kernels land).
#include "lsx_internal.h"

int lsx_blocked_det_residues(lsx_ctx* ctx, const int32_t*, int n, int, int, uint32_t*) {
    return lsx_fail(ctx, LSX_ERR_UNSUPPORTED, "det_large: n = %d is beyond the shared-memory tile path", n);
}
