// Fused register-resident kernels for small shapes (placeholder: nothing covered yet).
#include "lsx_internal.h"

int lsx_run_small(lsx_ctx*, const ElimJob&, int* handled) {
    *handled = 0;
    return LSX_OK;
}
