// Internal: aggregation-based additive coarse correction used as a right preconditioner by the Krylov driver.
#pragma once
#include "tefem.h"
#include <cuda_runtime.h>

// z = omega * r + P * Ac^{-1} * P^T r   (see tefem_precond.cu); r, z: owned fine rows (4 * n_rows doubles)
int tefem_precond_apply_stream(tefem_precond* pc, const double* r, double* z, cudaStream_t st);
