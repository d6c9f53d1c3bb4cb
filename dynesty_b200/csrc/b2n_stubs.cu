// Entry points that are declared in include/b200nest.h but not implemented yet
// return B2N_ERR_UNSUPPORTED (loud failure, never a CPU fallback).
#include "b2n_common.cuh"
extern "C" {
int b2n_rslice_batch(b2n_ctx*, const b2n_chain_args*, int32_t, int32_t, double*, double*, double*, int32_t*, int32_t*, int32_t*, uint32_t*) { return B2N_ERR_UNSUPPORTED; }
int b2n_slice_batch(b2n_ctx*, const b2n_chain_args*, int32_t, int32_t, double*, double*, double*, int32_t*, int32_t*, int32_t*, uint32_t*) { return B2N_ERR_UNSUPPORTED; }
int b2n_unif_batch(b2n_ctx*, const b2n_chain_args*, double*, double*, double*, int32_t*, int32_t*, uint32_t*) { return B2N_ERR_UNSUPPORTED; }
}
