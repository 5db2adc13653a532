// Internal interface of peer.cu (weight-gradient exchange over NVLink peer memory).
#pragma once
#include "common.cuh"

namespace spx {

// Sum `total` fp32 values over the ranks of `pg` and write them (x scale, rounded once) to `dst` in `dtype`.
// The local contribution is either the sum of `chunks` split-K partials (`partial` != NULL, row stride
// `stride` floats) or the tensor `src` (may alias `dst`).
int peer_reduce_exchange(const float *partial, int64_t stride, int chunks, const void *src, int64_t total, void *dst,
                         int dtype, const spx_peer_group *pg, float scale, cudaStream_t stream);

}  // namespace spx
