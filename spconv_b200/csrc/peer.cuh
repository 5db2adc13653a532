// Internal interface of peer.cu (weight-gradient exchange over NVLink peer memory).
#pragma once
#include "common.cuh"

namespace spx {

// push: this rank's `total` fp32 values -- the sum of `chunks` split-K partials (`partial` != NULL, row stride
// `stride` floats) or the tensor `src` of `dtype` -- go to every rank's exchange buffer.  Never waits.
int peer_push(const float *partial, int64_t stride, int chunks, const void *src, int64_t total, int dtype,
              const spx_peer_group *pg, cudaStream_t stream);
// finish: wait for every rank's push, dst = scale * sum over ranks (rank order, one rounding to `dtype`).
int peer_finish(void *dst, int64_t total, int dtype, const spx_peer_group *pg, float scale, cudaStream_t stream);

}  // namespace spx
