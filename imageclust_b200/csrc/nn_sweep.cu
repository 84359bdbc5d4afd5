// nn_sweep.cu -- K2: first nearest-neighbour sweep over the initial distance matrix.
//
// FindClosestClusters (clustering.go:119-133) rescans the whole lower triangle on
// every iteration.  Here each row s caches its best partner among clusters with a
// LOWER key (key order == the reference's slice order, so "lower key" == "column
// j < i" of the reference's scan); the global minimum is then a reduction over n
// cached candidates and a row is rescanned only when its partner dies (the
// rescans live inside the merge loop, merge_loop.cu).  Before the first merge
// key == slot, so row s sweeps exactly its columns u < s: the strict lower
// triangle, each element read once -- vectorised, coalesced, L1-bypassing loads,
// warp-shuffle + shared-memory min of the packed (dist, key) candidate whose u64
// order is the reference's (d, i, j) tie-break.
// Roofline: HBM.  Algorithmic bytes = 4 per pair = 4*N(N-1)/2 per sweep.
#include "common.cuh"
#include "kernels.h"

namespace ic {

namespace {
constexpr int kSweepThreads = 256;
}

__global__ void __launch_bounds__(kSweepThreads) nn_sweep_kernel(const float* __restrict__ dm, int64_t n, int64_t ld,
                                                                 SlotNN* __restrict__ nn) {
    // long rows first: the triangle's big rows start while the short ones fill the tail
    const int64_t s = n - 1 - static_cast<int64_t>(blockIdx.x);
    const float* row = dm + s * ld;
    uint64_t best = kPackInf;
    const int64_t limit = s;  // partners are exactly the columns u < s
    const int64_t vec_end = limit & ~int64_t(3);
    // 4 independent 16-byte loads in flight per thread
    int64_t u = static_cast<int64_t>(threadIdx.x) * 4;
    for (; u + 3 * kSweepThreads * 4 < vec_end; u += 4 * kSweepThreads * 4) {
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = ld_stream_f4(reinterpret_cast<const float4*>(row + u + q * kSweepThreads * 4));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t c = static_cast<uint32_t>(u + q * kSweepThreads * 4);
            best = umin64(best, pack_cand(v[q].x, c));
            best = umin64(best, pack_cand(v[q].y, c + 1));
            best = umin64(best, pack_cand(v[q].z, c + 2));
            best = umin64(best, pack_cand(v[q].w, c + 3));
        }
    }
    for (; u < vec_end; u += kSweepThreads * 4) {
        const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(row + u));
        const uint32_t c = static_cast<uint32_t>(u);
        best = umin64(best, pack_cand(v.x, c));
        best = umin64(best, pack_cand(v.y, c + 1));
        best = umin64(best, pack_cand(v.z, c + 2));
        best = umin64(best, pack_cand(v.w, c + 3));
    }
    if (threadIdx.x < limit - vec_end) {  // 0..3 tail columns
        const int64_t c = vec_end + threadIdx.x;
        best = umin64(best, pack_cand(row[c], static_cast<uint32_t>(c)));
    }
    __shared__ uint64_t red[kSweepThreads / 32];
    best = warp_min_u64(best);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < kSweepThreads / 32; ++w) best = umin64(best, red[w]);
        SlotNN out = make_uint4(kNoPartner, kNoPartner, kNoPartner, 0u);
        if (pack_selectable(best)) {  // entries >= MaxFloat32 never win (clustering.go:120,124)
            const uint32_t pk = pack_key(best);
            out = make_uint4(pk, static_cast<uint32_t>(best >> 32), pk, 1u);
        }
        nn[s] = out;
    }
}

cudaError_t launch_nn_sweep(const float* dm, int64_t n, int64_t ld, SlotNN* nn, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    nn_sweep_kernel<<<static_cast<unsigned>(n), kSweepThreads, 0, s>>>(dm, n, ld, nn);
    return cudaGetLastError();
}

}  // namespace ic
