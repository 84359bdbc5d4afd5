// nn_sweep.cu -- K2: first nearest-neighbour sweep over the initial distance matrix.
//
// FindClosestClusters (clustering.go:119-133) rescans the whole lower triangle on
// every iteration.  Here each row s caches its kNNK smallest partners among clusters
// with a LOWER key (key order == the reference's slice order, so "lower key" ==
// "column j < i" of the reference's scan); the global minimum is then a reduction
// over n cached heads, and a row is rescanned only when ALL its cached partners have
// died (the rescans live inside the merge loop, merge_loop.cu).  Before the first merge
// key == slot, so row s sweeps exactly its columns u < s: the strict lower triangle,
// each element read once -- vectorised, coalesced, L1-bypassing loads, then a
// block-wide exact top-k selection of the packed (dist, key) candidates whose u64
// order is the reference's (d, i, j) tie-break.
// Roofline: HBM.  Algorithmic bytes = 4 per pair = 4*N(N-1)/2 per sweep.
#include "common.cuh"
#include "kernels.h"

namespace ic {

namespace {
constexpr int kSweepThreads = 256;

IC_DEVINL void consider(Cand2& c, float v, uint32_t col) {
    const uint32_t bits = __float_as_uint(v);
    if (bits < kMaxFloatBits)  // entries >= MaxFloat32 (and NaN) never win (clustering.go:120,124)
        cand2_insert(c, (static_cast<uint64_t>(bits) << 32) | col, static_cast<int32_t>(col));
}
}  // namespace

__global__ void __launch_bounds__(kSweepThreads) nn_sweep_kernel(const float* __restrict__ dm, int64_t row_begin,
                                                                 int64_t row_end, int64_t ld, SlotNN* __restrict__ nn,
                                                                 int32_t* __restrict__ nn_more) {
    // dm holds the rows [row_begin, row_end).  Long rows first: the triangle's big rows start while the
    // short ones fill the tail
    const int64_t s = row_end - 1 - static_cast<int64_t>(blockIdx.x);
    const float* row = dm + (s - row_begin) * ld;
    Cand2 c;
    cand2_init(c);
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
            const uint32_t col = static_cast<uint32_t>(u + q * kSweepThreads * 4);
            consider(c, v[q].x, col);
            consider(c, v[q].y, col + 1);
            consider(c, v[q].z, col + 2);
            consider(c, v[q].w, col + 3);
        }
    }
    for (; u < vec_end; u += kSweepThreads * 4) {
        const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(row + u));
        const uint32_t col = static_cast<uint32_t>(u);
        consider(c, v.x, col);
        consider(c, v.y, col + 1);
        consider(c, v.z, col + 2);
        consider(c, v.w, col + 3);
    }
    if (threadIdx.x < limit - vec_end) {  // 0..3 tail columns
        const int64_t col = vec_end + threadIdx.x;
        consider(c, row[col], static_cast<uint32_t>(col));
    }
    __shared__ TopKScratch sc;
    bool more = false;
    const int m = block_select_topk<kSweepThreads>(c, sc, more);
    if (threadIdx.x < kNNK) {
        SlotNN out = make_uint4(kNoPartner, kNoPartner, kNoPartner, 0u);
        if (static_cast<int>(threadIdx.x) < m) {
            const uint64_t p = sc.pack[threadIdx.x];
            out = make_uint4(pack_key(p), static_cast<uint32_t>(p >> 32), static_cast<uint32_t>(sc.slot[threadIdx.x]), 1u);
        }
        nn[s * kNNK + threadIdx.x] = out;
    }
    if (threadIdx.x == 0) nn_more[s] = more ? 1 : 0;
}

cudaError_t launch_nn_sweep(const float* dm, int64_t row_begin, int64_t row_end, int64_t ld, SlotNN* nn,
                            int32_t* nn_more, cudaStream_t s) {
    if (row_end <= row_begin) return cudaSuccess;
    nn_sweep_kernel<<<static_cast<unsigned>(row_end - row_begin), kSweepThreads, 0, s>>>(dm, row_begin, row_end, ld, nn,
                                                                                        nn_more);
    return cudaGetLastError();
}

}  // namespace ic
