// nn_sweep.cu -- K2: per-row nearest-neighbour sweep over the distance matrix.
//
// FindClosestClusters (clustering.go:119-133) rescans the whole lower triangle on
// every iteration.  Here each row s caches its best partner among clusters with a
// LOWER key (key order == the reference's slice order, so "lower key" == "column
// j < i" of the reference's scan); the global minimum is then a reduction over n
// cached candidates and a row is rescanned only when its partner dies.  This
// kernel is the full sweep (one block per row): vectorised, coalesced, read-once
// loads; warp-shuffle + shared-memory min of the packed (dist, key) candidate,
// whose u64 order is the reference's (d, i, j) tie-break.
// Roofline: HBM.  Algorithmic bytes = 4 per live pair read.
#include "common.cuh"
#include "kernels.h"

namespace ic {

template <bool kIdentity>
__global__ void __launch_bounds__(256) nn_sweep_kernel(const float* __restrict__ dm, int64_t n, int64_t ld,
                                                       const int32_t* __restrict__ key,
                                                       const int32_t* __restrict__ size,
                                                       const int32_t* __restrict__ slot_of_key,
                                                       unsigned long long* __restrict__ nn_pack,
                                                       int2* __restrict__ nn_aux) {
    const int64_t s = blockIdx.x;
    const int32_t ks = kIdentity ? static_cast<int32_t>(s) : key[s];
    uint64_t best = kPackInf;
    if (ks >= 0) {
        const float* row = dm + s * ld;
        const int64_t limit = kIdentity ? s : n;  // identity keys: partners are exactly the columns j < s
        for (int64_t u = static_cast<int64_t>(threadIdx.x) * 4; u < limit; u += 256 * 4) {
            const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(row + u));
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int64_t uu = u + e;
                if (uu >= limit) continue;
                const int32_t ku = kIdentity ? static_cast<int32_t>(uu) : key[uu];
                if (ku < 0 || ku >= ks) continue;
                best = umin64(best, pack_cand(vv[e], static_cast<uint32_t>(ku)));
            }
        }
    }
    __shared__ uint64_t red[8];
    best = warp_min_u64(best);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) best = umin64(best, red[w]);
        int2 aux = make_int2(-1, 0);
        if (pack_selectable(best)) {
            const int32_t pk = static_cast<int32_t>(pack_key(best));
            aux.x = kIdentity ? pk : slot_of_key[pk];
            aux.y = size[aux.x];
        } else {
            best = kPackInf;
        }
        nn_pack[s] = best;
        nn_aux[s] = aux;
    }
}

cudaError_t launch_nn_sweep(const float* dm, int64_t n, int64_t ld, const int32_t* key, const int32_t* size,
                            const int32_t* slot_of_key, int identity_keys, unsigned long long* nn_pack,
                            int2* nn_aux, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    if (identity_keys)
        nn_sweep_kernel<true><<<static_cast<unsigned>(n), 256, 0, s>>>(dm, n, ld, key, size, slot_of_key, nn_pack, nn_aux);
    else
        nn_sweep_kernel<false><<<static_cast<unsigned>(n), 256, 0, s>>>(dm, n, ld, key, size, slot_of_key, nn_pack, nn_aux);
    return cudaGetLastError();
}

}  // namespace ic
