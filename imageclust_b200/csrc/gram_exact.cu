// gram_exact.cu -- audit path for K1: initial Ward distances with the reference's
// OWN arithmetic, bit for bit.
//
// WardDistance (clustering.go:136-145) for singletons: diff = a[k]-b[k] rounded to
// fp32, sum += diff*diff in index order with separately rounded multiply and add
// (DotFloat32, clustering.go:152-155; Go/amd64 does not fuse), then
// (float32(1*1)/float32(1+1)) * sum.  __fsub_rn/__fmul_rn/__fadd_rn forbid FMA
// contraction, so every entry equals the CPU oracle's exactly.  SIMT only: it is
// the on-device truth that the tensor-core kernel is checked against, and the
// re-evaluator for near-tie audits; 3 flops per (pair, dimension) on the FMA pipe.
#include "common.cuh"
#include "kernels.h"

namespace ic {

namespace {
constexpr int T = 64;    // output tile
constexpr int KT = 16;   // k tile
}

__global__ void __launch_bounds__(256) gram_exact_kernel(const float* __restrict__ x, int64_t n, int64_t d,
                                                         int64_t ldx, float* __restrict__ dm, int64_t ld,
                                                         int64_t row_begin, int64_t row_end) {
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;  // lower triangle of tiles only
    // dm holds the rows [row_begin, row_end) (a rank's row block, or the whole matrix)
    // (a tile is needed if its rows OR its columns -- the mirrored entries -- touch the resident rows)
    const int64_t r0 = static_cast<int64_t>(bi) * T, c0 = static_cast<int64_t>(bj) * T;
    if ((r0 >= row_end || r0 + T <= row_begin) && (c0 >= row_end || c0 + T <= row_begin)) return;
    __shared__ float sa[T][KT + 1];
    __shared__ float sb[T][KT + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t row0 = static_cast<int64_t>(bi) * T, col0 = static_cast<int64_t>(bj) * T;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;

    for (int64_t k0 = 0; k0 < d; k0 += KT) {
        for (int e = threadIdx.x; e < T * KT; e += 256) {
            const int r = e / KT, k = e % KT;
            const int64_t gk = k0 + k;
            const int64_t ga = row0 + r, gb = col0 + r;
            sa[r][k] = (ga < n && gk < d) ? x[ga * ldx + gk] : 0.0f;
            sb[r][k] = (gb < n && gk < d) ? x[gb * ldx + gk] : 0.0f;
        }
        __syncthreads();
        const int kmax = (d - k0) < KT ? static_cast<int>(d - k0) : KT;
        for (int k = 0; k < kmax; ++k) {  // ascending k: the reference's summation order
            float a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = sa[ty * 4 + r][k];
#pragma unroll
            for (int c = 0; c < 4; ++c) b[c] = sb[tx * 4 + c][k];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float diff = __fsub_rn(a[r], b[c]);
                    acc[r][c] = __fadd_rn(acc[r][c], __fmul_rn(diff, diff));
                }
        }
        __syncthreads();
    }
    const float w = __fdiv_rn(1.0f, 2.0f);  // float32(1*1) / float32(1+1), clustering.go:142-144
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int64_t gi = row0 + ty * 4 + r, gj = col0 + tx * 4 + c;
            if (gi >= n || gj >= n) continue;
            if (gj < gi) {
                const float v = __fmul_rn(w, acc[r][c]);
                if (gi >= row_begin && gi < row_end) dm[(gi - row_begin) * ld + gj] = v;
                if (gj >= row_begin && gj < row_end) dm[(gj - row_begin) * ld + gi] = v;  // mirrored entry, if resident
            } else if (gj == gi) {
                if (gi >= row_begin && gi < row_end) dm[(gi - row_begin) * ld + gi] = 0.0f;
            }
        }
}

cudaError_t launch_gram_exact(const float* x, int64_t n, int64_t d, int64_t ldx, float* dm, int64_t ld,
                              int64_t row_begin, int64_t row_end, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const unsigned nb = static_cast<unsigned>((n + T - 1) / T);
    dim3 grid(nb, nb);
    gram_exact_kernel<<<grid, 256, 0, s>>>(x, n, d, ldx, dm, ld, row_begin, row_end);
    return cudaGetLastError();
}

}  // namespace ic
