// prep.cu -- K0: centring, TF32 hi/lo split and row norms of the embedding matrix.
//
// Feeds K1 (gram_tcgen05.cu).  The reference computes WardDistance from raw fp32
// rows (clustering.go:136-145); the Gram identity ||x||^2+||y||^2-2<x,y> needs
//  (1) centred data (distances are translation invariant; removes the common-mean
//      cancellation of non-negative ResNet features),
//  (2) x = s1 + rt with s1 a short fixed-point slice (exact tensor-core accumulation of
//      s1*s1) and rt the TF32-rounded residual; s1*s1 + s1*rt + rt*s1 + rt*rt reproduces
//      the fp32 Gram value to ~1e-8 relative (see split_kernel),
//  (3) norms of the SAME represented values (s1+rt), in double.
// HBM-bound: reads 4ND bytes, writes 8 N_pad D_pad bytes.
#include "common.cuh"
#include "kernels.h"

namespace ic {

__global__ void __launch_bounds__(128) colsum_kernel(const float* __restrict__ x, int64_t n, int64_t d,
                                                     int64_t ldx, double* __restrict__ colsum) {
    const int64_t col = static_cast<int64_t>(blockIdx.x) * 128 + threadIdx.x;
    if (col >= d) return;
    double acc = 0.0;
    for (int64_t r = blockIdx.y; r < n; r += gridDim.y) acc += static_cast<double>(x[r * ldx + col]);
    atomicAdd(&colsum[col], acc);
}

cudaError_t launch_colsum(const float* x, int64_t n, int64_t d, int64_t ldx, double* colsum, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(colsum, 0, sizeof(double) * d, s);
    if (e != cudaSuccess) return e;
    if (n == 0 || d == 0) return cudaSuccess;
    dim3 grid(static_cast<unsigned>((d + 127) / 128), static_cast<unsigned>(n < 592 ? n : 592));
    colsum_kernel<<<grid, 128, 0, s>>>(x, n, d, ldx, colsum);
    return cudaGetLastError();
}

// One block per row.  x_c = x - mean is split as  x_c = s1 + rt (+ dropped remainder):
//   s1 = q_i * rint(x_c / q_i),  q_i = 2^(e_i - bits),  |x_c| < 2^e_i   (fixed point, per-row scale)
//   rt = tf32(x_c - s1)          (x_c - s1 is exact in fp32; |rt| <= q_i / 2)
// With bits = floor(log2(2^24 / d_pad) / 2) every partial sum of s1_i * s1_j over k is an
// integer multiple of q_i q_j below 2^24, so the tensor core's fp32 accumulation of the
// leading product is EXACT (measured: experiments/gram_error.py); the three small products
// go to a second accumulator whose truncation error is ~1e-9 of the Gram value.
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ x, int64_t n, int64_t d, int64_t ldx,
                                                    const double* __restrict__ colsum, int center, int bits,
                                                    float* __restrict__ hi, float* __restrict__ lo,
                                                    double* __restrict__ norms, int64_t d_pad) {
    const int64_t row = blockIdx.x;
    float* hrow = hi + row * d_pad;
    float* lrow = lo + row * d_pad;
    const double inv_n = n > 0 ? 1.0 / static_cast<double>(n) : 0.0;
    __shared__ float redf[8];
    __shared__ double red[8];
    // pass 1: row maximum of |x_c|
    float amax = 0.0f;
    if (row < n)
        for (int64_t k = threadIdx.x; k < d; k += blockDim.x) {
            float v = x[row * ldx + k];
            if (center) v = __fsub_rn(v, static_cast<float>(colsum[k] * inv_n));
            const float a = fabsf(v);
            if (a > amax && a <= 3.0e38f) amax = a;  // NaN / Inf never set the scale
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) redf[threadIdx.x >> 5] = amax;
    __syncthreads();
    amax = redf[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) amax = fmaxf(amax, redf[w]);
    if (amax < 1e-30f) amax = 0.0f;  // degenerate row: everything goes to the residual
    int e = 0;
    frexpf(amax, &e);                        // amax = f * 2^e, f in [0.5, 1)  ->  |x_c| < 2^e
    const float q = ldexpf(1.0f, e - bits);  // quantum of the fixed-point slice
    const float inv_q = ldexpf(1.0f, bits - e);
    double acc = 0.0;
    for (int64_t k = threadIdx.x; k < d_pad; k += blockDim.x) {
        float h = 0.0f, l = 0.0f;
        if (row < n && k < d) {
            float v = x[row * ldx + k];
            if (center) v = __fsub_rn(v, static_cast<float>(colsum[k] * inv_n));
            if (amax > 0.0f && fabsf(v) <= 3.0e38f) h = __fmul_rn(rintf(__fmul_rn(v, inv_q)), q);
            l = to_tf32(__fsub_rn(v, h));  // NaN / Inf inputs stay in the residual and poison the row, as in the reference
            const double rep = static_cast<double>(h) + static_cast<double>(l);
            acc += rep * rep;
        }
        hrow[k] = h;
        lrow[k] = l;
    }
    // block reduce
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        norms[row] = t;
    }
}

// K0' for the int8 path.  One block per row: 22-bit fixed point relative to the row's largest |x_c|, three balanced
// base-128 digits (the leading one uses the whole int8 range).  A row with a non-finite value gets norm = NaN: all its distances become NaN, which never win
// (the reference's strict '<', clustering.go:124).
__global__ void __launch_bounds__(256) split_i8_kernel(const float* __restrict__ x, int64_t n, int64_t d, int64_t ldx,
                                                       const double* __restrict__ colsum, int center,
                                                       int8_t* __restrict__ hs, int8_t* __restrict__ ms,
                                                       int8_t* __restrict__ ls, float* __restrict__ quanta,
                                                       double* __restrict__ norms, int64_t d_pad) {
    const int64_t row = blockIdx.x;
    const double inv_n = n > 0 ? 1.0 / static_cast<double>(n) : 0.0;
    __shared__ float redf[8];
    __shared__ double red[8];
    __shared__ int bad[8];
    float amax = 0.0f;
    int nonfinite = 0;
    if (row < n)
        for (int64_t k = threadIdx.x; k < d; k += blockDim.x) {
            float v = x[row * ldx + k];
            if (center) v = __fsub_rn(v, static_cast<float>(colsum[k] * inv_n));
            const float a = fabsf(v);
            if (!(a <= 3.0e38f)) nonfinite = 1;
            else if (a > amax) amax = a;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        nonfinite |= __shfl_xor_sync(0xffffffffu, nonfinite, o);
    }
    if ((threadIdx.x & 31) == 0) {
        redf[threadIdx.x >> 5] = amax;
        bad[threadIdx.x >> 5] = nonfinite;
    }
    __syncthreads();
    amax = redf[0];
    nonfinite = bad[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
        amax = fmaxf(amax, redf[w]);
        nonfinite |= bad[w];
    }
    if (amax < 1e-30f) amax = 1e-30f;  // an all-zero row: every digit is 0 whatever the scale
    int e = 0;
    const float fr = frexpf(amax, &e);         // amax = fr * 2^e, fr in [0.5, 1)
    // quantum: 22 bits including the sign, v = h*2^14 + m*2^7 + l with h in [-127, 127], m, l in [-64, 63]; the
    // largest representable |v| is 127*2^14 + 63*2^7 + 63 = 2088895 < 2^21: rows whose maximum sits in the top 0.4 %
    // of their binade take the next quantum
    if (fr * 2097152.0f > 2088000.0f) ++e;
    const float q = ldexpf(1.0f, e - 21);
    const float inv_q = ldexpf(1.0f, 21 - e);
    double acc = 0.0;
    for (int64_t k = threadIdx.x; k < d_pad; k += blockDim.x) {
        int hh = 0, mm = 0, ll = 0;
        if (row < n && k < d && !nonfinite) {
            float v = x[row * ldx + k];
            if (center) v = __fsub_rn(v, static_cast<float>(colsum[k] * inv_n));
            const int iv = __float2int_rn(v * inv_q);  // |iv| <= 2088000 (the product by a power of two is exact)
            ll = ((iv + 64) & 127) - 64;         // balanced digits in [-64, 63]; the leading one spans [-127, 127]
            const int v1 = (iv - ll) >> 7;
            mm = ((v1 + 64) & 127) - 64;
            hh = (v1 - mm) >> 7;
            acc += static_cast<double>(iv) * static_cast<double>(iv);
        }
        hs[row * d_pad + k] = static_cast<int8_t>(hh);
        ms[row * d_pad + k] = static_cast<int8_t>(mm);
        ls[row * d_pad + k] = static_cast<int8_t>(ll);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        const double qd = static_cast<double>(q);
        norms[row] = nonfinite ? __longlong_as_double(0x7FF8000000000000ll) : t * qd * qd;
        quanta[row] = q;
    }
}

cudaError_t launch_split_i8(const float* x, int64_t n, int64_t d, int64_t ldx, const double* colsum, int center,
                            int8_t* h, int8_t* m, int8_t* l, float* quanta, double* norms, int64_t n_pad, int64_t d_pad,
                            cudaStream_t s) {
    if (n_pad == 0) return cudaSuccess;
    split_i8_kernel<<<static_cast<unsigned>(n_pad), 256, 0, s>>>(x, n, d, ldx, colsum, center, h, m, l, quanta, norms, d_pad);
    return cudaGetLastError();
}

int split_slice_bits(int64_t d_pad) {
    int bits = 1;
    while (bits < 10 && (static_cast<int64_t>(d_pad) << (2 * (bits + 1))) <= (int64_t(1) << 24)) ++bits;
    return bits;
}

cudaError_t launch_split(const float* x, int64_t n, int64_t d, int64_t ldx, const double* colsum, int center,
                         float* hi, float* lo, double* norms, int64_t n_pad, int64_t d_pad, cudaStream_t s) {
    if (n_pad == 0) return cudaSuccess;
    split_kernel<<<static_cast<unsigned>(n_pad), 256, 0, s>>>(x, n, d, ldx, colsum, center, split_slice_bits(d_pad), hi,
                                                              lo, norms, d_pad);
    return cudaGetLastError();
}

__global__ void fill_kernel(float* __restrict__ p, int64_t count, float value) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        p[i] = value;
}

cudaError_t launch_fill(float* dm, int64_t count, float value, cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    fill_kernel<<<148 * 8, 256, 0, s>>>(dm, count, value);
    return cudaGetLastError();
}

// GenerateLabelVector + the label half of CombineEmbeddings (embeddings.go:166-183): one block per item writes the
// zeros of its label block, then 1.0 at every label index of the item that lies inside the label set.
__global__ void __launch_bounds__(128) label_block_kernel(float* __restrict__ x, int64_t d, int64_t d_img,
                                                          const int32_t* __restrict__ label_offsets,
                                                          const int32_t* __restrict__ label_ids) {
    const int64_t row = blockIdx.x;
    float* lab = x + row * d + d_img;
    const int64_t n_labels = d - d_img;
    for (int64_t j = threadIdx.x; j < n_labels; j += blockDim.x) lab[j] = 0.0f;  // make([]float32, len(labelSet))
    __syncthreads();
    const int32_t b = label_offsets[row], e = label_offsets[row + 1];
    for (int32_t q = b + threadIdx.x; q < e; q += blockDim.x) {
        const int32_t id = label_ids[q];
        if (id >= 0 && id < n_labels) lab[id] = 1.0f;  // labelVector[idx] = 1.0 if the label exists in the set
    }
}

cudaError_t launch_label_block(float* x, int64_t n, int64_t d, int64_t d_img, const int32_t* label_offsets,
                               const int32_t* label_ids, cudaStream_t s) {
    if (n == 0 || d == d_img) return cudaSuccess;
    label_block_kernel<<<static_cast<unsigned>(n), 128, 0, s>>>(x, d, d_img, label_offsets, label_ids);
    return cudaGetLastError();
}

__global__ void init_slots_kernel(SlotKS* __restrict__ ks, int32_t* __restrict__ gkey, int64_t n, int64_t n4) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) ks[i] = make_int2(static_cast<int32_t>(i), 1);  // NewCluster, clustering.go:18-26
    if (i < n4) gkey[i] = i < n ? static_cast<int32_t>(i) : -1;
}

cudaError_t launch_init_slots(SlotKS* ks, int32_t* gkey, int64_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const int64_t n4 = (n + 3) / 4 * 4;
    init_slots_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, s>>>(ks, gkey, n, n4);
    return cudaGetLastError();
}

}  // namespace ic
