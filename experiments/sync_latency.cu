// Experiment (GPU): latency of the primitives the merge loop's exchange is built from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o experiments/sync_latency experiments/sync_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t ld_vol(const uint32_t* p) { uint32_t v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint32_t ld_rlx(const uint32_t* p) { uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint32_t ld_acq(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_vol(uint32_t* p, uint32_t v) { asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_rlx(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_rel(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void fence() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// mode 0: volatile ld/st; 1: relaxed.gpu; 2: relaxed + fence before store; 3: st.release / ld.acquire;
// 4: relaxed + fence + 64 scattered stores before each fence
__global__ void pingpong(uint32_t* flags, float* junk, int iters, int mode, long long* out) {
    uint32_t* mine = flags + blockIdx.x * 64;          // separate lines
    uint32_t* other = flags + (1 - blockIdx.x) * 64;
    if (threadIdx.x != 0) {
        return;
    }
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (blockIdx.x == 0) {
            if (mode == 4) for (int j = 0; j < 64; ++j) junk[(size_t)(i * 64 + j) * 4099 % (1 << 24)] = (float)i;
            if (mode == 2 || mode == 4) fence();
            if (mode == 0) st_vol(mine, i); else if (mode == 3) st_rel(mine, i); else st_rlx(mine, i);
            if (mode == 0) while (ld_vol(other) != (uint32_t)i) {}
            else if (mode == 3) while (ld_acq(other) != (uint32_t)i) {}
            else while (ld_rlx(other) != (uint32_t)i) {}
            if (mode == 2 || mode == 4) fence();
        } else {
            if (mode == 0) while (ld_vol(other) != (uint32_t)i) {}
            else if (mode == 3) while (ld_acq(other) != (uint32_t)i) {}
            else while (ld_rlx(other) != (uint32_t)i) {}
            if (mode == 4) for (int j = 0; j < 64; ++j) junk[(size_t)(i * 64 + j + 7) * 4099 % (1 << 24)] = (float)i;
            if (mode == 2 || mode == 4) fence();
            if (mode == 0) st_vol(mine, i); else if (mode == 3) st_rel(mine, i); else st_rlx(mine, i);
        }
    }
    if (blockIdx.x == 0) out[0] = (clock64() - t0) / iters;
}

// all-to-all with private mailboxes: every block pushes a tagged 16-byte word to every reader, then polls its own
__global__ void alltoall(uint4* mail, int iters, int use_fence, long long* out) {
    const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        __syncthreads();
        if (t < G) {
            if (use_fence) fence();
            uint4* dst = mail + ((size_t)t * 2 + (i & 1)) * G + b;
            asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(b), "r"(t), "r"(i), "r"(i) : "memory");
            const uint4* src = mail + ((size_t)b * 2 + (i & 1)) * G + t;
            uint4 v;
            do {
                asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src) : "memory");
            } while (v.w != (uint32_t)i);
            if (use_fence) fence();
        }
        __syncthreads();
    }
    if (b == 0 && t == 0) out[0] = (clock64() - t0) / iters;
}

// dependent global loads: L2-hit and DRAM round trips as seen by one thread
__global__ void chase(const uint32_t* next, int iters, long long* out) {
    uint32_t p = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) p = __ldcg(next + p);
    out[0] = (clock64() - t0) / iters;
    out[1] = p;
}

int main() {
    uint32_t* flags; float* junk; long long* out; uint4* mail;
    cudaMalloc(&flags, 4096); cudaMalloc(&junk, sizeof(float) << 24); cudaMalloc(&out, 64);
    cudaMalloc(&mail, sizeof(uint4) * 148 * 2 * 148);
    long long h[2];
    const char* names[] = {"volatile", "relaxed.gpu", "relaxed+fence", "release/acquire", "relaxed+fence+64 scattered stores"};
    for (int mode = 0; mode < 5; ++mode) {
        cudaMemset(flags, 0, 4096);
        void* args[] = {&flags, &junk, nullptr, &mode, &out};
        int iters = 2000; args[2] = &iters;
        cudaLaunchCooperativeKernel((void*)pingpong, dim3(2), dim3(32), args, 0, 0);
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("pingpong %-36s %6lld cycles per round trip (2 one-way hops)\n", names[mode], h[0]);
    }
    for (int G : {2, 8, 37, 74, 148})
        for (int f = 0; f < 2; ++f) {
            cudaMemset(mail, 0, sizeof(uint4) * 148 * 2 * 148);
            int iters = 2000;
            void* args[] = {&mail, &iters, &f, &out};
            cudaLaunchCooperativeKernel((void*)alltoall, dim3(G), dim3(256), args, 0, 0);
            cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("alltoall G=%3d fence=%d: %6lld cycles per exchange (%s)\n", G, f, h[0], cudaGetErrorString(e));
        }
    {
        // pointer chase over 32 MB (L2 resident) and 1 GB (DRAM)
        for (size_t words : {size_t(1) << 23, size_t(1) << 28}) {
            uint32_t* next; cudaMalloc(&next, words * 4);
            uint32_t* hn = (uint32_t*)malloc(words * 4);
            const size_t stride = 1031 * 32;  // jump ~132 KB
            for (size_t i = 0; i < words; ++i) hn[i] = (uint32_t)((i + stride) % words);
            cudaMemcpy(next, hn, words * 4, cudaMemcpyHostToDevice);
            int iters = 4000;
            chase<<<1, 1>>>(next, iters, out);  // warm
            chase<<<1, 1>>>(next, iters, out);
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("dependent ld.cg over %4zu MB: %lld cycles per load\n", words * 4 >> 20, h[0]);
            cudaFree(next); free(hn);
        }
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SM clock (attr) %d kHz\n", clk);
    return 0;
}
