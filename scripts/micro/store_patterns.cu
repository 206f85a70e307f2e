// Store-pattern microbenchmark (sm_100a): how fast can a grid stream 256-bit stores to HBM when every lane owns
// `own` contiguous bytes (written 32 B per instruction) and `group` neighbouring lanes interleave their chunks?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/store_patterns scripts/micro/store_patterns.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void st32(void* p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    asm volatile("st.global.cs.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

// A warp owns 32 * own contiguous bytes per round.  group = 1: lane l writes bytes [l*own, (l+1)*own) in own/32
// instructions.  group = g: g lanes share g*own bytes; instruction i of lane l writes chunk (i*g + l%g).
__global__ void __launch_bounds__(256) pattern_kernel(uint8_t* out, uint64_t bytes, int own, int group) {
    const uint64_t warp_bytes = 32ull * own;
    const uint64_t n_warp_blocks = bytes / warp_bytes;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t gwarp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int chunks = own / 32;
    for (uint64_t wb = gwarp; wb < n_warp_blocks; wb += n_warps) {
        uint8_t* base = out + wb * warp_bytes + (uint64_t)(lane / group) * group * own;
        for (int i = 0; i < chunks; ++i) {
            const uint64_t v = wb + i;
            st32(base + ((uint64_t)i * group + (lane % group)) * 32, v, v + 1, v + 2, v + 3);
        }
    }
}

int main() {
    const uint64_t bytes = 8ull << 30;
    uint8_t* d;
    cudaMalloc(&d, bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int owns[] = {32, 64, 128, 256};
    const int groups[] = {1, 2, 4, 8};
    for (int own : owns)
        for (int g : groups) {
            if (g > 1 && own == 32) continue;
            for (int ctas_per_sm : {4, 8}) {
                const int grid = 148 * ctas_per_sm;
                pattern_kernel<<<grid, 256>>>(d, bytes, own, g);
                cudaEventRecord(a);
                for (int r = 0; r < 5; ++r) pattern_kernel<<<grid, 256>>>(d, bytes, own, g);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                printf("own %3d B/lane  group %d  ctas/sm %d : %7.1f GB/s  (%s)\n", own, g, ctas_per_sm, 5.0 * bytes / ms / 1e6,
                       cudaGetErrorString(cudaGetLastError()));
            }
        }
    return 0;
}
