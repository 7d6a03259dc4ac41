// Which (row, column) of a tensor-memory tile lands in which thread register for tcgen05.ld shape 16x256b?
// Fill 32 lanes x 32 columns through the 32x32b shape with value = row * 100 + col, read back with 16x256b.x2 at lane
// offsets 0 and 16, print thread -> values.   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_layout tmem_layout.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(uint32_t *out) {
    __shared__ uint32_t s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        uint32_t a = (uint32_t)__cvta_generic_to_shared(&s_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(a));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    __syncthreads();
    const uint32_t tmem = s_base + ((uint32_t)(warp * 32) << 16);
    // write: thread = row (lane of this warp's quarter), 32 columns
    for (int c = 0; c < 32; ++c) {
        uint32_t v = (uint32_t)((warp * 32 + lane) * 100 + c);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + c), "r"(v));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    __syncthreads();
    uint32_t r[8];
    for (int half = 0; half < 2; ++half) {
        const uint32_t addr = tmem + ((uint32_t)(half * 16) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int q = 0; q < 8; ++q) out[((warp * 2 + half) * 32 + lane) * 8 + q] = r[q];
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(s_base));
}

int main() {
    uint32_t *d, h[4 * 2 * 32 * 8];
    cudaMalloc(&d, sizeof(h));
    k<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int warp = 0; warp < 2; ++warp)
        for (int half = 0; half < 2; ++half) {
            printf("warp %d lane offset %d\n", warp, half * 16);
            for (int lane = 0; lane < 32; ++lane) {
                printf("  t%02d:", lane);
                for (int q = 0; q < 8; ++q) printf(" %5u", h[((warp * 2 + half) * 32 + lane) * 8 + q]);
                printf("\n");
            }
        }
    return 0;
}
