// latency of the warp-collective primitives the FPS round is built from (dependent chains, cycles per op)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_prims warp_prims.cu ; run: ./warp_prims
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 512
__device__ __forceinline__ float credux_max_f32(float v) { float r; asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int OP>
__global__ void k(float *out, long long *cyc, int dummy) {
    __shared__ float s[64];
    __shared__ __align__(8) unsigned long long bar;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 64) s[threadIdx.x] = (float)threadIdx.x * 0.5f;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"((int)blockDim.x));
    __syncthreads();
    float v = (float)lane + (float)dummy;
    int iv = lane + dummy;
    unsigned phase = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        if (OP == 0) { v = credux_max_f32(v) + (float)lane * 1e-3f; }
        if (OP == 1) { iv = __reduce_max_sync(0xffffffffu, iv) + lane; }
        if (OP == 2) { unsigned b = __ballot_sync(0xffffffffu, iv & 1); iv = (int)b + lane; }
        if (OP == 3) { iv = __shfl_sync(0xffffffffu, iv, (iv + 1) & 31) + 1; }
        if (OP == 4) { iv = __ffs(iv | 0x10000) + lane; }
        if (OP == 5) { v = s[((int)v) & 31] + 1.f; }
        if (OP == 6) { __syncthreads(); }
        if (OP == 7) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
            asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar)), "r"(phase) : "memory");
            phase ^= 1;
        }
        if (OP == 8) {   // full in-warp argmax stage: credux + vote + ffs + shfl
            float m = credux_max_f32(v);
            unsigned b = __ballot_sync(0xffffffffu, v == m);
            int src = __ffs(b) - 1;
            iv = __shfl_sync(0xffffffffu, iv, src);
            v = v + (float)(iv & 1) + (float)lane * 1e-3f;
        }
        if (OP == 9) {   // two-credux stage: max value then max payload among ties
            float m = credux_max_f32(v);
            unsigned pl = __reduce_max_sync(0xffffffffu, v == m ? (unsigned)iv : 0u);
            iv = (int)pl + lane;
            v = v + (float)(iv & 1) + (float)lane * 1e-3f;
        }
        if (OP == 10) { v = fmaxf(v * 1.0001f, 0.5f) ; }   // FP dependent op for scale
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = v + (float)iv;
}
template <int OP> void run(const char *name, int threads) {
    float *o; long long *c; cudaMalloc(&o, 4096 * 4); cudaMalloc(&c, 8);
    k<OP><<<1, threads>>>(o, c, 0); k<OP><<<1, threads>>>(o, c, 0);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-44s threads %4d : %.1f cycles/op  (%s)\n", name, threads, (double)h / N, cudaGetErrorString(cudaGetLastError()));
    cudaFree(o); cudaFree(c);
}
int main() {
    for (int th : {32, 512}) {
        run<0>("CREDUX.MAX.F32 (+FADD)", th); run<1>("REDUX.MAX.S32 (+IADD)", th); run<2>("VOTE ballot (+IADD)", th);
        run<3>("SHFL.IDX (+IADD)", th); run<4>("ffs (BREV+FLO) (+IADD)", th); run<5>("LDS dependent (+FADD)", th);
        run<6>("__syncthreads", th); run<7>("mbarrier arrive + try_wait (all threads)", th);
        run<8>("stage: credux+vote+ffs+shfl", th); run<9>("stage: credux + redux.max.u32 payload", th); run<10>("FMUL+FMNMX chain", th);
    }
    return 0;
}
