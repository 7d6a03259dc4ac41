// mlp_tc.cu -- the per-group / per-point shared MLP as a tcgen05 (5th-gen tensor core) layer chain.
//
// Replaces, for eval-mode forward, everything between the ball query and the next level's features:
//   grouping_operation x2 + cat + SharedMLP (Conv2d 1x1 -> BN -> ReLU) x L + max_pool2d
//     (pointnet2_lib/pointnet2/pointnet2_utils.py:249-257, pointnet2_modules.py:40-52, pytorch_utils.py:5-101)
//   three_interpolate + cat + SharedMLP x L
//     (pointnet2_modules.py:144-156)
// The reference materialises every (B,C,npoint,nsample) tensor in HBM between ~10 library kernels per
// scale; here a 128-row tile never leaves the SM between its gather and its pooled output.
//
// Structure (one CTA per SM, persistent over 128-row tiles; rows = (centre, sample) pairs or points):
//   warps 0-3  "row threads": build the A operand of every layer, chunk by chunk (32 K-columns =
//              one 128-byte swizzle row), into a ring of shared-memory stages:
//                layer 0  : gathered neighbour features / interpolated + skip features / plain rows
//                layer l>0: tcgen05.ld of layer l-1's accumulator columns -> scale/shift/ReLU -> tf32
//              and run the final epilogue (max over nsample / channel-major store).
//   warp 4     weight producer: one thread streams pre-packed weight tiles (already in the UMMA
//              K-major SWIZZLE_128B image, see prb_mlp_pack_weights) with cp.async.bulk + mbarrier tx.
//   warp 5     MMA issuer: one thread issues tcgen05.mma kind::tf32 (M=128, N<=256, K=8) with fp32
//              accumulators in TMEM; tcgen05.commit releases stages / publishes finished layers.
// TMEM plan: layer 0 at column 0, layer 1 at the top (512-N1), layer 2 at column 0 again; chains that
// do not fit (N_l + N_{l+1} > 512) are split into several launches by the host wrapper.
// Precision: operands are rounded to TF32 (cvt.rna), products accumulate in fp32 -- the same contract
// as the reference's cuDNN convolutions under torch's default allow_tf32=True.
#include <string.h>

#include <vector>

#include "common.cuh"

namespace prb {

constexpr int TM = 128;              // rows per tile (= TMEM lanes, UMMA M)
constexpr int KC = 32;               // K columns per chunk (128 bytes of fp32/tf32)
constexpr int A_STAGE_BYTES = TM * KC * 4;        // 16 KB
constexpr int B_TILE_ROWS = 256;                  // max N per MMA / per weight tile
constexpr int B_STAGE_BYTES = B_TILE_ROWS * KC * 4;  // 32 KB
constexpr int NA = 4, NB = 4;        // ring depths
constexpr int ROW_THREADS = 128;
constexpr int CHAIN_THREADS = 192;   // 4 row warps + producer warp + MMA warp
constexpr int MAX_LAYERS = 3;
constexpr int MAX_NP = 512;

enum { IN_SA = 0, IN_FP = 1, IN_DIRECT = 2 };
enum { OUT_ROWS = 0, OUT_SA_MAX = 1, OUT_FP = 2 };

struct ChainParams {
    int mode_in, mode_out, num_layers;
    int nchunks[MAX_LAYERS];   // K chunks per layer
    int np[MAX_LAYERS];        // padded N (multiple of 32)
    int dcol[MAX_LAYERS];      // TMEM column of the accumulator
    const float *w[MAX_LAYERS];      // packed weight images
    const float *scale[MAX_LAYERS];  // np floats (zero padded)
    const float *shift[MAX_LAYERS];
    // layer-0 K segments (each padded to a multiple of KC)
    int nseg, seg_chunks[2], seg_width[2];
    long total_rows;
    int num_tiles;
    // SA
    int n, npoint, ns, c_feat;
    const float *xyz, *new_xyz, *feats_pm;
    const int *idx;
    // FP
    int m, c_known, c_skip;
    const float *known_pm, *weight, *skip;
    // DIRECT
    const float *x_rows;
    int x_pitch;
    // output
    float *out;
    int c_last;          // true channel count of the last layer
    int out_stride_c, out_c_off, out_pitch;
};

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t s2u(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, tf32 inputs, fp32 accumulate, M=128
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void named_bar_rows() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 8-row atoms of 1024 bytes (SBO), version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    // c=f32 (1<<4), a=tf32 (2<<7), b=tf32 (2<<10), a,b K-major, N>>3 at bit 17, M>>4 at bit 24
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
// byte offset of (row r, 16-byte unit j) inside a K-major SWIZZLE_128B stage
__device__ __forceinline__ uint32_t swz(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

struct RingPos {
    uint32_t stage, phase;
    __device__ void advance(int depth) {
        if (++stage == (uint32_t)depth) { stage = 0; phase ^= 1; }
    }
};

struct Smem {
    // 1024-byte aligned stages first
    uint8_t a[NA][A_STAGE_BYTES];
    uint8_t b[NB][B_STAGE_BYTES];
    float scale[MAX_LAYERS][MAX_NP];
    float shift[MAX_LAYERS][MAX_NP];
    // per-tile row metadata
    int row_src[TM][3];      // SA: global point row (slot 0); FP: 3 known rows
    float row_aux[TM][3];    // SA: centre xyz; FP: 3 weights
    int row_valid[TM];
    float red[32][33];       // [channel][partial group] staging of the max-pool epilogue
    uint64_t a_full[NA], a_empty[NA], b_full[NB], b_empty[NB], d_full[MAX_LAYERS];
    uint32_t tmem_base;
};

// ------------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(CHAIN_THREADS, 1) mlp_chain_kernel(const ChainParams p) {
    extern __shared__ uint8_t smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.num_layers;

    if (tid == 0) {
        for (int i = 0; i < NA; ++i) { mbar_init(s2u(&S.a_full[i]), ROW_THREADS); mbar_init(s2u(&S.a_empty[i]), 1); }
        for (int i = 0; i < NB; ++i) { mbar_init(s2u(&S.b_full[i]), 1); mbar_init(s2u(&S.b_empty[i]), 1); }
        for (int i = 0; i < MAX_LAYERS; ++i) mbar_init(s2u(&S.d_full[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) tmem_alloc(s2u(&S.tmem_base), 512);
    for (int l = 0; l < L; ++l)
        for (int i = tid; i < p.np[l]; i += CHAIN_THREADS) { S.scale[l][i] = p.scale[l][i]; S.shift[l][i] = p.shift[l][i]; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    if (warp == 4) {
        // ===================================================== weight producer
        if (lane == 0) {
            RingPos rb = {0, 0};
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                for (int l = 0; l < L; ++l) {
                    const int halves = (p.np[l] + B_TILE_ROWS - 1) / B_TILE_ROWS;
                    for (int kc = 0; kc < p.nchunks[l]; ++kc)
                        for (int h = 0; h < halves; ++h) {
                            const int rows = min(B_TILE_ROWS, p.np[l] - h * B_TILE_ROWS);
                            const uint32_t bytes = (uint32_t)rows * KC * 4;
                            mbar_wait(s2u(&S.b_empty[rb.stage]), rb.phase ^ 1);
                            mbar_expect_tx(s2u(&S.b_full[rb.stage]), bytes);
                            const float *src = p.w[l] + ((size_t)kc * p.np[l] + (size_t)h * B_TILE_ROWS) * KC;
                            bulk_g2s(s2u(S.b[rb.stage]), src, bytes, s2u(&S.b_full[rb.stage]));
                            rb.advance(NB);
                        }
                }
            }
        }
    } else if (warp == 5) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            RingPos ra = {0, 0}, rb = {0, 0};
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                for (int l = 0; l < L; ++l) {
                    const int halves = (p.np[l] + B_TILE_ROWS - 1) / B_TILE_ROWS;
                    for (int kc = 0; kc < p.nchunks[l]; ++kc) {
                        // valid K in this chunk -> number of K=8 steps
                        int valid = KC;
                        if (l == 0) {
                            int c = kc, s = 0;
                            if (p.nseg > 1 && c >= p.seg_chunks[0]) { c -= p.seg_chunks[0]; s = 1; }
                            valid = min(KC, p.seg_width[s] - c * KC);
                        }
                        const int ksteps = (valid + 7) >> 3;
                        mbar_wait(s2u(&S.a_full[ra.stage]), ra.phase);
                        const uint64_t adesc = make_desc(s2u(S.a[ra.stage]));
                        for (int h = 0; h < halves; ++h) {
                            const int rows = min(B_TILE_ROWS, p.np[l] - h * B_TILE_ROWS);
                            mbar_wait(s2u(&S.b_full[rb.stage]), rb.phase);
                            tc_fence_after();
                            const uint64_t bdesc = make_desc(s2u(S.b[rb.stage]));
                            const uint32_t idesc = make_idesc(rows);
                            const uint32_t d = tmem + (uint32_t)(p.dcol[l] + h * B_TILE_ROWS);
                            for (int ks = 0; ks < ksteps; ++ks)  // +32 bytes (= 2 x 16 B) per K=8 step inside the swizzle row
                                umma_tf32(d, adesc + (uint64_t)(2 * ks), bdesc + (uint64_t)(2 * ks), idesc, (kc > 0 || ks > 0) ? 1u : 0u);
                            umma_commit(s2u(&S.b_empty[rb.stage]));
                            rb.advance(NB);
                        }
                        umma_commit(s2u(&S.a_empty[ra.stage]));
                        ra.advance(NA);
                    }
                    umma_commit(s2u(&S.d_full[l]));
                }
            }
        }
    } else {
        // ===================================================== row threads (warps 0-3)
        RingPos ra = {0, 0};
        uint32_t dphase = 0;
        const int r = tid;  // my row inside the tile / my TMEM lane
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const long R = (long)tile * TM + r;
            const bool valid = R < p.total_rows;
            // ---- tile metadata
            named_bar_rows();  // previous tile's readers of S.row_* are done
            S.row_valid[r] = valid;
            if (p.mode_in == IN_SA) {
                long pr = valid ? R / p.ns : 0;                 // global centre index
                int scene = (int)(pr / p.npoint);
                int k = valid ? p.idx[R] : 0;
                S.row_src[r][0] = scene * p.n + k;
                S.row_aux[r][0] = p.new_xyz[pr * 3 + 0];
                S.row_aux[r][1] = p.new_xyz[pr * 3 + 1];
                S.row_aux[r][2] = p.new_xyz[pr * 3 + 2];
            } else if (p.mode_in == IN_FP) {
                long rr = valid ? R : 0;
                int scene = (int)(rr / p.n);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    S.row_src[r][q] = scene * p.m + p.idx[rr * 3 + q];
                    S.row_aux[r][q] = p.weight[rr * 3 + q];
                }
            }
            named_bar_rows();

            // ---- layer 0: build A chunks from global memory
            for (int kc = 0; kc < p.nchunks[0]; ++kc) {
                int c = kc, seg = 0;
                if (p.nseg > 1 && c >= p.seg_chunks[0]) { c -= p.seg_chunks[0]; seg = 1; }
                const int k0 = c * KC;                       // first column of this chunk inside its segment
                const int width = p.seg_width[seg];
                mbar_wait(s2u(&S.a_empty[ra.stage]), ra.phase ^ 1);
                uint8_t *A = S.a[ra.stage];
                const bool rows_seg = (p.mode_in == IN_DIRECT) || (p.mode_in == IN_SA && seg == 0 && p.c_feat > 0) ||
                                      (p.mode_in == IN_FP && seg == 0);
                if (rows_seg) {
                    // point-major sources: 8 lanes cover one row's 128 bytes, a warp covers 4 rows per step
                    const int j = lane & 7;
                    const int kk = k0 + 4 * j;
                    for (int rr = warp * 32 + (lane >> 3); rr < warp * 32 + 32; rr += 4) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (S.row_valid[rr] && kk < width) {
                            if (p.mode_in == IN_FP) {
                                const int C = p.c_known;
                                const float w0 = S.row_aux[rr][0], w1 = S.row_aux[rr][1], w2 = S.row_aux[rr][2];
                                const float *s0 = p.known_pm + (size_t)S.row_src[rr][0] * C + kk;
                                const float *s1 = p.known_pm + (size_t)S.row_src[rr][1] * C + kk;
                                const float *s2 = p.known_pm + (size_t)S.row_src[rr][2] * C + kk;
                                float a0[4], a1[4], a2[4];
                                if ((C & 3) == 0) {
                                    const float4 t0 = __ldg((const float4 *)s0), t1 = __ldg((const float4 *)s1), t2 = __ldg((const float4 *)s2);
                                    a0[0] = t0.x; a0[1] = t0.y; a0[2] = t0.z; a0[3] = t0.w;
                                    a1[0] = t1.x; a1[1] = t1.y; a1[2] = t1.z; a1[3] = t1.w;
                                    a2[0] = t2.x; a2[1] = t2.y; a2[2] = t2.z; a2[3] = t2.w;
                                } else {
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        const bool in = kk + q < width;
                                        a0[q] = in ? __ldg(s0 + q) : 0.f; a1[q] = in ? __ldg(s1 + q) : 0.f; a2[q] = in ? __ldg(s2 + q) : 0.f;
                                    }
                                }
                                float o[4];
#pragma unroll
                                for (int q = 0; q < 4; ++q)  // same contraction as three_interpolate (interpolate_gpu.cu:96)
                                    o[q] = __fmaf_rn(w2, a2[q], __fmaf_rn(w0, a0[q], __fmul_rn(w1, a1[q])));
                                v = make_float4(o[0], o[1], o[2], o[3]);
                            } else {
                                const float *src;
                                int pitch;
                                if (p.mode_in == IN_DIRECT) { pitch = p.x_pitch; src = p.x_rows + ((size_t)tile * TM + rr) * pitch + kk; }
                                else { pitch = p.c_feat; src = p.feats_pm + (size_t)S.row_src[rr][0] * pitch + kk; }
                                if ((pitch & 3) == 0 && kk + 3 < width) {
                                    v = __ldg((const float4 *)src);
                                } else {
                                    float o[4];
#pragma unroll
                                    for (int q = 0; q < 4; ++q) o[q] = (kk + q < width) ? __ldg(src + q) : 0.f;
                                    v = make_float4(o[0], o[1], o[2], o[3]);
                                }
                            }
                        }
                        v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
                        *reinterpret_cast<float4 *>(A + swz(rr, j)) = v;
                    }
                } else if (p.mode_in == IN_SA) {
                    // relative xyz segment: [x - cx, y - cy, z - cz, 0 ...]; one K=8 step is consumed
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid) {
                        const float *q = p.xyz + (size_t)S.row_src[r][0] * 3;
                        v.x = to_tf32(q[0] - S.row_aux[r][0]);
                        v.y = to_tf32(q[1] - S.row_aux[r][1]);
                        v.z = to_tf32(q[2] - S.row_aux[r][2]);
                    }
                    *reinterpret_cast<float4 *>(A + swz(r, 0)) = v;
                    *reinterpret_cast<float4 *>(A + swz(r, 1)) = make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    // FP skip segment: channel-major (b, c_skip, n); lanes run along consecutive points
                    const long rr0 = valid ? R : 0;
                    const int scene = (int)(rr0 / p.n), u = (int)(rr0 - (long)scene * p.n);
                    const float *base = p.skip + (size_t)scene * p.c_skip * p.n + u;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float o[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int ch = k0 + 4 * j + q;
                            o[q] = (valid && ch < width) ? to_tf32(__ldg(base + (size_t)ch * p.n)) : 0.f;
                        }
                        *reinterpret_cast<float4 *>(A + swz(r, j)) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                }
                fence_async_smem();
                mbar_arrive(s2u(&S.a_full[ra.stage]));
                ra.advance(NA);
            }

            // ---- layers 1..L-1: previous accumulator -> scale/shift/ReLU -> next A operand
            for (int l = 1; l < L; ++l) {
                mbar_wait(s2u(&S.d_full[l - 1]), dphase);
                tc_fence_after();
                for (int kc = 0; kc < p.nchunks[l]; ++kc) {
                    uint32_t acc[32];
                    tmem_ld32(tmem + lane_base + (uint32_t)(p.dcol[l - 1] + kc * KC), acc);
                    mbar_wait(s2u(&S.a_empty[ra.stage]), ra.phase ^ 1);
                    uint8_t *A = S.a[ra.stage];
                    const float *sc = &S.scale[l - 1][kc * KC], *sh = &S.shift[l - 1][kc * KC];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float o[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            o[q] = to_tf32(fmaxf(fmaf(__uint_as_float(acc[4 * j + q]), sc[4 * j + q], sh[4 * j + q]), 0.f));
                        *reinterpret_cast<float4 *>(A + swz(r, j)) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                    tc_fence_before();
                    fence_async_smem();
                    mbar_arrive(s2u(&S.a_full[ra.stage]));
                    ra.advance(NA);
                }
            }

            // ---- final epilogue
            mbar_wait(s2u(&S.d_full[L - 1]), dphase);
            tc_fence_after();
            const int Cl = p.c_last;
            for (int c0 = 0; c0 < Cl; c0 += 32) {
                uint32_t acc[32];
                tmem_ld32(tmem + lane_base + (uint32_t)(p.dcol[L - 1] + c0), acc);
                float v[32];
#pragma unroll
                for (int q = 0; q < 32; ++q)
                    v[q] = fmaxf(fmaf(__uint_as_float(acc[q]), S.scale[L - 1][c0 + q], S.shift[L - 1][c0 + q]), 0.f);
                if (p.mode_out == OUT_ROWS) {
                    if (valid) {
                        float *o = p.out + (size_t)R * p.out_pitch + c0;
#pragma unroll
                        for (int q = 0; q < 32; q += 4)
                            *reinterpret_cast<float4 *>(o + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
                    }
                } else if (p.mode_out == OUT_FP) {
                    if (valid) {
                        const int scene = (int)(R / p.n), u = (int)(R - (long)scene * p.n);
                        float *o = p.out + ((size_t)scene * p.out_stride_c + p.out_c_off + c0) * p.n + u;
#pragma unroll
                        for (int q = 0; q < 32; ++q)
                            if (c0 + q < Cl) o[(size_t)q * p.n] = v[q];
                    }
                } else {
                    // max over the nsample consecutive rows of each centre.  Groups are whole (total_rows is a
                    // multiple of nsample), so rows of the tail tile past total_rows form groups that are skipped.
                    const int ns = p.ns;
                    const int w = ns < 32 ? ns : 32;
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        float x = v[q];
                        for (int off = 1; off < w; off <<= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, off));
                        v[q] = x;
                    }
                    const int gpw = 32 / w;                   // partial groups per warp
                    const int ppc = ns > 32 ? ns / 32 : 1;    // partials per centre
                    const int G = TM / ns;                    // centres per tile
                    named_bar_rows();                         // previous readers of S.red are done
                    if ((lane % w) == 0) {
                        const int pg = warp * gpw + lane / w;
#pragma unroll
                        for (int q = 0; q < 32; ++q) S.red[q][pg] = v[q];
                    }
                    named_bar_rows();
                    for (int e = r; e < 32 * G; e += ROW_THREADS) {
                        const int q = e / G, g = e - q * G;
                        float x = S.red[q][g * ppc];
                        for (int t = 1; t < ppc; ++t) x = fmaxf(x, S.red[q][g * ppc + t]);
                        const long Rg = (long)tile * TM + (long)g * ns;
                        if (Rg < p.total_rows && c0 + q < Cl) {
                            const long pr = Rg / ns;
                            const int scene = (int)(pr / p.npoint), pp = (int)(pr - (long)scene * p.npoint);
                            p.out[((size_t)scene * p.out_stride_c + p.out_c_off + c0 + q) * p.npoint + pp] = x;
                        }
                    }
                }
            }
            tc_fence_before();
            dphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ host side
static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct LayerGeom {
    int k_chunks;  // K chunks
    int np;        // padded N
    size_t w_off;  // float offset of this layer's image in the packed buffer
};

// K layout of layer 0 = the given segments, each padded to KC; deeper layers: previous np
static void chain_geometry(int L, int nseg, const int *seg_width, const int *c_out, LayerGeom *g, size_t *total_floats) {
    size_t off = 0;
    for (int l = 0; l < L; ++l) {
        int chunks = 0;
        if (l == 0) for (int s = 0; s < nseg; ++s) chunks += round_up(seg_width[s], KC) / KC;
        else chunks = g[l - 1].np / KC;
        g[l].k_chunks = chunks;
        g[l].np = round_up(c_out[l], 32);
        g[l].w_off = off;
        off += (size_t)chunks * g[l].np * KC;
    }
    if (total_floats) *total_floats = off;
}

static inline float tf32_rna_host(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) != 0x7f800000u) u += 0x1000u;  // round to nearest, ties away (cvt.rna)
    u &= 0xffffe000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

}  // namespace prb

using namespace prb;

// segment description used by the packer: our K order is seg 0 then seg 1; each segment names the first
// column it takes from the ORIGINAL (c_out0, c_in) weight and its width
struct PackSegs {
    int nseg;
    int src_off[2], width[2];
};

static void default_segs(int kind, int c_in, int a, PackSegs *ps) {
    // kind 0 (SA): original columns [xyz(3), feats(a)] -> ours [feats(a), xyz(3)];  a = c_feat
    // kind 1 (FP): original [interp(a), skip(c_in-a)] kept in order, split in two padded segments
    // kind 2 (DIRECT): one segment
    if (kind == 0) {
        if (a > 0) { ps->nseg = 2; ps->src_off[0] = 3; ps->width[0] = a; ps->src_off[1] = 0; ps->width[1] = 3; }
        else { ps->nseg = 1; ps->src_off[0] = 0; ps->width[0] = 3; }
    } else if (kind == 1 && c_in - a > 0) {
        ps->nseg = 2; ps->src_off[0] = 0; ps->width[0] = a; ps->src_off[1] = a; ps->width[1] = c_in - a;
    } else {
        ps->nseg = 1; ps->src_off[0] = 0; ps->width[0] = c_in;
    }
}

extern "C" {

// extended packing entry points (kind / split) -- the header's prb_mlp_packed_bytes / prb_mlp_pack_weights
// are the DIRECT (kind 2) forms
PRB_API size_t prb_mlp_packed_bytes_ex(int kind, int split, int num_layers, int c_in, const int *c_out) {
    PackSegs ps;
    default_segs(kind, c_in, split, &ps);
    LayerGeom g[MAX_LAYERS];
    size_t total = 0;
    chain_geometry(num_layers, ps.nseg, ps.width, c_out, g, &total);
    return total * sizeof(float);
}

PRB_API int prb_mlp_pack_weights_ex(int kind, int split, int num_layers, int c_in, const int *c_out, const float *const *w, void *dst) {
    PRB_REQUIRE(num_layers >= 1 && num_layers <= MAX_LAYERS && w && dst && c_out, "mlp_pack: bad arguments");
    PackSegs ps;
    default_segs(kind, c_in, split, &ps);
    LayerGeom g[MAX_LAYERS];
    size_t total = 0;
    chain_geometry(num_layers, ps.nseg, ps.width, c_out, g, &total);
    float *out = (float *)dst;
    memset(out, 0, total * sizeof(float));
    for (int l = 0; l < num_layers; ++l) {
        const int np = g[l].np;
        const int kin = l == 0 ? c_in : c_out[l - 1];
        // map our K index -> original column (or -1 for padding)
        std::vector<int> kmap((size_t)g[l].k_chunks * KC, -1);
        if (l == 0) {
            int base = 0;
            for (int s = 0; s < ps.nseg; ++s) {
                for (int i = 0; i < ps.width[s]; ++i) kmap[base + i] = ps.src_off[s] + i;
                base += round_up(ps.width[s], KC);
            }
        } else {
            for (int i = 0; i < kin; ++i) kmap[i] = i;
        }
        for (int kc = 0; kc < g[l].k_chunks; ++kc)
            for (int n = 0; n < np; ++n)
                for (int kk = 0; kk < KC; ++kk) {
                    const int src = kmap[(size_t)kc * KC + kk];
                    float v = 0.f;
                    if (n < c_out[l] && src >= 0) v = tf32_rna_host(w[l][(size_t)n * kin + src]);
                    // K-major SWIZZLE_128B image of an (np x 32) tile: row n at n*128 B, 16-B unit j at j ^ (n & 7)
                    const int j = kk >> 2, q = kk & 3;
                    const size_t o = g[l].w_off + (size_t)kc * np * KC + (size_t)n * KC + (size_t)(((j ^ (n & 7)) << 2) + q);
                    out[o] = v;
                }
    }
    return 0;
}

size_t prb_mlp_packed_bytes(int num_layers, int c_in, const int *c_out) {
    return prb_mlp_packed_bytes_ex(2, 0, num_layers, c_in, c_out);
}
int prb_mlp_pack_weights(int num_layers, int c_in, const int *c_out, const float *const *w, void *dst) {
    return prb_mlp_pack_weights_ex(2, 0, num_layers, c_in, c_out, w, dst);
}

}  // extern "C"

namespace prb {

// launch one fused segment [l0, l1) of the chain
static int launch_chain(ChainParams &p, cudaStream_t st) {
    static bool attr_set = false;
    const size_t smem = sizeof(Smem) + 1024;
    if (!attr_set) {
        PRB_CUDA(cudaFuncSetAttribute(mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    p.num_tiles = (int)((p.total_rows + TM - 1) / TM);
    if (p.num_tiles == 0) return 0;
    int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    mlp_chain_kernel<<<grid, CHAIN_THREADS, smem, st>>>(p);
    return check_launch("mlp_chain_kernel");
}

// TMEM feasibility of fusing layers [l0, l1): accumulators ping-pong between column 0 and the top
static bool fits(const LayerGeom *g, int l0, int l1) {
    const int n = l1 - l0;
    if (n == 1) return g[l0].np <= 512;
    if (n == 2) return g[l0].np + g[l0 + 1].np <= 512;
    return g[l0].np + g[l0 + 1].np <= 512 && g[l0 + 2].np + g[l0 + 1].np <= 512;
}

struct ChainIO {
    int kind, split;          // packing kind (0 SA, 1 FP, 2 DIRECT) and its split argument
    ChainParams base;         // mode_in / sources / output filled by the caller
    long rows;
};

static size_t chain_workspace_bytes(long rows, int L, int kind, int c_in, int split, const int *c_out) {
    PackSegs ps;
    default_segs(kind, c_in, split, &ps);
    LayerGeom g[MAX_LAYERS];
    chain_geometry(L, ps.nseg, ps.width, c_out, g, nullptr);
    // worst case: two ping-pong row buffers of the widest padded layer
    int wmax = 0;
    for (int l = 0; l < L; ++l) wmax = g[l].np > wmax ? g[l].np : wmax;
    bool split_needed = !fits(g, 0, L);
    const size_t rows_pad = (size_t)((rows + TM - 1) / TM * TM);
    return split_needed ? 2 * (rows_pad * wmax * sizeof(float) + 256) : 256;
}

// run the whole chain, splitting where TMEM cannot hold two consecutive accumulators
static int run_chain(const ChainIO &io, const prb_mlp_desc *mlp, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    const int L = mlp->num_layers;
    PRB_REQUIRE(L >= 1 && L <= MAX_LAYERS, "mlp: num_layers %d unsupported", L);
    PackSegs ps;
    default_segs(io.kind, mlp->c_in, io.split, &ps);
    LayerGeom g[MAX_LAYERS];
    chain_geometry(L, ps.nseg, ps.width, mlp->c_out, g, nullptr);
    for (int l = 0; l < L; ++l) PRB_REQUIRE(g[l].np <= MAX_NP, "mlp: layer width %d > %d unsupported", mlp->c_out[l], MAX_NP);
    size_t soff[MAX_LAYERS];
    size_t so = 0;
    for (int l = 0; l < L; ++l) { soff[l] = so; so += (size_t)g[l].np; }

    const long rows_pad = (io.rows + TM - 1) / TM * TM;
    int l0 = 0;
    const float *cur_rows = nullptr;
    int cur_pitch = 0;
    int pingpong = 0;
    while (l0 < L) {
        int l1 = L;
        while (l1 > l0 + 1 && !fits(g, l0, l1)) --l1;
        ChainParams p = io.base;
        p.total_rows = io.rows;
        p.num_layers = l1 - l0;
        for (int l = l0; l < l1; ++l) {
            const int i = l - l0;
            p.nchunks[i] = g[l].k_chunks;
            p.np[i] = g[l].np;
            p.w[i] = mlp->packed_w + g[l].w_off;
            p.scale[i] = mlp->scale + soff[l];
            p.shift[i] = mlp->shift + soff[l];
        }
        p.dcol[0] = 0;
        if (p.num_layers >= 2) p.dcol[1] = 512 - p.np[1];
        if (p.num_layers >= 3) p.dcol[2] = 0;
        if (l0 > 0) {  // continue from materialised rows
            p.mode_in = IN_DIRECT;
            p.x_rows = cur_rows;
            p.x_pitch = cur_pitch;
            p.nseg = 1; p.seg_chunks[0] = g[l0].k_chunks; p.seg_width[0] = cur_pitch;
        } else {
            p.nseg = ps.nseg;
            for (int s = 0; s < ps.nseg; ++s) { p.seg_chunks[s] = round_up(ps.width[s], KC) / KC; p.seg_width[s] = ps.width[s]; }
        }
        if (l1 < L) {  // materialise this segment's activations as rows
            size_t buf_bytes = (size_t)rows_pad * g[l1 - 1].np * sizeof(float);
            PRB_REQUIRE(workspace && workspace_bytes >= 2 * buf_bytes, "mlp: workspace too small (%zu < %zu)", workspace_bytes, 2 * buf_bytes);
            float *dst = (float *)((char *)workspace + (pingpong ? workspace_bytes / 2 / 256 * 256 : 0));
            p.mode_out = OUT_ROWS;
            p.out = dst;
            p.out_pitch = g[l1 - 1].np;
            p.c_last = g[l1 - 1].np;
            int rc = launch_chain(p, st);
            if (rc) return rc;
            cur_rows = dst;
            cur_pitch = g[l1 - 1].np;
            pingpong ^= 1;
        } else {
            p.c_last = mlp->c_out[L - 1];
            int rc = launch_chain(p, st);
            if (rc) return rc;
        }
        l0 = l1;
    }
    return 0;
}

}  // namespace prb

extern "C" {

PRB_API size_t prb_sa_workspace_bytes(int b, int npoint, int nsample, int c_feat, int num_layers, const int *c_out) {
    return chain_workspace_bytes((long)b * npoint * nsample, num_layers, 0, 3 + c_feat, c_feat, c_out);
}
PRB_API size_t prb_fp_workspace_bytes(int b, int n, int c_known, int c_skip, int num_layers, const int *c_out) {
    return chain_workspace_bytes((long)b * n, num_layers, 1, c_known + c_skip, c_known, c_out);
}
PRB_API size_t prb_rows_workspace_bytes(long rows, int c_in, int num_layers, const int *c_out) {
    return chain_workspace_bytes(rows, num_layers, 2, c_in, 0, c_out);
}

PRB_API int prb_sa_group_mlp_max_ws(int b, int n, int npoint, int nsample, int c_feat, const float *xyz, const float *new_xyz,
                                    const float *feats_pm, const int *idx, const prb_mlp_desc *mlp, float *out, int out_stride_c,
                                    int out_c_off, void *workspace, size_t workspace_bytes, void *stream) {
    PRB_REQUIRE(b >= 0 && n > 0 && npoint > 0 && nsample > 0 && xyz && new_xyz && idx && mlp && out, "sa_group_mlp_max: bad arguments");
    PRB_REQUIRE(mlp->c_in == 3 + c_feat, "sa_group_mlp_max: c_in %d != 3 + c_feat %d", mlp->c_in, c_feat);
    PRB_REQUIRE(c_feat == 0 || feats_pm, "sa_group_mlp_max: features missing");
    PRB_REQUIRE(nsample >= 4 && nsample <= 128 && (nsample & (nsample - 1)) == 0, "sa_group_mlp_max: nsample %d must be a power of two in [4,128]", nsample);
    if (b == 0) return 0;
    ChainIO io;
    memset(&io, 0, sizeof(io));
    io.kind = 0; io.split = c_feat;
    io.rows = (long)b * npoint * nsample;
    io.base.mode_in = IN_SA; io.base.mode_out = OUT_SA_MAX;
    io.base.n = n; io.base.npoint = npoint; io.base.ns = nsample; io.base.c_feat = c_feat;
    io.base.xyz = xyz; io.base.new_xyz = new_xyz; io.base.feats_pm = feats_pm; io.base.idx = idx;
    io.base.out = out; io.base.out_stride_c = out_stride_c; io.base.out_c_off = out_c_off;
    return run_chain(io, mlp, workspace, workspace_bytes, (cudaStream_t)stream);
}

PRB_API int prb_fp_interp_mlp_ws(int b, int n, int m, int c_known, int c_skip, const float *known_pm, const int *idx,
                                 const float *weight, const float *skip, const prb_mlp_desc *mlp, float *out, void *workspace,
                                 size_t workspace_bytes, void *stream) {
    PRB_REQUIRE(b >= 0 && n > 0 && m > 0 && c_known > 0 && known_pm && idx && weight && mlp && out, "fp_interp_mlp: bad arguments");
    PRB_REQUIRE(mlp->c_in == c_known + c_skip, "fp_interp_mlp: c_in %d != %d + %d", mlp->c_in, c_known, c_skip);
    PRB_REQUIRE(c_skip == 0 || skip, "fp_interp_mlp: skip features missing");
    if (b == 0) return 0;
    ChainIO io;
    memset(&io, 0, sizeof(io));
    io.kind = 1; io.split = c_known;
    io.rows = (long)b * n;
    io.base.mode_in = IN_FP; io.base.mode_out = OUT_FP;
    io.base.n = n; io.base.m = m; io.base.c_known = c_known; io.base.c_skip = c_skip;
    io.base.known_pm = known_pm; io.base.idx = idx; io.base.weight = weight; io.base.skip = skip;
    io.base.out = out; io.base.out_stride_c = mlp->c_out[mlp->num_layers - 1]; io.base.out_c_off = 0;
    return run_chain(io, mlp, workspace, workspace_bytes, (cudaStream_t)stream);
}

// plain rows -> MLP -> rows (also the unit-test doorway of the tensor-core chain)
PRB_API int prb_mlp_rows(long rows, int c_in, const float *x_rows, const prb_mlp_desc *mlp, float *out_rows, int out_pitch,
                         void *workspace, size_t workspace_bytes, void *stream) {
    PRB_REQUIRE(rows >= 0 && c_in > 0 && x_rows && mlp && out_rows, "mlp_rows: bad arguments");
    PRB_REQUIRE(mlp->c_in == c_in, "mlp_rows: c_in mismatch");
    const int np_last = (mlp->c_out[mlp->num_layers - 1] + 31) / 32 * 32;
    PRB_REQUIRE(out_pitch >= np_last && (out_pitch & 3) == 0, "mlp_rows: out_pitch %d must be >= %d and a multiple of 4", out_pitch, np_last);
    if (rows == 0) return 0;
    ChainIO io;
    memset(&io, 0, sizeof(io));
    io.kind = 2; io.split = 0;
    io.rows = rows;
    io.base.mode_in = IN_DIRECT; io.base.mode_out = OUT_ROWS;
    io.base.x_rows = x_rows; io.base.x_pitch = c_in;
    io.base.out = out_rows; io.base.out_pitch = out_pitch;
    return run_chain(io, mlp, workspace, workspace_bytes, (cudaStream_t)stream);
}

// header forms without an explicit workspace: valid only for chains that need no split
int prb_sa_group_mlp_max(int b, int n, int npoint, int nsample, int c_feat, const float *xyz, const float *new_xyz,
                         const float *feats_pm, const int *idx, const prb_mlp_desc *mlp, float *out, int out_stride_c,
                         int out_c_off, void *stream) {
    return prb_sa_group_mlp_max_ws(b, n, npoint, nsample, c_feat, xyz, new_xyz, feats_pm, idx, mlp, out, out_stride_c, out_c_off,
                                   nullptr, 0, stream);
}
int prb_fp_interp_mlp(int b, int n, int m, int c_known, int c_skip, const float *known_pm, const int *idx, const float *weight,
                      const float *skip, const prb_mlp_desc *mlp, float *out, void *stream) {
    return prb_fp_interp_mlp_ws(b, n, m, c_known, c_skip, known_pm, idx, weight, skip, mlp, out, nullptr, 0, stream);
}

}  // extern "C"
