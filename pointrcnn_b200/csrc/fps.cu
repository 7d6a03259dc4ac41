// fps.cu -- furthest point sampling for sm_100a: register-resident, cluster-parallel.
//
// Replaces furthest_point_sampling_wrapper (pointnet2_lib/pointnet2/src/sampling.cpp:36-46 ->
// sampling_gpu.cu:93-253) and the gather_operation that follows it (pointnet2_modules.py:32-35).
//
// What the reference computes (the spec): idx[0]=0; every round updates temp[k]=min(temp[k],
// |p_k - p_last|^2) for all k and picks argmax_k temp[k].  Among equal maxima its S-thread strided
// scan + shared-memory tree picks the point with the smallest "rank"
//      rank(k) = bitrev_log2(S)(k mod S) * ceil(n/S) + k div S,   S = 2^floor(log2(min(n,1024)))
// (thread k mod S keeps its lowest k on ties; the tree keeps the lower slot on ties, i.e. compares
// thread ids LSB-first).  That rule is an observable part of the output (RoI-pooled inputs are full
// of duplicates), so it is reproduced exactly -- but not by copying the tree:
//
// Design: points are laid out in RANK order.  Thread g of a scene owns ranks [g*PPT,(g+1)*PPT) with
// xyz and the running min-distance in registers, so "lowest position wins" IS the tie rule at every
// level: strict '>' inside a thread, lowest lane inside a warp (redux.max + ballot + ffs), lowest
// warp inside a CTA, lowest CTA inside a cluster.  One __syncthreads per round.  A scene is spread
// over a thread-block cluster of CS CTAs (CS*THREADS*PPT >= n): each CTA's winner -- distance, rank AND
// coordinates, 32 bytes -- goes to every peer with two st.async stores that complete a transaction count on the
// peer's mbarrier (no cluster barrier in the loop).  A CTA mirrors only its OWN points in shared memory (<= 48 KB, so several clusters share an
// SM); the winner's coordinates travel with the message.  new_xyz is emitted on the fly.
#include <limits.h>

#include "common.cuh"

namespace prb {

struct FpsParams {
    int b, n, m;
    int S, logS, Q;      // reference block size, its log2, ceil(n/S)
    int use_smem_xyz;    // rank-ordered xyz copy fits in shared memory
    const float *xyz;    // (b,n,3)
    float *temp;         // (b,n)
    int *idx;            // (b,m)
    float *new_xyz;      // (b,m,3) or nullptr
};

__device__ __forceinline__ int rank_to_k(int r, int S, int logS, int Q) {
    int brev = r / Q, q = r - brev * Q;
    if (brev >= S) return INT_MAX;
    int t = logS ? (int)(__brev((unsigned)brev) >> (32 - logS)) : 0;
    return q * S + t;
}
__device__ __forceinline__ int k_to_rank(int k, int S, int logS, int Q) {
    int t = k & (S - 1);
    int rv = logS == 0 ? 0 : (int)(__brev((unsigned)t) >> (32 - logS));
    return rv * Q + (k >> logS);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
// ---- cluster exchange primitives
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity)
        : "memory");
}
// 16-byte store into a peer CTA's shared memory that completes 16 tx-bytes on the peer's mbarrier.
// (Measured alternatives, profiles/r1_fps_sweep.json: plain remote stores + tag polling, cluster barriers and
//  per-warp direct pushes are all 1.3-5x slower per round than st.async + mbarrier complete_tx.)
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote_addr),
                 "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_bar)
                 : "memory");
}

constexpr int kMaxWarps = 32;
constexpr int kMaxCluster = 8;

// one exchanged winner: 32 bytes = two 16-byte st.async: {dist bits, rank, x, y} {z, -, -, -}
struct __align__(16) FpsMsg {
    int dist_bits;
    int rank;
    float x, y;
    float z;
    int pad[3];
};

template <int THREADS, int PPT, int CS>
__global__ void __launch_bounds__(THREADS, 1) fps_rank_kernel(const FpsParams p) {
    constexpr int W = THREADS / 32;
    constexpr int CAP = THREADS * PPT;               // ranks owned by this CTA
    extern __shared__ __align__(16) float s_pts[];   // xyz of MY ranks (3 floats per rank), optional for CS == 1
    __shared__ int2 s_wkey[2][kMaxWarps];            // per-warp (dist bits, rank), double buffered
    __shared__ FpsMsg s_slot[2][kMaxCluster];        // per-CTA winners (cluster exchange), double buffered
    __shared__ __align__(8) uint64_t s_bar[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int crank = CS > 1 ? (int)cluster_ctarank() : 0;
    const int scene = blockIdx.x / CS;
    const int n = p.n, m = p.m, S = p.S, logS = p.logS, Q = p.Q;
    const float *xyz = p.xyz + (size_t)scene * n * 3;
    float *temp = p.temp + (size_t)scene * n;
    int *idx = p.idx + (size_t)scene * m;
    float *new_xyz = p.new_xyz ? p.new_xyz + (size_t)scene * m * 3 : nullptr;

    if (CS > 1 && tid == 0) {
        mbar_init(smem_u32(&s_bar[0]), 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // my PPT consecutive ranks: coordinates + running min distance in registers, coordinates mirrored in shared
    // memory so that the CTA's winner can be looked up by rank
    const int g = crank * THREADS + tid;
    float px[PPT], py[PPT], pz[PPT], pd[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        int k = rank_to_k(g * PPT + i, S, logS, Q);
        if (k < n) {
            px[i] = xyz[k * 3 + 0];
            py[i] = xyz[k * 3 + 1];
            pz[i] = xyz[k * 3 + 2];
            pd[i] = temp[k];
        } else {  // hole in the rank space: can never win (real distances are >= 0)
            px[i] = py[i] = pz[i] = 0.f;
            pd[i] = -1.f;
        }
        if (p.use_smem_xyz) {
            s_pts[(tid * PPT + i) * 3 + 0] = px[i];
            s_pts[(tid * PPT + i) * 3 + 1] = py[i];
            s_pts[(tid * PPT + i) * 3 + 2] = pz[i];
        }
    }
    __syncthreads();
    if (CS > 1) cluster_sync_all();

    // round 0: idx[0] = 0 (rank 0 == point 0)
    float cx = xyz[0], cy = xyz[1], cz = xyz[2];
    const bool writer = (tid == 0 && crank == 0);
    if (writer) {
        idx[0] = 0;  // holds RANKS until the fix-up pass below
        if (new_xyz) { new_xyz[0] = cx; new_xyz[1] = cy; new_xyz[2] = cz; }
    }
    uint32_t phase0 = 0, phase1 = 0;

    for (int j = 1; j < m; ++j) {
        const int par = j & 1;
        if (CS > 1 && tid == 0) mbar_arrive_expect_tx(smem_u32(&s_bar[par]), CS * 32);
        float best = -1.f;
        int bslot = 0;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            float d = dist2_ref(px[i] - cx, py[i] - cy, pz[i] - cz);
            d = fminf(d, pd[i]);
            pd[i] = d;
            if (d > best) { best = d; bslot = i; }
        }
        // warp arg-max: distances are >= 0 (or -1 for holes), so their bit patterns order as signed ints
        int vb = __float_as_int(best);
        int wm = __reduce_max_sync(0xffffffffu, vb);
        unsigned bal = __ballot_sync(0xffffffffu, vb == wm);
        int wr = __shfl_sync(0xffffffffu, g * PPT + bslot, __ffs(bal) - 1);
        if (lane == 0) s_wkey[par][warp] = make_int2(wm, wr);
        __syncthreads();

        // CTA-level winner (lowest warp wins ties)
        int2 kv = lane < W ? s_wkey[par][lane] : make_int2(INT_MIN, 0);
        int cm = __reduce_max_sync(0xffffffffu, kv.x);
        unsigned b2 = __ballot_sync(0xffffffffu, kv.x == cm);
        int r = __shfl_sync(0xffffffffu, kv.y, __ffs(b2) - 1);   // winning rank, uniform in the CTA

        if (CS == 1) {
            if (p.use_smem_xyz) {
                cx = s_pts[r * 3 + 0]; cy = s_pts[r * 3 + 1]; cz = s_pts[r * 3 + 2];
            } else {
                int k = rank_to_k(r, S, logS, Q);
                cx = xyz[k * 3 + 0]; cy = xyz[k * 3 + 1]; cz = xyz[k * 3 + 2];
            }
        } else {
            // my CTA's winner (with its coordinates) goes to every CTA of the cluster: two 16-byte st.async per peer
            if (warp == 0 && lane < CS) {
                const int lr = r - crank * CAP;
                const float wx = s_pts[lr * 3 + 0], wy = s_pts[lr * 3 + 1], wz = s_pts[lr * 3 + 2];
                const uint32_t dst = map_to_cta(smem_u32(&s_slot[par][crank]), lane);
                const uint32_t bar = map_to_cta(smem_u32(&s_bar[par]), lane);
                st_async_v4(dst, (uint32_t)cm, (uint32_t)r, __float_as_uint(wx), __float_as_uint(wy), bar);
                st_async_v4(dst + 16, __float_as_uint(wz), 0u, 0u, 0u, bar);
            }
            mbar_wait_cluster(smem_u32(&s_bar[par]), par ? phase1 : phase0);
            if (par) phase1 ^= 1; else phase0 ^= 1;
            const int kx = lane < CS ? s_slot[par][lane].dist_bits : INT_MIN;
            const int gm = __reduce_max_sync(0xffffffffu, kx);
            const unsigned b3 = __ballot_sync(0xffffffffu, kx == gm);
            const FpsMsg &win = s_slot[par][__ffs(b3) - 1];      // lowest CTA wins ties: broadcast LDS
            r = win.rank;
            cx = win.x; cy = win.y; cz = win.z;
        }
        if (writer) {
            idx[j] = r;
            if (new_xyz) { new_xyz[j * 3 + 0] = cx; new_xyz[j * 3 + 1] = cy; new_xyz[j * 3 + 2] = cz; }
        }
    }

    // write back the running minimum distances (the reference leaves them in temp)
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        int k = rank_to_k(g * PPT + i, S, logS, Q);
        if (k < n) temp[k] = pd[i];
    }
    // ranks -> point indices (same CTA wrote them; __syncthreads orders the global accesses)
    __syncthreads();
    if (crank == 0)
        for (int j = tid; j < m; j += THREADS) idx[j] = rank_to_k(idx[j], S, logS, Q);
    if (CS > 1) cluster_sync_all();
}

// Any n: distances stay in global memory (temp), points are visited in the reference's strided order;
// the tie rule is applied as (max distance, then min rank) with two redux ops per level.
__global__ void __launch_bounds__(1024, 1) fps_generic_kernel(const FpsParams p) {
    __shared__ int s_val[2][32];
    __shared__ unsigned s_rank[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int scene = blockIdx.x, n = p.n, m = p.m, S = p.S, logS = p.logS, Q = p.Q;
    const int W = blockDim.x / 32;  // blockDim.x = max(S, 32)
    const float *xyz = p.xyz + (size_t)scene * n * 3;
    float *temp = p.temp + (size_t)scene * n;
    int *idx = p.idx + (size_t)scene * m;
    float *new_xyz = p.new_xyz ? p.new_xyz + (size_t)scene * m * 3 : nullptr;
    int old = 0;
    if (tid == 0) idx[0] = 0;
    for (int j = 0; j < m; ++j) {
        float cx = xyz[old * 3], cy = xyz[old * 3 + 1], cz = xyz[old * 3 + 2];
        if (tid == 0 && new_xyz) { new_xyz[j * 3] = cx; new_xyz[j * 3 + 1] = cy; new_xyz[j * 3 + 2] = cz; }
        if (j == m - 1) break;
        float best = -1.f;
        int besti = 0;
        for (int k = tid; k < n && tid < S; k += S) {  // thread tid plays reference thread tid
            float d = dist2_ref(xyz[k * 3] - cx, xyz[k * 3 + 1] - cy, xyz[k * 3 + 2] - cz);
            d = fminf(d, temp[k]);
            temp[k] = d;
            if (d > best) { best = d; besti = k; }
        }
        unsigned rank = (unsigned)k_to_rank(besti, S, logS, Q);
        int vb = __float_as_int(best);
        int wm = __reduce_max_sync(0xffffffffu, vb);
        unsigned wr = __reduce_min_sync(0xffffffffu, vb == wm ? rank : 0xffffffffu);
        const int par = j & 1;
        if (lane == 0) { s_val[par][warp] = wm; s_rank[par][warp] = wr; }
        __syncthreads();
        int v = lane < W ? s_val[par][lane] : INT_MIN;
        unsigned rk = lane < W ? s_rank[par][lane] : 0xffffffffu;
        int cm = __reduce_max_sync(0xffffffffu, v);
        unsigned cr = __reduce_min_sync(0xffffffffu, v == cm ? rk : 0xffffffffu);
        old = rank_to_k((int)cr, S, logS, Q);
        if (tid == 0) idx[j + 1] = old;
    }
}

template <int THREADS, int PPT, int CS>
static int launch_rank(const FpsParams &p, size_t smem, cudaStream_t stream) {
    auto kern = fps_rank_kernel<THREADS, PPT, CS>;
    // static shared memory counts against the 48 KB default too
    if (smem + 2048 > 48 * 1024) PRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.b * CS);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    PRB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    return check_launch("fps_rank_kernel");
}

template <int CS>
static int dispatch_cs(const FpsParams &p, int threads, int ppt, size_t smem, cudaStream_t st) {
#define PRB_FPS_CASE(T, P) \
    if (threads == T && ppt == P) return launch_rank<T, P, CS>(p, smem, st);
    if (CS == 1) {
        PRB_FPS_CASE(128, 1) PRB_FPS_CASE(128, 2) PRB_FPS_CASE(128, 4)
        PRB_FPS_CASE(256, 8)
        PRB_FPS_CASE(1024, 4) PRB_FPS_CASE(1024, 8)
    }
    PRB_FPS_CASE(256, 4) PRB_FPS_CASE(512, 2) PRB_FPS_CASE(512, 4) PRB_FPS_CASE(512, 8) PRB_FPS_CASE(512, 16)
#undef PRB_FPS_CASE
    set_error("fps: no kernel instance for threads=%d ppt=%d cs=%d", threads, ppt, CS);
    return -1;
}

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

}  // namespace prb

using namespace prb;

extern "C" int prb_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                           float *new_xyz, void *stream) {
    PRB_REQUIRE(b >= 0 && n > 0 && xyz && temp && idx, "fps: bad arguments (b=%d n=%d)", b, n);
    if (m <= 0 || b == 0) return 0;  // reference kernel returns immediately for m <= 0
    cudaStream_t st = (cudaStream_t)stream;
    FpsParams p;
    p.b = b; p.n = n; p.m = m;
    int logS = 0;
    while ((2 << logS) <= n && logS < 10) ++logS;  // S = 2^floor(log2(min(n,1024))), cuda_utils.h:10-14
    p.S = 1 << logS; p.logS = logS; p.Q = ceil_div(n, p.S);
    p.xyz = xyz; p.temp = temp; p.idx = idx; p.new_xyz = new_xyz;
    const int n_pad = p.S * p.Q;

    // configuration: cluster size, threads per CTA, points per thread
    // A cluster of CS CTAs per scene cuts the per-round compute CS-fold; each CTA mirrors only its own slice of the
    // scene in shared memory (<= 48 KB).  Measured (profiles/r1_fps_sweep.json, n=16384, m=4096, ns per round):
    // b=2: CS=8 538, CS=4 614, CS=2 803;  b=16: CS=8 691 (CTAs start sharing SMs), CS=4 617;  b=32: CS=4 617.
    int cs = env_int("PRB_FPS_CS", 0);
    if (cs != 0 && n_pad < 4096) cs = 1;       // the override is meant for the big levels only
    if (cs == 0) {
        cs = 1;
        if (n_pad >= 8192) {
            cs = (long)b * 8 <= num_sms() / 2 ? 8 : 4;
            while (cs > 1 && (long)b * cs > 2L * num_sms()) cs >>= 1;
        }
    }
    while (cs < 8 && ceil_div(n_pad, cs) > 8192) cs <<= 1;  // keep the register-resident path
    while (cs > 1 && (n_pad % (cs * 128)) != 0) cs >>= 1;
    int P = ceil_div(n_pad, cs);
    int threads = env_int("PRB_FPS_THREADS", 0);
    if (threads == 0) {
        threads = 128;
        while (threads < 512 && threads * 4 < P) threads <<= 1;
        if (threads * 16 < P) threads = 1024;
    }
    int ppt = 1;
    while (threads * ppt < P) ppt <<= 1;
    bool generic = (ppt > 16) || (threads == 1024 && ppt > 8) || env_int("PRB_FPS_GENERIC", 0);
    // round (threads, ppt) to an instantiated pair
    if (!generic) {
        if (threads == 128 && ppt > 4) { threads = 256; ppt = ppt / 2; }
        if (threads == 256 && ppt < 4) ppt = 4;
        if (threads == 256 && ppt > 8) { threads = 512; ppt = ppt / 2; }
        if (threads == 512 && ppt < 2) ppt = 2;
        if (threads == 1024 && ppt < 4) ppt = 4;
    }
    if (generic) {
        fps_generic_kernel<<<b, p.S < 32 ? 32 : p.S, 0, st>>>(p);
        return check_launch("fps_generic_kernel");
    }
    // shared copy of the CTA's own coordinates: whole (padded) scene for a single CTA, one slice in a cluster
    size_t smem_pts = (size_t)threads * ppt * 3 * sizeof(float);
    p.use_smem_xyz = (cs > 1 || smem_pts <= 200 * 1024) ? 1 : 0;
    size_t smem = p.use_smem_xyz ? smem_pts : 0;
    switch (cs) {
        case 1: return dispatch_cs<1>(p, threads, ppt, smem, st);
        case 2: return dispatch_cs<2>(p, threads, ppt, smem, st);
        case 4: return dispatch_cs<4>(p, threads, ppt, smem, st);
        case 8: return dispatch_cs<8>(p, threads, ppt, smem, st);
    }
    set_error("fps: bad cluster size %d", cs);
    return -1;
}
