// sweep_cluster.cu -- the temporal-layer recurrences on thread-block clusters (H = 128, 256, 512).
//
// Same arithmetic as sweep.cu (reference loops brnnet.py:144-152 forward, :208-224 BPTT), but the
// hidden state never touches global memory on the serial chain:
//
//   * one CLUSTER of CS CTAs per (direction, tile of NB utterances); CTA r owns ROWS output units
//     (ROWS x NB = 256) and keeps its ROWS x H slice of the recurrent matrix in registers for the whole sweep;
//   * every CTA holds the complete previous state of its 8 utterances in shared memory, laid out
//     [slice][utterance][32] so that one CTA's contribution is one contiguous 1 KB block;
//   * after a step, each CTA pushes its 1 KB block into the shared memory of all CS CTAs of the
//     cluster with cp.async.bulk (shared::cta -> shared::cluster); the copies complete_tx on the
//     DESTINATION's mbarrier, which is the only synchronisation of the step: a CTA starts step s as
//     soon as the CS blocks of step s-1 have landed.  Two state buffers / two barriers alternate;
//   * For/Back (dFor/dBack) are still streamed to HBM for the GEMMs that follow, off the chain.
#include "common.cuh"
#include <stdlib.h>

namespace ctcb {

constexpr int SC_THREADS = 256;

struct SweepClusterArgs {
    int mode, T, B, H;
    const int32_t *Tlen;
    const float *pre;
    const float *W[2];
    float *out[2];
    const float *act[2];
    float maxAct;
    unsigned int *err;      // [0] set to 2 if a barrier wait timed out (never a hang)
    int ndir;               // gridDim.z: 2 = both directions, 1 = forward in time only
    int opt;                // tuning bits (CTCB_SWEEP_OPT): 1 = one polling lane per warp, 2 = HBM store after the push
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 1-D bulk async copy: this CTA's shared memory -> a cluster peer's shared memory, signalling the
// peer's mbarrier with the byte count when the data has landed.
__device__ __forceinline__ void bulk_push(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}

// KI = H/32.  RPW = output units per warp (4 or 8); a CTA owns ROWS = 8*RPW units of NB = 32/RPW
// utterances, so its block is always 1 KB and the cluster has CS = H/ROWS CTAs.  Fewer, fatter CTAs
// (RPW = 8) halve both the number of pushes and the shared-memory traffic per step; the price is
// 8*KI weight registers per lane.  grid = (CS, ntiles, 2), cluster = (CS, 1, 1).
template <int KI, int RPW>
__global__ void __launch_bounds__(SC_THREADS, 1) sweep_cluster_kernel(SweepClusterArgs a) {
    constexpr int H = 32 * KI;
    constexpr int SC_NB = 32 / RPW;
    constexpr int ROWS = 8 * RPW;
    constexpr int CS = H / ROWS;
    constexpr uint32_t BLK_BYTES = SC_NB * ROWS * sizeof(float);         // one CTA's block: 1 KB
    __shared__ __align__(128) float hbuf[2][CS * SC_NB * ROWS];          // [buffer][slice][utterance][ROWS]
    __shared__ __align__(128) float stage[2][SC_NB * ROWS];              // this CTA's new outputs [utterance][ROWS]
    __shared__ __align__(8) unsigned long long mbar[2];

    const int B = a.B, T = a.T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rank = blockIdx.x;                 // == rank in cluster (cluster spans gridDim.x)
    const int dir = blockIdx.z;
    const int b0 = blockIdx.y * SC_NB;
    const float *W = a.W[dir];
    float *out = a.out[dir];
    const float *act = a.act[dir];
    const bool bptt = (a.mode == 1);
    const bool ascending = (dir == 0) != bptt;
    const int j0 = rank * ROWS + warp * RPW;

    float wreg[RPW][KI];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int i = 0; i < KI; ++i) {
            const int j = j0 + r, k = lane + 32 * i;
            wreg[r][i] = bptt ? W[(int64_t)k * H + j] : W[(int64_t)j * H + k];
        }

    const uint32_t bar0 = smem_u32(&mbar[0]), bar1 = smem_u32(&mbar[1]);
    if (threadIdx.x == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // arm both barriers for their first use (CS blocks of 1 KB each)
        mbar_arrive_expect_tx(bar0, CS * BLK_BYTES);
        mbar_arrive_expect_tx(bar1, CS * BLK_BYTES);
    }
    cluster_sync_all();      // barriers initialised cluster-wide before any peer pushes into them

    const int orow = lane / SC_NB, ob = lane % SC_NB;
    const int oj = j0 + orow, b = b0 + ob;
    const bool valid = (b < B);
    const int Tb = valid ? __ldg(a.Tlen + b) : 0;

    bool dead = false;
    for (int s = 0; s < T; ++s) {
        const int t = ascending ? s : T - 1 - s;
        float pre_v = 0.f, act_v = 0.f;
        if (valid) {
            const int64_t o = ((int64_t)t * B + b) * H + oj;
            pre_v = __ldg(a.pre + o);
            if (bptt) act_v = __ldg(act + o);
        }
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        if (s > 0) {
            // wait for the CS blocks of step s-1 (barrier s&1, use number (s-1)>>1)
            const uint32_t bar = (s & 1) ? bar1 : bar0;
            const uint32_t parity = (uint32_t)(((s - 1) >> 1) & 1);
            if ((lane == 0 || !(a.opt & 1)) && !dead && !mbar_try_wait(bar, parity)) {
                const long long t_start = clock64();
                while (!mbar_try_wait(bar, parity)) {
                    if (clock64() - t_start > 1000000000LL) {   // ~0.5 s: report, then run on without waiting
                        dead = true;
                        atomicExch(a.err, 2u);
                        break;
                    }
                }
            }
            __syncwarp();
            const float *hs = hbuf[s & 1];
#pragma unroll
            for (int i = 0; i < KI; ++i) {
                float hv[SC_NB];
#pragma unroll
                for (int bb = 0; bb < SC_NB; ++bb)
                    hv[bb] = hs[((32 * i) / ROWS * SC_NB + bb) * ROWS + (32 * i) % ROWS + lane];
#pragma unroll
                for (int r = 0; r < RPW; ++r)
#pragma unroll
                    for (int bb = 0; bb < SC_NB; ++bb)
                        acc[r * SC_NB + bb] = fmaf(wreg[r][i], hv[bb], acc[r * SC_NB + bb]);
            }
        }
        // transposing butterfly (unconditional: no collective inside a branch): lane l ends with acc[l] summed
#pragma unroll
        for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < n / 2; ++i) {
                const float send = up ? acc[i] : acc[i + n / 2];
                const float keep = up ? acc[i + n / 2] : acc[i];
                acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
        float v = 0.f;
        if (valid) {
            v = pre_v + acc[0];
            if (!bptt) v = fminf(fmaxf(v, 0.f), a.maxAct);                     // minmax(0, maxAct)
            else v = (act_v > 0.f && act_v < a.maxAct) ? v : 0.f;              // within(0, maxAct)
            if (t >= Tb) v = 0.f;
            if (!(a.opt & 2)) out[((int64_t)t * B + b) * H + oj] = v;
        }
        if (s + 1 < T) {
            stage[s & 1][ob * ROWS + warp * RPW + orow] = v;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic writes -> async proxy
            __syncthreads();   // block complete; every warp is done reading hbuf[s&1]
            if (warp == 0) {
                if (s > 0 && lane == 0) {
                    // re-arm the barrier we just consumed for its next use (step s+2)
                    mbar_arrive_expect_tx((s & 1) ? bar1 : bar0, CS * BLK_BYTES);
                }
                __syncwarp();
                if (lane < CS) {
                    const int nb = (s + 1) & 1;
                    const uint32_t dst = map_to_cta(smem_u32(&hbuf[nb][rank * SC_NB * ROWS]), (uint32_t)lane);
                    const uint32_t rbar = map_to_cta(nb ? bar1 : bar0, (uint32_t)lane);
                    bulk_push(dst, smem_u32(&stage[s & 1][0]), BLK_BYTES, rbar);
                }
            }
        }
        // stream the state to HBM for the GEMMs that follow -- after the push, so that the proxy fence
        // of the exchange never waits for this store
        if (valid && (a.opt & 2)) out[((int64_t)t * B + b) * H + oj] = v;
    }
    cluster_sync_all();      // no CTA may exit while peers can still address its shared memory
}

template <int KI, int RPW>
static int launch_cluster(const SweepClusterArgs &a, cudaStream_t st, bool *handled) {
    constexpr int CS = 4 * KI / RPW;
    constexpr int NB = 32 / RPW;
    const int ntiles = (a.B + NB - 1) / NB;
    if (ntiles > 65535) return CTCB_OK;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CS, ntiles, a.ndir);
    cfg.blockDim = dim3(SC_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (CS > 8)
        CTCB_CUDA_CHECK(cudaFuncSetAttribute((sweep_cluster_kernel<KI, RPW>), cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, (sweep_cluster_kernel<KI, RPW>), &cfg) != cudaSuccess || nclusters < 1) {
        cudaGetLastError();
        *handled = false;     // this device/partition cannot host the cluster: use the general kernel
        return CTCB_OK;
    }
    if (getenv("CTCB_DEBUG")) {
        static bool once = false;
        if (!once) { once = true; fprintf(stderr, "[ctcb] sweep v2 H=%d rpw=%d: cluster=%d, grid clusters=%d, max active clusters=%d\n", 32 * KI, RPW, CS, ntiles * a.ndir, nclusters); }
    }
    CTCB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, (sweep_cluster_kernel<KI, RPW>), a));
    count_launch();
    *handled = true;
    return CTCB_OK;
}

// ---------------------------------------------------------------------------------------------------
// v3: 16 warps per CTA (4 per scheduler, to hide the latency of the serial chain), 64 output units x 4
// utterances per CTA, packed FFMA2 over pairs of adjacent input units.  CS = H/64 CTAs per cluster.
//   block layout (1 KB per CTA and step): [utt pair (2)][unit pair (32)][{k,k+1} x {u,u+1}]
//   lane l of every warp owns input units 64*ip + 2l, 2l+1 of every slice ip: one slice per iteration,
//   two conflict-free LDS.128 feed 16 FFMA2.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float sum2(unsigned long long v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}
__device__ __forceinline__ void ffma2(unsigned long long &d, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

// NB = utterances per cluster (4..8).  B200 holds at most 15 clusters of 8 CTAs at once, so the launcher picks
// the smallest NB with 2*ceil(B/NB) <= that limit (B = 32 -> NB = 5, 14 clusters): a second wave of clusters
// would double the time of the whole sweep.
template <int KI, int NB>
__global__ void __launch_bounds__(512, 1) sweep_cluster_kernel_v3(SweepClusterArgs a) {
    constexpr int H = 32 * KI;
    constexpr int CS = H / 64;
    constexpr int NP = (NB + 1) / 2;                                     // utterance pairs
    constexpr int NV = 4 * NB;                                           // outputs per warp (<= 32)
    constexpr int BLKF = NP * 128;                                       // floats per CTA block
    constexpr uint32_t BLK_BYTES = BLKF * sizeof(float);
    __shared__ __align__(128) float hbuf[2][CS * BLKF];                  // [buffer][slice][utt pair][unit pair][4]
    __shared__ __align__(128) float stage[2][BLKF];
    __shared__ __align__(8) unsigned long long mbar[2];

    const int B = a.B, T = a.T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;        // 16 warps
    const int rank = blockIdx.x;
    const int dir = blockIdx.z;
    const int b0 = blockIdx.y * NB;
    const float *W = a.W[dir];
    float *out = a.out[dir];
    const float *act = a.act[dir];
    const bool bptt = (a.mode == 1);
    const bool ascending = (dir == 0) != bptt;
    const int j0 = rank * 64 + warp * 4;

    unsigned long long w2[4][CS];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int ip = 0; ip < CS; ++ip) {
            const int j = j0 + r, k = 64 * ip + 2 * lane;
            w2[r][ip] = bptt ? pack2(W[(int64_t)k * H + j], W[(int64_t)(k + 1) * H + j])
                             : pack2(W[(int64_t)j * H + k], W[(int64_t)j * H + k + 1]);
        }
    // the padding utterance of an odd NB must read as zero in every buffer
    for (int i = threadIdx.x; i < 2 * BLKF; i += 512) (&stage[0][0])[i] = 0.f;

    const uint32_t bar0 = smem_u32(&mbar[0]), bar1 = smem_u32(&mbar[1]);
    if (threadIdx.x == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arrive_expect_tx(bar0, CS * BLK_BYTES);
        mbar_arrive_expect_tx(bar1, CS * BLK_BYTES);
    }
    cluster_sync_all();

    // after the reduction lane l owns output index l = orow*NB + ob (lanes >= NV idle)
    const int orow = lane / NB, ob = lane % NB;
    const int oj = j0 + orow, b = b0 + ob;
    const bool valid = (lane < NV) && (b < B);
    const int Tb = valid ? __ldg(a.Tlen + b) : 0;
    bool dead = false;

    for (int s = 0; s < T; ++s) {
        const int t = ascending ? s : T - 1 - s;
        float pre_v = 0.f, act_v = 0.f;
        if (valid) {
            const int64_t o = ((int64_t)t * B + b) * H + oj;
            pre_v = __ldg(a.pre + o);
            if (bptt) act_v = __ldg(act + o);
        }
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        if (s > 0) {
            const uint32_t bar = (s & 1) ? bar1 : bar0;
            const uint32_t parity = (uint32_t)(((s - 1) >> 1) & 1);
            if (!dead && !mbar_try_wait(bar, parity)) {
                const long long t_start = clock64();
                while (!mbar_try_wait(bar, parity)) {
                    if (clock64() - t_start > 1000000000LL) {   // ~0.5 s: report, then run on without waiting
                        dead = true;
                        atomicExch(a.err, 2u);
                        break;
                    }
                }
            }
            const ulonglong2 *hs = reinterpret_cast<const ulonglong2 *>(hbuf[s & 1]);
            unsigned long long acc2[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) acc2[i] = 0ull;
#pragma unroll
            for (int ip = 0; ip < CS; ++ip) {
                ulonglong2 hv[NP];
#pragma unroll
                for (int q = 0; q < NP; ++q) hv[q] = hs[ip * (BLKF / 4) + q * 32 + lane];   // {k,k+1} x utterances 2q,2q+1
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int u = 0; u < NB; ++u)
                        ffma2(acc2[r * NB + u], w2[r][ip], (u & 1) ? hv[u >> 1].y : hv[u >> 1].x);
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] = sum2(acc2[i]);
        }
        // transposing butterfly over 32 slots (slots >= NV are zero): lane l ends with the full sum of slot l
#pragma unroll
        for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < n / 2; ++i) {
                const float send = up ? acc[i] : acc[i + n / 2];
                const float keep = up ? acc[i + n / 2] : acc[i];
                acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
        float v = 0.f;
        if (valid) {
            v = pre_v + acc[0];
            if (!bptt) v = fminf(fmaxf(v, 0.f), a.maxAct);
            else v = (act_v > 0.f && act_v < a.maxAct) ? v : 0.f;
            if (t >= Tb) v = 0.f;
            out[((int64_t)t * B + b) * H + oj] = v;
        }
        if (s + 1 < T) {
            if (lane < NV) {
                const int jl = warp * 4 + orow;     // local unit 0..63
                stage[s & 1][(ob >> 1) * 128 + (jl >> 1) * 4 + (ob & 1) * 2 + (jl & 1)] = v;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (warp == 0) {
                if (s > 0 && lane == 0) mbar_arrive_expect_tx((s & 1) ? bar1 : bar0, CS * BLK_BYTES);
                __syncwarp();
                if (lane < CS) {
                    const int nb = (s + 1) & 1;
                    const uint32_t dst = map_to_cta(smem_u32(&hbuf[nb][rank * BLKF]), (uint32_t)lane);
                    const uint32_t rbar = map_to_cta(nb ? bar1 : bar0, (uint32_t)lane);
                    bulk_push(dst, smem_u32(&stage[s & 1][0]), BLK_BYTES, rbar);
                }
            }
        }
    }
    cluster_sync_all();
}

template <int KI, int NB>
static int launch_cluster_v3_nb(const SweepClusterArgs &a, int ntiles, cudaStream_t st) {
    constexpr int CS = KI / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CS, ntiles, a.ndir);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CTCB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, (sweep_cluster_kernel_v3<KI, NB>), a));
    count_launch();
    return CTCB_OK;
}

template <int KI>
static int launch_cluster_v3(const SweepClusterArgs &a, cudaStream_t st, bool *handled) {
    constexpr int CS = KI / 2;
    // how many clusters of this shape the device holds at once (shape-independent of NB up to shared memory)
    static int max_clusters = -1;
    if (max_clusters < 0) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(CS, 64, 2);
        cfg.blockDim = dim3(512);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, (sweep_cluster_kernel_v3<KI, 8>), &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
        max_clusters = n;
        if (getenv("CTCB_DEBUG")) fprintf(stderr, "[ctcb] sweep v3 H=%d: cluster=%d CTAs, max active clusters=%d\n", 32 * KI, CS, n);
    }
    *handled = false;
    if (max_clusters < 2) return CTCB_OK;
    static int nb_env = -1;   // CTCB_SWEEP_NB=4..8 forces the group size
    if (nb_env < 0) { const char *e = getenv("CTCB_SWEEP_NB"); nb_env = e ? atoi(e) : 0; }
    int nb = 0;
    for (int c = 4; c <= 8; ++c)
        if (a.ndir * ((a.B + c - 1) / c) <= max_clusters) { nb = c; break; }
    if (nb_env >= 4 && nb_env <= 8) nb = nb_env;
    if (nb == 0) return CTCB_OK;          // would need a second wave of clusters: the general kernel is faster
    const int ntiles = (a.B + nb - 1) / nb;
    int rc;
    switch (nb) {
        case 4: rc = launch_cluster_v3_nb<KI, 4>(a, ntiles, st); break;
        case 5: rc = launch_cluster_v3_nb<KI, 5>(a, ntiles, st); break;
        case 6: rc = launch_cluster_v3_nb<KI, 6>(a, ntiles, st); break;
        case 7: rc = launch_cluster_v3_nb<KI, 7>(a, ntiles, st); break;
        default: rc = launch_cluster_v3_nb<KI, 8>(a, ntiles, st); break;
    }
    if (rc == CTCB_OK) *handled = true;
    return rc;
}

int run_sweep_cluster(int mode, int T, int B, int H, const int32_t *Tlen, const float *pre, const float *Wf,
                      const float *Wb, float *outF, float *outB, const float *actF, const float *actB, float maxAct,
                      unsigned int *err, cudaStream_t st, bool *handled) {
    *handled = false;
    if (H != 128 && H != 256 && H != 512) return CTCB_OK;
    static int force_cluster = -1;   // CTCB_SWEEP=cluster takes this kernel for H = 512 too
    if (force_cluster < 0) { const char *e = getenv("CTCB_SWEEP"); force_cluster = (e && e[0] == 'c') ? 1 : 0; }
    (void)force_cluster;
    SweepClusterArgs a;
    a.mode = mode; a.T = T; a.B = B; a.H = H; a.Tlen = Tlen; a.pre = pre;
    a.W[0] = Wf; a.W[1] = Wb; a.out[0] = outF; a.out[1] = outB; a.act[0] = actF; a.act[1] = actB; a.maxAct = maxAct; a.err = err;
    a.ndir = Wb ? 2 : 1;
    if (!Wb) { a.W[1] = Wf; a.out[1] = outF; a.act[1] = actF; }
    {
        static int opt = -1;
        if (opt < 0) { const char *e = getenv("CTCB_SWEEP_OPT"); opt = e ? atoi(e) : 0; }
        a.opt = opt;
    }
    static int ver_env = -1;   // CTCB_SWEEP_V=2 selects the 8-warp kernels below instead of v3
    if (ver_env < 0) { const char *e = getenv("CTCB_SWEEP_V"); ver_env = e ? atoi(e) : 0; }
    // measured (tools/sweep_time.py, B=32, T=200): H=512 v3 0.46 ms vs v2 0.69 vs barrier 0.63; H=256 v2 0.23 vs v3 0.27
    if (ver_env == 3 || (ver_env == 0 && H == 512)) {
        switch (H / 32) {
            case 4: return launch_cluster_v3<4>(a, st, handled);
            case 8: return launch_cluster_v3<8>(a, st, handled);
            case 16: return launch_cluster_v3<16>(a, st, handled);
            default: return CTCB_OK;
        }
    }
    static int rpw_env = -1;   // CTCB_SWEEP_RPW=4|8 overrides the default shape
    if (rpw_env < 0) { const char *e = getenv("CTCB_SWEEP_RPW"); rpw_env = e ? atoi(e) : 0; }
    // defaults from measurements on B200 (tools/sweep_time.py): clusters of 8 CTAs are the sweet spot
    const int rpw = rpw_env ? rpw_env : (H >= 256 ? 8 : 4);
    switch (H / 32) {
        case 4: return launch_cluster<4, 4>(a, st, handled);
        case 8: return rpw == 8 ? launch_cluster<8, 8>(a, st, handled) : launch_cluster<8, 4>(a, st, handled);
        case 16: return rpw == 8 ? launch_cluster<16, 8>(a, st, handled) : launch_cluster<16, 4>(a, st, handled);
        default: return CTCB_OK;
    }
}

}  // namespace ctcb
