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
    cfg.gridDim = dim3(CS, ntiles, 2);
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
        if (!once) { once = true; fprintf(stderr, "[ctcb] sweep v2 H=%d rpw=%d: cluster=%d, grid clusters=%d, max active clusters=%d\n", 32 * KI, RPW, CS, ntiles * 2, nclusters); }
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

// HALVES = 1: one group of 4 utterances per cluster.  HALVES = 2: two groups that share the weight
// registers and are processed back to back in every step, each with its own state buffers and mbarriers,
// so the exchange of one group overlaps the arithmetic of the other (used when the grid would otherwise
// need more clusters than the device can hold at once: 15 clusters of 8 on B200).
template <int KI, int HALVES>
__global__ void __launch_bounds__(512, 1) sweep_cluster_kernel_v3(SweepClusterArgs a) {
    constexpr int H = 32 * KI;
    constexpr int CS = H / 64;
    constexpr int NB = 4;
    constexpr uint32_t BLK_BYTES = 64 * NB * sizeof(float);              // 1 KB
    __shared__ __align__(128) float hbuf[HALVES][2][CS * 256];
    __shared__ __align__(128) float stage[HALVES][2][256];
    __shared__ __align__(8) unsigned long long mbar[HALVES][2];

    const int B = a.B, T = a.T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;        // 16 warps
    const int rank = blockIdx.x;
    const int dir = blockIdx.z;
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

    if (threadIdx.x == 0) {
#pragma unroll
        for (int hf = 0; hf < HALVES; ++hf)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                mbar_init(smem_u32(&mbar[hf][q]), 1);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int hf = 0; hf < HALVES; ++hf)
#pragma unroll
            for (int q = 0; q < 2; ++q) mbar_arrive_expect_tx(smem_u32(&mbar[hf][q]), CS * BLK_BYTES);
    }
    cluster_sync_all();

    // after the reduction lane l (and l^16) owns output index l & 15 = orow*4 + ob; lanes < 16 write it
    const int oidx = lane & 15, orow = oidx >> 2, ob = oidx & 3;
    const int oj = j0 + orow;
    int bq[HALVES], Tbq[HALVES];
    bool validq[HALVES];
#pragma unroll
    for (int hf = 0; hf < HALVES; ++hf) {
        bq[hf] = (blockIdx.y * HALVES + hf) * NB + ob;
        validq[hf] = (bq[hf] < B) && (lane < 16);
        Tbq[hf] = (bq[hf] < B) ? __ldg(a.Tlen + bq[hf]) : 0;
    }
    bool dead = false;

    for (int s = 0; s < T; ++s) {
        const int t = ascending ? s : T - 1 - s;
#pragma unroll
        for (int hf = 0; hf < HALVES; ++hf) {
            const int b = bq[hf];
            const bool valid = validq[hf];
            float pre_v = 0.f, act_v = 0.f;
            if (valid) {
                const int64_t o = ((int64_t)t * B + b) * H + oj;
                pre_v = __ldg(a.pre + o);
                if (bptt) act_v = __ldg(act + o);
            }
            float acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
            if (s > 0) {
                const uint32_t bar = smem_u32(&mbar[hf][s & 1]);
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
                const ulonglong2 *hs = reinterpret_cast<const ulonglong2 *>(hbuf[hf][s & 1]);
                unsigned long long acc2[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) acc2[i] = 0ull;
#pragma unroll
                for (int ip = 0; ip < CS; ++ip) {
                    const ulonglong2 h01 = hs[ip * 64 + lane];          // {k,k+1} x utterances 0,1
                    const ulonglong2 h23 = hs[ip * 64 + 32 + lane];     // {k,k+1} x utterances 2,3
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        ffma2(acc2[r * 4 + 0], w2[r][ip], h01.x);
                        ffma2(acc2[r * 4 + 1], w2[r][ip], h01.y);
                        ffma2(acc2[r * 4 + 2], w2[r][ip], h23.x);
                        ffma2(acc2[r * 4 + 3], w2[r][ip], h23.y);
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = sum2(acc2[i]);
            }
            // 16 values x 32 lanes -> lane l holds the full sum of value l & 15
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
#pragma unroll
            for (int off = 8, n = 16; off >= 1; off >>= 1, n >>= 1) {
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
                if (t >= Tbq[hf]) v = 0.f;
                out[((int64_t)t * B + b) * H + oj] = v;
            }
            if (s + 1 < T) {
                if (lane < 16) {
                    const int jl = warp * 4 + orow;     // local unit 0..63
                    stage[hf][s & 1][(ob >> 1) * 128 + (jl >> 1) * 4 + (ob & 1) * 2 + (jl & 1)] = v;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                if (warp == 0) {
                    if (s > 0 && lane == 0) mbar_arrive_expect_tx(smem_u32(&mbar[hf][s & 1]), CS * BLK_BYTES);
                    __syncwarp();
                    if (lane < CS) {
                        const int nb = (s + 1) & 1;
                        const uint32_t dst = map_to_cta(smem_u32(&hbuf[hf][nb][rank * 256]), (uint32_t)lane);
                        const uint32_t rbar = map_to_cta(smem_u32(&mbar[hf][nb]), (uint32_t)lane);
                        bulk_push(dst, smem_u32(&stage[hf][s & 1][0]), BLK_BYTES, rbar);
                    }
                }
            }
        }
    }
    cluster_sync_all();
}

template <int KI, int HALVES>
static int launch_cluster_v3(const SweepClusterArgs &a, cudaStream_t st, bool *handled) {
    constexpr int CS = KI / 2;
    const int ntiles = (a.B + 4 * HALVES - 1) / (4 * HALVES);
    if (ntiles > 65535) return CTCB_OK;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CS, ntiles, 2);
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
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, (sweep_cluster_kernel_v3<KI, HALVES>), &cfg) != cudaSuccess || nclusters < 1) {
        cudaGetLastError();
        *handled = false;
        return CTCB_OK;
    }
    if (HALVES == 1 && ntiles * 2 > nclusters) {
        // more clusters than the device holds at once would run in two waves: let every cluster carry two
        // utterance groups instead (their exchange/compute phases interleave)
        return launch_cluster_v3<KI, 2>(a, st, handled);
    }
    if (getenv("CTCB_DEBUG")) {
        static bool once = false;
        if (!once) { once = true; fprintf(stderr, "[ctcb] sweep v3 H=%d halves=%d: cluster=%d, grid clusters=%d, max active clusters=%d\n", 32 * KI, HALVES, CS, ntiles * 2, nclusters); }
    }
    CTCB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, (sweep_cluster_kernel_v3<KI, HALVES>), a));
    count_launch();
    *handled = true;
    return CTCB_OK;
}

int run_sweep_cluster(int mode, int T, int B, int H, const int32_t *Tlen, const float *pre, const float *Wf,
                      const float *Wb, float *outF, float *outB, const float *actF, const float *actB, float maxAct,
                      unsigned int *err, cudaStream_t st, bool *handled) {
    *handled = false;
    if (H != 128 && H != 256 && H != 512) return CTCB_OK;
    static int force_cluster = -1;   // CTCB_SWEEP=cluster takes this kernel for H = 512 too
    if (force_cluster < 0) { const char *e = getenv("CTCB_SWEEP"); force_cluster = (e && e[0] == 'c') ? 1 : 0; }
    // measured on B200 (tools/sweep_time.py, bench.py): clusters win 2x at H <= 256 (1.2 vs 2.6 us/step);
    // at H = 512 the counter-barrier kernel is still ~5% ahead (3.2 vs 3.3 us/step), so it stays the default there
    if (H == 512 && !force_cluster) return CTCB_OK;
    SweepClusterArgs a;
    a.mode = mode; a.T = T; a.B = B; a.H = H; a.Tlen = Tlen; a.pre = pre;
    a.W[0] = Wf; a.W[1] = Wb; a.out[0] = outF; a.out[1] = outB; a.act[0] = actF; a.act[1] = actB; a.maxAct = maxAct; a.err = err;
    {
        static int opt = -1;
        if (opt < 0) { const char *e = getenv("CTCB_SWEEP_OPT"); opt = e ? atoi(e) : 0; }
        a.opt = opt;
    }
    static int ver_env = -1;   // CTCB_SWEEP_V=2 selects the 8-warp kernels below instead of v3
    if (ver_env < 0) { const char *e = getenv("CTCB_SWEEP_V"); ver_env = e ? atoi(e) : 3; }
    if (ver_env == 3) {
        switch (H / 32) {
            case 4: return launch_cluster_v3<4, 1>(a, st, handled);
            case 8: return launch_cluster_v3<8, 1>(a, st, handled);
            case 16: return launch_cluster_v3<16, 1>(a, st, handled);
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
