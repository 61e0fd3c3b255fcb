// sweep_cluster.cu -- the temporal-layer recurrences on thread-block clusters (H = 128, 256, 512).
//
// Same arithmetic as sweep.cu (reference loops brnnet.py:144-152 forward, :208-224 BPTT), but the
// hidden state never touches global memory on the serial chain:
//
//   * one CLUSTER of CS = H/32 CTAs per (direction, tile of 8 utterances); CTA r owns 32 output units
//     and keeps its 32 x H slice of the recurrent matrix in registers for the whole sweep;
//   * every CTA holds the complete previous state of its 8 utterances in shared memory, laid out
//     [slice][utterance][32] so that one CTA's contribution is one contiguous 1 KB block;
//   * after a step, each CTA pushes its 1 KB block into the shared memory of all CS CTAs of the
//     cluster with cp.async.bulk (shared::cta -> shared::cluster); the copies complete_tx on the
//     DESTINATION's mbarrier, which is the only synchronisation of the step: a CTA starts step s as
//     soon as the CS blocks of step s-1 have landed.  Two state buffers / two barriers alternate;
//   * For/Back (dFor/dBack) are still streamed to HBM for the GEMMs that follow, off the chain.
#include "common.cuh"

namespace ctcb {

constexpr int SC_THREADS = 256;
constexpr int SC_NB = 8;

struct SweepClusterArgs {
    int mode, T, B, H;
    const int32_t *Tlen;
    const float *pre;
    const float *W[2];
    float *out[2];
    const float *act[2];
    float maxAct;
    unsigned int *err;      // [0] set to 2 if a barrier wait timed out (never a hang)
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

// KI = H/32 = cluster size.  grid = (KI, ntiles, 2), cluster = (KI, 1, 1).
template <int KI>
__global__ void __launch_bounds__(SC_THREADS, 1) sweep_cluster_kernel(SweepClusterArgs a) {
    constexpr int H = 32 * KI;
    constexpr uint32_t BLK_BYTES = SC_NB * 32 * sizeof(float);          // one CTA's block: 1 KB
    __shared__ __align__(128) float hbuf[2][KI * SC_NB * 32];           // [buffer][slice][utterance][32]
    __shared__ __align__(128) float stage[2][SC_NB * 32];               // this CTA's new outputs [utterance][32]
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
    const int j0 = rank * 32 + warp * 4;

    float wreg[4][KI];
#pragma unroll
    for (int r = 0; r < 4; ++r)
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
        mbar_arrive_expect_tx(bar0, KI * BLK_BYTES);
        mbar_arrive_expect_tx(bar1, KI * BLK_BYTES);
    }
    cluster_sync_all();      // barriers initialised cluster-wide before any peer pushes into them

    const int orow = lane >> 3, ob = lane & 7;
    const int oj = j0 + orow, b = b0 + ob;
    const bool valid = (b < B);
    const int Tb = valid ? __ldg(a.Tlen + b) : 0;

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
            if (!mbar_try_wait(bar, parity)) {
                const long long t_start = clock64();
                bool dead = false;
                while (!mbar_try_wait(bar, parity)) {
                    if (clock64() - t_start > 1000000000LL) { dead = true; break; }   // ~0.5 s
                }
                if (dead) { atomicExch(a.err, 2u); break; }
            }
            const float *hs = hbuf[s & 1];
#pragma unroll
            for (int i = 0; i < KI; ++i) {
                float hv[SC_NB];
#pragma unroll
                for (int bb = 0; bb < SC_NB; ++bb) hv[bb] = hs[(i * SC_NB + bb) * 32 + lane];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int bb = 0; bb < SC_NB; ++bb) acc[r * 8 + bb] = fmaf(wreg[r][i], hv[bb], acc[r * 8 + bb]);
            }
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
        }
        float v = 0.f;
        if (valid) {
            v = pre_v + acc[0];
            if (!bptt) v = fminf(fmaxf(v, 0.f), a.maxAct);                     // minmax(0, maxAct)
            else v = (act_v > 0.f && act_v < a.maxAct) ? v : 0.f;              // within(0, maxAct)
            if (t >= Tb) v = 0.f;
            out[((int64_t)t * B + b) * H + oj] = v;
        }
        if (s + 1 < T) {
            stage[s & 1][ob * 32 + warp * 4 + orow] = v;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic writes -> async proxy
            __syncthreads();   // block complete; every warp is done reading hbuf[s&1]
            if (warp == 0) {
                if (s > 0 && lane == 0) {
                    // re-arm the barrier we just consumed for its next use (step s+2)
                    mbar_arrive_expect_tx((s & 1) ? bar1 : bar0, KI * BLK_BYTES);
                }
                __syncwarp();
                if (lane < KI) {
                    const int nb = (s + 1) & 1;
                    const uint32_t dst = map_to_cta(smem_u32(&hbuf[nb][rank * SC_NB * 32]), (uint32_t)lane);
                    const uint32_t rbar = map_to_cta(nb ? bar1 : bar0, (uint32_t)lane);
                    bulk_push(dst, smem_u32(&stage[s & 1][0]), BLK_BYTES, rbar);
                }
            }
        }
    }
    cluster_sync_all();      // no CTA may exit while peers can still address its shared memory
}

template <int KI>
static int launch_cluster(const SweepClusterArgs &a, int ntiles, cudaStream_t st, bool *handled) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(KI, ntiles, 2);
    cfg.blockDim = dim3(SC_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = KI;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (KI > 8)
        CTCB_CUDA_CHECK(cudaFuncSetAttribute(sweep_cluster_kernel<KI>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, sweep_cluster_kernel<KI>, &cfg) != cudaSuccess || nclusters < 1) {
        cudaGetLastError();
        *handled = false;     // this device/partition cannot host the cluster: use the general kernel
        return CTCB_OK;
    }
    CTCB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, sweep_cluster_kernel<KI>, a));
    count_launch();
    *handled = true;
    return CTCB_OK;
}

int run_sweep_cluster(int mode, int T, int B, int H, const int32_t *Tlen, const float *pre, const float *Wf,
                      const float *Wb, float *outF, float *outB, const float *actF, const float *actB, float maxAct,
                      unsigned int *err, cudaStream_t st, bool *handled) {
    *handled = false;
    if (H != 128 && H != 256 && H != 512) return CTCB_OK;
    SweepClusterArgs a;
    a.mode = mode; a.T = T; a.B = B; a.H = H; a.Tlen = Tlen; a.pre = pre;
    a.W[0] = Wf; a.W[1] = Wb; a.out[0] = outF; a.out[1] = outB; a.act[0] = actF; a.act[1] = actB; a.maxAct = maxAct; a.err = err;
    const int ntiles = (B + SC_NB - 1) / SC_NB;
    if (ntiles > 65535) return CTCB_OK;
    switch (H / 32) {
        case 4: return launch_cluster<4>(a, ntiles, st, handled);
        case 8: return launch_cluster<8>(a, ntiles, st, handled);
        case 16: return launch_cluster<16>(a, ntiles, st, handled);
        default: return CTCB_OK;
    }
}

}  // namespace ctcb
