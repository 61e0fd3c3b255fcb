// sweep.cu -- the bi-directional time recurrences of the temporal layer as ONE persistent kernel.
//
// Replaces the 4(T-1) cudamat launches per utterance of the reference's hot loops
//   forward : For[:,t]  = clip(pre[:,t] + Wtf.For[:,t-1], 0, 20),  Back mirrored
//             (/root/reference/ctc_fast/nnets/brnnet.py:144-152: mvdot_col_slice + minmax)
//   BPTT    : dFor[:,t] = within(For[:,t]) * (d[:,t] + Wtf^T.dFor[:,t+1]),  dBack mirrored
//             (brnnet.py:208-224: mvdot_col_slice on W.T + mult_slice)
// with a batch of utterances as the second matrix dimension.
//
// Design (B200): the H x H recurrent matrix never leaves the register file.  A CTA of 8 warps owns
// 32 output units; lane l of warp w keeps W[j][l + 32 i] for its 4 rows j in registers for the
// whole sweep (H/32 * 4 registers), so a time step only moves the previous hidden state
// (8 utterances x H floats, read from L2 into shared memory) and the 32 x 8 new outputs.  Both
// directions run concurrently in the same launch (blockIdx.z).
//
// There is NO barrier between time steps.  The hidden state is exchanged through the output array
// itself (it has to be written anyway) with a flag-in-data protocol: the producers pre-fill their own
// words with a sentinel bit pattern (0xffffffff, a NaN no arithmetic produces; one grid barrier at
// kernel start orders the fills), then overwrite them with
// st.relaxed.gpu, and consumers poll their operand loads (ld.acquire.gpu) until no word is the sentinel.
// One L2 round trip per step replaces fence + atomic + poll + load of a counter barrier, and CTAs run
// as a decoupled dataflow pipeline (cooperative launch guarantees the co-residency polling needs).
//
// Data layout: time-major [T][B][H] fp32, so one time step of all utterances is contiguous.
#include "common.cuh"
#include <cooperative_groups.h>

namespace ctcb {

constexpr int SW_THREADS = 256;
constexpr int SW_ROWS = 32;   // output units per CTA (4 per warp)
constexpr int SW_NB = 8;      // utterances per inner tile

struct SweepArgs {
    int mode;               // 0: forward recurrences, 1: BPTT
    int T, B, H;
    const int32_t *Tlen;    // [B]
    const float *pre;       // mode 0: pre-activations [T][B][H]; mode 1: incoming deltas
    const float *W[2];      // Wtf, Wtb  (H x H, row-major out x in)
    float *out[2];          // mode 0: For, Back; mode 1: dFor, dBack
    const float *act[2];    // mode 1: For, Back (for the within(0,maxAct) masks)
    float maxAct;
    int parts;              // utterance partitions (gridDim.y)
    unsigned int *counters; // [0]: sticky poll-timeout flag, zeroed before launch
};

constexpr unsigned SENTINEL = 0xffffffffu;
constexpr long long POLL_CYCLES = 1000000000LL;   // ~0.5 s of polling: a lost producer becomes an error flag, not a hang

__device__ __forceinline__ float4 ld_relaxed4(const float *p) {
    float4 v;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
// Re-polls must be ACQUIRE loads: a relaxed gpu-scope load can keep hitting the SM's stale L1 copy of
// a line it fetched while the producer was still writing it (measured on B200: the poll never sees the
// update); the acquire's CCTL.IVALL drops that copy.  A stale copy only ever holds sentinel-or-final
// words, so a word that does not read as the sentinel is always the final value.
__device__ __forceinline__ float4 ld_volatile4(const float *p) {
    float4 v;
    asm volatile("ld.acquire.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_volatile1(const float *p) {
    float v;
    asm volatile("ld.acquire.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
// Strong (gpu-scope, relaxed) store: a weak st.global may linger in the SM and race with the polls.
__device__ __forceinline__ void st_relaxed_gpu(float *p, float v) {
    asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ bool is_sentinel(float x) { return __float_as_uint(x) == SENTINEL; }
__device__ __forceinline__ bool any_sentinel(const float4 &v) {
    return is_sentinel(v.x) || is_sentinel(v.y) || is_sentinel(v.z) || is_sentinel(v.w);
}

// KI > 0: H == 32*KI and the weights live in registers.  KI == 0: any H, weights re-read through L1/L2.
template <int KI>
__global__ void __launch_bounds__(SW_THREADS, 1) sweep_kernel(SweepArgs a) {
    extern __shared__ __align__(16) float hs[];   // [SW_NB][Hp]
    const int H = a.H, B = a.B, T = a.T;
    const int Hp = (H + 3) / 4 * 4 + 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dir = blockIdx.z;
    const int j0 = blockIdx.x * SW_ROWS + warp * 4;
    const float *W = a.W[dir];
    float *out = a.out[dir];
    const bool bptt = (a.mode == 1);
    // time runs forward for (forward-mode, dir 0) and (BPTT, dir 1)
    const bool ascending = (dir == 0) != bptt;
    // contraction: forward uses W (row j), BPTT uses W^T (column j)
    float wreg[4][KI > 0 ? KI : 1];
    if (KI > 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < KI; ++i) {
                const int j = j0 + r, k = lane + 32 * i;
                wreg[r][i] = (j < H) ? (bptt ? W[(int64_t)k * H + j] : W[(int64_t)j * H + k]) : 0.f;
            }
    }
    // utterance range of this partition, in tiles of SW_NB
    const int ntiles = (B + SW_NB - 1) / SW_NB;
    const int tiles_per_part = (ntiles + a.parts - 1) / a.parts;
    const int tile_beg = blockIdx.y * tiles_per_part;
    const int tile_end = min(ntiles, tile_beg + tiles_per_part);

    // lane -> (row r, utterance b) of the value it ends up owning after the transposing reduction
    const int orow = lane >> 3, ob = lane & 7;
    const int oj = j0 + orow;

    // Sentinel pre-fill by the producers themselves: every lane marks the words it will produce later
    // as "not produced yet", then ONE grid-wide barrier (the only one of the sweep) orders all fills
    // before any poll.  Same-thread program order then guarantees fill-before-produce.
    for (int tile = tile_beg; tile < tile_end; ++tile) {
        const int b = tile * SW_NB + ob;
        if (oj < H && b < B)
            for (int t = 0; t < T; ++t) st_relaxed_gpu(out + ((int64_t)t * B + b) * H + oj, __uint_as_float(SENTINEL));
    }
    __threadfence();
    cooperative_groups::this_grid().sync();

    for (int s = 0; s < T; ++s) {
        const int t = ascending ? s : T - 1 - s;
        const int tprev = ascending ? t - 1 : t + 1;
        for (int tile = tile_beg; tile < tile_end; ++tile) {
            const int b0 = tile * SW_NB;
            const int b = b0 + ob;
            // issue the epilogue operands early: they do not depend on the recurrence
            float pre_v = 0.f, act_v = 0.f;
            int Tb = 0;
            const bool valid = (oj < H) && (b < B);
            if (valid) {
                const int64_t o = ((int64_t)t * B + b) * H + oj;
                pre_v = __ldg(a.pre + o);
                if (bptt) act_v = __ldg(a.act[dir] + o);
                Tb = __ldg(a.Tlen + b);
            }
            float acc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0.f;
            if (s > 0) {
                // previous state of this tile's utterances: [SW_NB][H] from L2 -> shared
                const float *src = out + ((int64_t)tprev * B + b0) * H;
                const int nb = min(SW_NB, B - b0);
                if (KI >= 4) {
                    // all operand loads in flight at once, then re-poll only the words still unwritten
                    constexpr int NCH = (KI >= 4) ? KI / 4 : 1;     // float4 chunks per thread: 8*H/4/256
                    const int h4 = H >> 2;
                    float4 v[NCH];
                    const float *p[NCH];
                    bool need[NCH];
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const int idx = threadIdx.x + c * SW_THREADS;
                        const int bb = idx / h4, k4 = idx - bb * h4;
                        need[c] = bb < nb;
                        p[c] = src + (int64_t)bb * H + 4 * k4;
                        v[c] = need[c] ? ld_relaxed4(p[c]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    int spins = 0;
                    bool dead = false;
                    long long t_start = 0;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        while (need[c] && !dead && any_sentinel(v[c])) {   // producer has not stored this word yet
                            if (spins == 0) t_start = clock64();
                            if (((++spins & 63) == 0) &&
                                (clock64() - t_start > POLL_CYCLES || *(volatile unsigned int *)a.counters != 0u)) {
                                if (atomicCAS(a.counters, 0u, 1u) == 0u) {   // first timeout: who waited for what
                                    a.counters[3] = (unsigned)s; a.counters[4] = blockIdx.x; a.counters[5] = blockIdx.y;
                                    a.counters[6] = blockIdx.z; a.counters[7] = (unsigned)c; a.counters[8] = threadIdx.x;
                                    a.counters[9] = (unsigned)tprev;
                                }
                                atomicAdd(a.counters + 2, 1u);
                                if (threadIdx.x == 0 || a.counters[16 + (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] == 0u)
                                    a.counters[16 + (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = 1000u + (unsigned)s;
                                dead = true;
                                break;
                            }
                            v[c] = ld_volatile4(p[c]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const int idx = threadIdx.x + c * SW_THREADS;
                        const int bb = idx / h4, k4 = idx - bb * h4;
                        if (need[c]) *reinterpret_cast<float4 *>(hs + bb * Hp + 4 * k4) = v[c];
                    }
                } else {
                    for (int idx = threadIdx.x; idx < nb * H; idx += SW_THREADS) {
                        const int bb = idx / H, k = idx - bb * H;
                        const float *p = src + (int64_t)bb * H + k;
                        float v = ld_volatile1(p);
                        int spins = 0;
                        long long t_start = 0;
                        while (is_sentinel(v)) {
                            if (spins == 0) t_start = clock64();
                            if (((++spins & 63) == 0) &&
                                (clock64() - t_start > POLL_CYCLES || *(volatile unsigned int *)a.counters != 0u)) {
                                atomicExch(a.counters, 1u);
                                break;
                            }
                            v = ld_volatile1(p);
                        }
                        hs[bb * Hp + k] = v;
                    }
                }
                for (int idx = threadIdx.x + nb * Hp; idx < SW_NB * Hp; idx += SW_THREADS) hs[idx] = 0.f;
                __syncthreads();
                if (KI > 0) {
#pragma unroll
                    for (int i = 0; i < KI; ++i) {
                        float hv[SW_NB];
#pragma unroll
                        for (int bb = 0; bb < SW_NB; ++bb) hv[bb] = hs[bb * Hp + lane + 32 * i];
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int bb = 0; bb < SW_NB; ++bb)
                                acc[r * 8 + bb] = fmaf(wreg[r][i], hv[bb], acc[r * 8 + bb]);
                    }
                } else {
                    for (int k = lane; k < H; k += 32) {
                        float hv[SW_NB], wv[4];
#pragma unroll
                        for (int bb = 0; bb < SW_NB; ++bb) hv[bb] = hs[bb * Hp + k];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int j = j0 + r;
                            wv[r] = (j < H) ? (bptt ? __ldg(W + (int64_t)k * H + j) : __ldg(W + (int64_t)j * H + k)) : 0.f;
                        }
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int bb = 0; bb < SW_NB; ++bb)
                                acc[r * 8 + bb] = fmaf(wv[r], hv[bb], acc[r * 8 + bb]);
                    }
                }
                // transposing butterfly: lane l ends with sum over lanes of acc[l]
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
                __syncthreads();   // hs is reused by the next tile
            }
            if (valid) {
                float v = pre_v + acc[0];
                if (!bptt) {
                    v = fminf(fmaxf(v, 0.f), a.maxAct);          // minmax(0, maxAct)
                    if (t >= Tb) v = 0.f;                        // beyond the utterance: zero state
                } else {
                    v = (act_v > 0.f && act_v < a.maxAct) ? v : 0.f;   // within(0, maxAct)
                    if (t >= Tb) v = 0.f;
                }
                if (is_sentinel(v)) v = __uint_as_float(0x7fc00000u);   // a NaN result must not look unwritten
                st_relaxed_gpu(out + ((int64_t)t * B + b) * H + oj, v);
            }
        }
    }
}

__global__ void add2_kernel(const float *__restrict__ x, const float *__restrict__ y, float *__restrict__ z, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) z[i] = x[i] + y[i];
}

template <int KI>
static int launch_sweep(SweepArgs &a, int slices, size_t smem, cudaStream_t st) {
    int dev = 0, per_sm = 0;
    CTCB_CUDA_CHECK(cudaGetDevice(&dev));
    CTCB_CUDA_CHECK(cudaFuncSetAttribute(sweep_kernel<KI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CTCB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sweep_kernel<KI>, SW_THREADS, smem));
    const int capacity = per_sm * num_sms();
    if (2 * slices > capacity)
        return set_error(CTCB_EINVAL, "recurrent sweep: layerSize %d needs %d co-resident CTAs, device holds %d",
                         a.H, 2 * slices, capacity);
    const int ntiles = (a.B + SW_NB - 1) / SW_NB;
    int parts = capacity / (2 * slices);
    if (parts > ntiles) parts = ntiles;
    if (parts < 1) parts = 1;
    a.parts = parts;
    CTCB_CUDA_CHECK(cudaMemsetAsync(a.counters, 0, sizeof(unsigned int) * 1024, st));
    dim3 grid(slices, parts, 2);
    void *params[] = {&a};
    CTCB_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)sweep_kernel<KI>, grid, dim3(SW_THREADS), params, smem, st));
    count_launch();
    return CTCB_OK;
}

// counters: >= 2*ceil(B/8) uint32 of device scratch
int run_sweep(int mode, int T, int B, int H, const int32_t *Tlen, const float *pre, const float *Wf,
              const float *Wb, float *outF, float *outB, const float *actF, const float *actB, float maxAct,
              unsigned int *counters, cudaStream_t st) {
    SweepArgs a;
    a.mode = mode; a.T = T; a.B = B; a.H = H; a.Tlen = Tlen; a.pre = pre;
    a.W[0] = Wf; a.W[1] = Wb; a.out[0] = outF; a.out[1] = outB; a.act[0] = actF; a.act[1] = actB;
    a.maxAct = maxAct; a.counters = counters; a.parts = 1;
    const int slices = (H + SW_ROWS - 1) / SW_ROWS;
    const int Hp = (H + 3) / 4 * 4 + 4;
    const size_t smem = (size_t)SW_NB * Hp * sizeof(float);
    if (H % 32 == 0) {
        switch (H / 32) {
            case 4: return launch_sweep<4>(a, slices, smem, st);
            case 8: return launch_sweep<8>(a, slices, smem, st);
            case 16: return launch_sweep<16>(a, slices, smem, st);
            case 32: return launch_sweep<32>(a, slices, smem, st);
            default: break;
        }
    }
    return launch_sweep<0>(a, slices, smem, st);
}

int run_add2(const float *x, const float *y, float *z, int64_t n, cudaStream_t st) {
    int blocks = (int)((n + 1023) / 1024);
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
    if (blocks < 1) blocks = 1;
    add2_kernel<<<blocks, 256, 0, st>>>(x, y, z, n);
    CTCB_LAUNCH_CHECK();
    return CTCB_OK;
}

}  // namespace ctcb
