// ctc.cu -- batched CTC loss + gradient on sm_100a, softmax fused.
//
// Replaces the CPU/float64 Cython routine /root/reference/ctc_fast/ctc-loss/ctc_fast.pyx:13-152
// (called once per utterance from ctc_fast/nnets/brnnet.py:175) and the six-kernel column softmax
// in front of it (brnnet.py:161-168).  Same algorithm -- per-frame RESCALED alpha/beta in the
// probability domain, window pruning [start,end) (:49-54), repeat-label rule (:64), gradient
// grad = p - occupancy/(p*absum) (:139-145), skip when a normaliser is 0 (:147-149) -- but laid
// out for the GPU:
//
//   * one WARP per utterance.  Trellis states are held in registers: lane owns P consecutive
//     (blank,label) pairs, s = 2i and 2i+1; the s-1/s-2 neighbours of the recurrence come from one
//     __shfl_up (alpha) / two __shfl_down (beta) per frame, the frame normaliser from a
//     warp-shuffle butterfly.  No block barrier anywhere.
//   * the T x K activations are streamed in TIME TILES of TT frames: a coalesced copy into shared
//     memory, then the softmax statistics of all TT frames at once (2 lanes per frame, off the
//     serial chain); the recurrence then gathers p[label] from the shared-memory tile.
//   * alpha-tilde is spilled to a [T][2*32*P] fp32 workspace (coalesced float2 per pair) and read
//     back, prefetched, by the beta sweep, which scatters alpha*beta into a shared-memory
//     occupancy tile; the gradient of a whole tile is then written with coalesced row stores.
//
// Scaling is arbitrary per frame (the gradient divides by absum[t], ctc_fast.pyx:133-145), so the
// recurrences run on e = exp(x - max) and the log-partition is added to the loss separately.
#include "common.cuh"
#include <math_constants.h>

namespace ctcb {

constexpr int TT = 16;  // frames per time tile

struct CtcArgs {
    const float *acts;
    int is_prob;
    int64_t us, fs;
    const int32_t *labels, *loff, *Tlen;
    int B, Tmax, K, Kp, blank;
    float *grad, *nll;
    int32_t *skip;
    float *ws;            // alpha-tilde spill: [B][Tmax][Lpad]
    int64_t ws_utt;       // floats per utterance in ws
};

// Load TT rows of activations into the shared tile and turn them into e = exp(x - rowmax).
// Returns (in lane r < TT) Z_r = sum_k e[r][k]; rows beyond T are zero-filled with Z = 1.
__device__ __forceinline__ float load_tile(const CtcArgs &a, const float *base, int t0, int T, float *te,
                                           int lane) {
    const int K = a.K, Kp = a.Kp;
#pragma unroll 4
    for (int r = 0; r < TT; ++r) {
        const int t = t0 + r;
        const float *row = base + (int64_t)t * a.fs;
        for (int k = lane; k < K; k += 32) te[r * Kp + k] = (t < T) ? __ldg(row + k) : 0.f;
    }
    __syncwarp();
    float Z = 1.f;
    if (!a.is_prob) {
        // two lanes per frame: lane = r + 16*h handles k = h, h+2, ...
        const int r = lane & (TT - 1), h = lane >> 4;
        float *row = te + r * Kp;
        float m = -CUDART_INF_F;
        for (int k = h; k < K; k += 2) m = fmaxf(m, row[k]);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
        float z = 0.f;
        for (int k = h; k < K; k += 2) {
            const float e = __expf(row[k] - m);
            row[k] = e;
            z += e;
        }
        z += __shfl_xor_sync(0xffffffffu, z, 16);
        Z = (t0 + r < T) ? z : 1.f;
    }
    __syncwarp();
    return Z;
}

template <int P>
__global__ void __launch_bounds__(256) ctc_warp_kernel(CtcArgs a) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int u = blockIdx.x * (blockDim.x >> 5) + wib;
    if (u >= a.B) return;
    const int K = a.K, Kp = a.Kp, blank = a.blank;
    float *te = smem + (size_t)wib * 2 * TT * Kp;  // e tile
    float *tg = te + TT * Kp;                      // occupancy tile
    const int T = min(a.Tlen[u], a.Tmax);
    const int lo = a.loff[u];
    const int nlab = a.loff[u + 1] - lo;
    const int L = 2 * nlab + 1;
    const float *base = a.acts + (int64_t)u * a.us;
    float *gbase = a.grad + (int64_t)u * a.us;
    float *wsu = a.ws + (int64_t)u * a.ws_utt;
    constexpr int LP = 64 * P;  // padded trellis row in the workspace

    // per-lane label data for pairs i = lane*P + j
    int lab[P];
    bool allow_a[P], allow_b[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int i = lane * P + j;
        lab[j] = (i < nlab) ? a.labels[lo + i] : -1;
        const int lprev = (i >= 1 && i < nlab) ? a.labels[lo + i - 1] : -1;
        const int lnext = (i + 1 < nlab) ? a.labels[lo + i + 1] : -1;
        allow_a[j] = (i >= 1 && i < nlab && lab[j] != lprev);       // ctc_fast.pyx:64-68
        allow_b[j] = (i + 1 < nlab && lab[j] != lnext);             // ctc_fast.pyx:104-108
    }

    const bool short_utt = (T < nlab);  // every window empty: reference returns (inf, p, False)
    bool fail = (T <= 0);
    float mant = 1.f;   // running product of frame normalisers: mantissa ...
    int expo = 0;       // ... and exponent
    float logZ = 0.f;   // lane r accumulates log Z of the rows it owns

    // ------------------------------------------------------------------ alpha sweep (:42-76)
    if (!short_utt && !fail) {
        float ab[P], al[P];
#pragma unroll
        for (int j = 0; j < P; ++j) ab[j] = al[j] = 0.f;
        for (int t0 = 0; t0 < T && !fail; t0 += TT) {
            const float Z = load_tile(a, base, t0, T, te, lane);
            if (lane < TT) logZ += logf(Z);
            const int rmax = min(TT, T - t0);
            for (int r = 0; r < rmax; ++r) {
                const int t = t0 + r;
                const float *row = te + r * Kp;
                const float eb = row[blank];
                float nb[P], nl[P], csum = 0.f;
                if (t == 0) {
#pragma unroll
                    for (int j = 0; j < P; ++j) {
                        const int i = lane * P + j;
                        nb[j] = (i == 0) ? eb : 0.f;
                        nl[j] = (i == 0 && nlab > 0) ? row[lab[j]] : 0.f;
                        csum += nb[j] + nl[j];
                    }
                } else {
                    int start = 2 * (T - t);
                    start = (L <= start) ? 0 : L - start;
                    float pl = __shfl_up_sync(0xffffffffu, al[P - 1], 1);
                    if (lane == 0) pl = 0.f;
#pragma unroll
                    for (int j = 0; j < P; ++j) {
                        const int i = lane * P + j;
                        const float el = (lab[j] >= 0) ? row[lab[j]] : 0.f;
                        float b = (ab[j] + pl) * eb;
                        float l = (al[j] + ab[j] + (allow_a[j] ? pl : 0.f)) * el;
                        if (2 * i < start || i > nlab) b = 0.f;
                        if (2 * i + 1 < start) l = 0.f;
                        pl = al[j];
                        nb[j] = b;
                        nl[j] = l;
                        csum += b + l;
                    }
                }
                const float c = warp_sum(csum);
                if (c == 0.f) { fail = true; break; }
                const float inv = 1.f / c;
                int e2;
                mant *= frexpf(c, &e2);
                expo += e2;
                int e3;
                mant = frexpf(mant, &e3);
                expo += e3;
                float2 *wrow = reinterpret_cast<float2 *>(wsu + (int64_t)t * LP) + lane * P;
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    ab[j] = nb[j] * inv;
                    al[j] = nl[j] * inv;
                    wrow[j] = make_float2(ab[j], al[j]);
                }
            }
            __syncwarp();
        }
    }

    // ------------------------------------------------------------------ beta sweep + gradient
    float my_absum = 0.f, my_Zinv = 1.f;
    if (!fail) {
        float bb[P], bl[P];
#pragma unroll
        for (int j = 0; j < P; ++j) bb[j] = bl[j] = 0.f;
        const int ntiles = (T + TT - 1) / TT;
        for (int tile = ntiles - 1; tile >= 0 && !fail; --tile) {
            const int t0 = tile * TT;
            const float Z = load_tile(a, base, t0, T, te, lane);
            my_Zinv = 1.f / Z;
            my_absum = 0.f;
            for (int idx = lane; idx < TT * Kp; idx += 32) tg[idx] = 0.f;
            __syncwarp();
            const int rmax = min(TT, T - t0);
            if (!short_utt) {
                // prefetch alpha-tilde of the first frame of this tile's sweep
                float2 an[P];
                {
                    const float2 *arow = reinterpret_cast<const float2 *>(wsu + (int64_t)(t0 + rmax - 1) * LP) + lane * P;
#pragma unroll
                    for (int j = 0; j < P; ++j) an[j] = arow[j];
                }
                for (int r = rmax - 1; r >= 0; --r) {
                    const int t = t0 + r;
                    const float *row = te + r * Kp;
                    float *grow = tg + r * Kp;
                    const float eb = row[blank];
                    float2 av[P];
#pragma unroll
                    for (int j = 0; j < P; ++j) av[j] = an[j];
                    if (t > 0) {  // prefetch next (earlier) frame; crosses into the previous tile's rows
                        const float2 *arow = reinterpret_cast<const float2 *>(wsu + (int64_t)(t - 1) * LP) + lane * P;
#pragma unroll
                        for (int j = 0; j < P; ++j) an[j] = arow[j];
                    }
                    float nb[P], nl[P], el[P], csum = 0.f;
                    if (t == T - 1) {   // :78-83
#pragma unroll
                        for (int j = 0; j < P; ++j) {
                            const int i = lane * P + j;
                            el[j] = (lab[j] >= 0) ? row[lab[j]] : 0.f;
                            nb[j] = (i == nlab) ? eb : 0.f;
                            nl[j] = (i == nlab - 1) ? el[j] : 0.f;
                            csum += nb[j] + nl[j];
                        }
                    } else {            // :84-114
                        const int end = min(2 * t + 2, L);
                        float nxb = __shfl_down_sync(0xffffffffu, bb[0], 1);
                        float nxl = __shfl_down_sync(0xffffffffu, bl[0], 1);
                        if (lane == 31) nxb = nxl = 0.f;
#pragma unroll
                        for (int j = 0; j < P; ++j) {
                            const int i = lane * P + j;
                            el[j] = (lab[j] >= 0) ? row[lab[j]] : 0.f;
                            const float b1 = (j + 1 < P) ? bb[(j + 1 < P) ? j + 1 : j] : nxb;
                            const float l1 = (j + 1 < P) ? bl[(j + 1 < P) ? j + 1 : j] : nxl;
                            float b = (bb[j] + bl[j]) * eb;
                            float l = (bl[j] + b1 + (allow_b[j] ? l1 : 0.f)) * el[j];
                            if (2 * i >= end || i > nlab) b = 0.f;
                            if (2 * i + 1 >= end) l = 0.f;
                            nb[j] = b;
                            nl[j] = l;
                            csum += b + l;
                        }
                    }
                    // occupancy numerators on the UNscaled beta; one interleaved butterfly for
                    // (normaliser, blank occupancy, absum)
                    float sb = 0.f, sa = 0.f, abl[P];
#pragma unroll
                    for (int j = 0; j < P; ++j) {
                        const float x = av[j].x * nb[j];
                        abl[j] = av[j].y * nl[j];
                        sb += x;
                        if (x != 0.f) sa += x / eb;                   // :122-125
                        if (abl[j] != 0.f) sa += abl[j] / el[j];      // :127-131
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        csum += __shfl_xor_sync(0xffffffffu, csum, o);
                        sb += __shfl_xor_sync(0xffffffffu, sb, o);
                        sa += __shfl_xor_sync(0xffffffffu, sa, o);
                    }
                    if (csum == 0.f) { fail = true; break; }
                    const float inv = 1.f / csum;
#pragma unroll
                    for (int j = 0; j < P; ++j) {
                        bb[j] = nb[j] * inv;
                        bl[j] = nl[j] * inv;
                        if (abl[j] != 0.f) atomicAdd(grow + lab[j], abl[j] * inv);
                    }
                    if (lane == 0) atomicAdd(grow + blank, sb * inv);
                    if (lane == r) my_absum = sa * inv;
                }
            }
            __syncwarp();
            if (fail) break;
            // tile epilogue: grad = p - G/(e*absum)  (:139-145), coalesced row stores
            for (int r = 0; r < rmax; ++r) {
                const float absum = __shfl_sync(0xffffffffu, my_absum, r);
                const float zinv = __shfl_sync(0xffffffffu, my_Zinv, r);
                float *orow = gbase + (int64_t)(t0 + r) * a.fs;
                for (int k = lane; k < K; k += 32) {
                    const float e = te[r * Kp + k];
                    const float tmp = e * absum;
                    const float p = e * zinv;
                    orow[k] = (tmp > 0.f) ? p - tg[r * Kp + k] / tmp : p;
                }
            }
            __syncwarp();
        }
    }

    // ------------------------------------------------------------------ outputs
    if (fail) {  // reference returns the zero-initialised grad on its failure path
        for (int t = 0; t < T; ++t) {
            float *orow = gbase + (int64_t)t * a.fs;
            for (int k = lane; k < K; k += 32) orow[k] = 0.f;
        }
    }
    for (int t = max(T, 0); t < a.Tmax; ++t) {  // padded frames carry no gradient
        float *orow = gbase + (int64_t)t * a.fs;
        for (int k = lane; k < K; k += 32) orow[k] = 0.f;
    }
    const float lz = warp_sum(logZ);
    if (lane == 0) {
        float nll;
        if (short_utt && !fail) nll = CUDART_INF_F;
        else nll = -(logf(mant) + (float)expo * 0.69314718055994531f - lz);
        a.nll[u] = nll;
        a.skip[u] = fail ? 1 : 0;
    }
}

// -------------------------------------------------------------------------------------------
// best path: per-frame argmax + collapse (ctc_fast.pyx:154-187).  One warp per utterance.
// -------------------------------------------------------------------------------------------
__global__ void ctc_best_path_kernel(const float *acts, int64_t us, int64_t fs, const int32_t *Tlen, int B,
                                     int Tmax, int K, int blank, int drop, int32_t *hyp, int32_t *align,
                                     int32_t *hlen) {
    const int lane = threadIdx.x & 31;
    const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (u >= B) return;
    const int T = min(Tlen[u], Tmax);
    int n = 0, prev = -1;
    for (int t = 0; t < T; ++t) {
        const float *row = acts + (int64_t)u * us + (int64_t)t * fs;
        float bv = -CUDART_INF_F;
        int bk = 0x7fffffff;
        for (int k = lane; k < K; k += 32) {
            const float v = row[k];
            if (v > bv) { bv = v; bk = k; }   // first maximum within the lane's stride
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (ov > bv || (ov == bv && ok < bk)) { bv = ov; bk = ok; }  // np.argmax: lowest index wins ties
        }
        const int b = bk;
        if (b == blank) { prev = b; continue; }
        if (drop && (b == 1 || b == 2 || b == 8)) { prev = b; continue; }
        if (t != 0 && b == prev) {
            if (lane == 0 && n > 0) align[(int64_t)u * Tmax + n - 1] = t;
            continue;
        }
        if (lane == 0) { hyp[(int64_t)u * Tmax + n] = b; align[(int64_t)u * Tmax + n] = t; }
        ++n;
        prev = b;
    }
    if (lane == 0) hlen[u] = n;
}

static int pairs_per_lane(int max_labels) {
    const int npairs = max_labels + 1;
    for (int p = 1; p <= 16; p *= 2)
        if (npairs <= 32 * p) return p;
    return 0;
}

}  // namespace ctcb

using namespace ctcb;

extern "C" size_t ctcb_ctc_workspace_bytes(int B, int Tmax, int max_labels) {
    const int P = pairs_per_lane(max_labels);
    if (P == 0 || B <= 0 || Tmax <= 0) return 0;
    return (size_t)B * (size_t)Tmax * (size_t)(64 * P) * sizeof(float);
}

extern "C" int ctcb_ctc_loss_grad_f32(const float *acts, int is_prob, int64_t utt_stride, int64_t frame_stride,
                                      const int32_t *labels, const int32_t *label_off, const int32_t *T_per_utt,
                                      int B, int Tmax, int K, int max_labels, int blank, float *grad_out,
                                      float *nll_out, int32_t *skip_out, void *workspace, size_t ws_bytes,
                                      void *stream) {
    if (B <= 0) return CTCB_OK;
    if (!acts || !labels || !label_off || !T_per_utt || !grad_out || !nll_out || !skip_out)
        return set_error(CTCB_EINVAL, "ctcb_ctc_loss_grad_f32: null pointer argument");
    if (K <= 0 || Tmax <= 0 || blank < 0 || blank >= K)
        return set_error(CTCB_EINVAL, "ctcb_ctc_loss_grad_f32: bad sizes K=%d Tmax=%d blank=%d", K, Tmax, blank);
    const int P = pairs_per_lane(max_labels);
    if (P == 0)
        return set_error(CTCB_EINVAL, "ctcb_ctc_loss_grad_f32: label sequences longer than 511 are not supported (got %d)", max_labels);
    const size_t need = ctcb_ctc_workspace_bytes(B, Tmax, max_labels);
    if (!workspace || ws_bytes < need)
        return set_error(CTCB_ENOMEM, "ctcb_ctc_loss_grad_f32: workspace %zu < %zu bytes", ws_bytes, need);

    CtcArgs a;
    a.acts = acts; a.is_prob = is_prob; a.us = utt_stride; a.fs = frame_stride;
    a.labels = labels; a.loff = label_off; a.Tlen = T_per_utt;
    a.B = B; a.Tmax = Tmax; a.K = K; a.Kp = K | 1; a.blank = blank;
    a.grad = grad_out; a.nll = nll_out; a.skip = skip_out;
    a.ws = (float *)workspace; a.ws_utt = (int64_t)Tmax * 64 * P;

    const size_t per_warp = (size_t)2 * TT * a.Kp * sizeof(float);
    int wpb = 8;
    while (wpb > 1 && per_warp * wpb > 72 * 1024) wpb >>= 1;
    const size_t smem = per_warp * wpb;
    if (smem > 200 * 1024)
        return set_error(CTCB_EINVAL, "ctcb_ctc_loss_grad_f32: K=%d too large for the shared-memory tile", K);
    const int grid = (B + wpb - 1) / wpb;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_P(PP)                                                                                   \
    case PP: {                                                                                         \
        CTCB_CUDA_CHECK(cudaFuncSetAttribute(ctc_warp_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        ctc_warp_kernel<PP><<<grid, wpb * 32, smem, st>>>(a);                                          \
        break;                                                                                         \
    }
    switch (P) {
        LAUNCH_P(1) LAUNCH_P(2) LAUNCH_P(4) LAUNCH_P(8) LAUNCH_P(16)
        default: return set_error(CTCB_EINVAL, "bad P");
    }
#undef LAUNCH_P
    CTCB_LAUNCH_CHECK();
    return CTCB_OK;
}

extern "C" int ctcb_ctc_best_path_f32(const float *acts, int64_t utt_stride, int64_t frame_stride,
                                      const int32_t *T_per_utt, int B, int Tmax, int K, int blank,
                                      int drop_swbd_noise, int32_t *hyp_out, int32_t *align_out,
                                      int32_t *hyp_len_out, void *stream) {
    if (B <= 0) return CTCB_OK;
    if (!acts || !T_per_utt || !hyp_out || !align_out || !hyp_len_out)
        return set_error(CTCB_EINVAL, "ctcb_ctc_best_path_f32: null pointer argument");
    const int wpb = 4;
    ctc_best_path_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
        acts, utt_stride, frame_stride, T_per_utt, B, Tmax, K, blank, drop_swbd_noise, hyp_out, align_out,
        hyp_len_out);
    CTCB_LAUNCH_CHECK();
    return CTCB_OK;
}
