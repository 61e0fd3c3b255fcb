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
//   * alpha-tilde is spilled to a [T][2*32*P] fp64 workspace (coalesced double2 per pair) and read
//     back, prefetched, by the beta sweep, which scatters the normalised occupancies into a
//     shared-memory tile; the gradient of a whole tile is then written with coalesced row stores.
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
    int B, Tmax, K, Kp, blank, vec2, nbuf;
    float *grad, *nll;
    int32_t *skip;
    float *ws;            // alpha-tilde spill: [B][Tmax][Lpad] doubles
    int64_t ws_utt;       // doubles per utterance in ws
};

// ---- asynchronous tile copy: global -> shared without staging registers (LDGSTS) ------------
__device__ __forceinline__ void cp_async4(float *dst, const float *src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(float *dst, const float *src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Issue the copy of TT activation rows (frames t0 .. t0+TT-1) into a shared tile; rows beyond T are
// zero-filled.  Coalesced: consecutive lanes fetch consecutive classes of one frame.
__device__ __forceinline__ void issue_tile(const CtcArgs &a, const float *base, int t0, int T, float *te, int lane) {
    const int K = a.K, Kp = a.Kp;
    for (int r = 0; r < TT; ++r) {
        const int t = t0 + r;
        float *drow = te + r * Kp;
        if (t < T) {
            const float *row = base + (int64_t)t * a.fs;
            if (a.vec2) {        // rows and the padded shared rows are 8-byte aligned: half the copies
                for (int k = 2 * lane; k < K; k += 64) cp_async8(drow + k, row + k);
            } else {
                for (int k = lane; k < K; k += 32) cp_async4(drow + k, row + k);
            }
        } else {
            for (int k = lane; k < K; k += 32) drow[k] = 0.f;
        }
    }
    cp_async_commit();
}

// Turn a landed tile into e = exp(x - rowmax) in place; two lanes per frame (lane = r + 16 h).
// Returns Z_r = sum_k e[r][k] in the lanes that own row r (1 for rows beyond T or probability input).
__device__ __forceinline__ float tile_stats(const CtcArgs &a, int t0, int T, float *te, int lane) {
    float Z = 1.f;
    if (!a.is_prob) {
        const int K = a.K, Kp = a.Kp;
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

__device__ __forceinline__ int dexp_field(double v) { return (__double2hiint(v) >> 20) & 0x7ff; }
__device__ __forceinline__ double pow2_from_field(int biased) { return __hiloint2double(biased << 20, 0); }

constexpr float GFIX = 1073741824.f;   // occupancies are accumulated as 2^30 fixed point (native ATOMS.ADD)

// One warp per utterance.  alpha/beta live in registers in FLOAT64 (the reference's own arithmetic,
// ctc_fast.pyx:23-37): float32 cannot hold the product of the alpha and beta tails, which is what the
// gradient is made of.  Instead of dividing by the frame normaliser every frame (a warp reduction on
// the serial chain) the state is rescaled by a power of two derived from the PREVIOUS frame's largest
// exponent (one REDUX.MAX off the chain); the accumulated exponent goes into the loss.
template <int P>
__global__ void __launch_bounds__(256) ctc_warp_kernel(CtcArgs a) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int u = blockIdx.x * (blockDim.x >> 5) + wib;
    if (u >= a.B) return;
    const int K = a.K, Kp = a.Kp, blank = a.blank;
    // nbuf = 2: two e tiles (the copy of tile n+1 overlaps the work on tile n; used for small batches, where one
    // warp has an SM almost to itself); nbuf = 1: one e tile, 1/3 less shared memory -> 24 instead of 16 warps per
    // SM, the other warps hide the copy (large batches).  Then the occupancy tile.
    const int nbuf = a.nbuf;
    float *te0 = smem + (size_t)wib * (nbuf + 1) * TT * Kp;
    float *tg = te0 + nbuf * TT * Kp;
    const int T = min(a.Tlen[u], a.Tmax);
    const int lo = a.loff[u];
    const int nlab = a.loff[u + 1] - lo;
    const int L = 2 * nlab + 1;
    const float *base = a.acts + (int64_t)u * a.us;
    float *gbase = a.grad + (int64_t)u * a.us;
    double *wsu = reinterpret_cast<double *>(a.ws) + (int64_t)u * a.ws_utt;
    constexpr int LP = 64 * P;  // padded trellis row (doubles) in the workspace

    // per-lane label data for pairs i = lane*P + j  (blank s = 2i, label s = 2i+1)
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
    float logZ = 0.f;     // lanes < TT accumulate log Z of the rows they own
    int S = 0;            // accumulated power-of-two scaling of alpha
    double final_sum = 1.0;
    const int ntiles = (T + TT - 1) / TT;

    // ------------------------------------------------------------------ alpha sweep (:42-76)
    if (!short_utt && !fail) {
        double ab[P], al[P];
#pragma unroll
        for (int j = 0; j < P; ++j) ab[j] = al[j] = 0.0;
        // virtual state before frame 0: a unit mass on the first blank makes frame 0 an ordinary frame
        // (alpha[0,0] = p_blank, alpha[1,0] = p_label0, ctc_fast.pyx:42-47) -- no special case in the loop
        if (lane == 0) ab[0] = 1.0;
        int kscale = 0;   // power of two applied to the next frame
        if (nbuf == 2) issue_tile(a, base, 0, T, te0, lane);
        for (int tile = 0; tile < ntiles && !fail; ++tile) {
            const int t0 = tile * TT;
            float *te = te0 + ((nbuf == 2) ? (tile & 1) : 0) * TT * Kp;
            if (nbuf == 2 && tile + 1 < ntiles) {
                issue_tile(a, base, t0 + TT, T, te0 + ((tile + 1) & 1) * TT * Kp, lane);
                cp_async_wait<1>();
            } else {
                if (nbuf == 1) issue_tile(a, base, t0, T, te, lane);
                cp_async_wait<0>();
            }
            __syncwarp();
            const float Z = tile_stats(a, t0, T, te, lane);
            if (lane < TT) logZ += logf(Z);
            const int rmax = min(TT, T - t0);
            for (int r = 0; r < rmax; ++r) {
                const int t = t0 + r;
                const float *row = te + r * Kp;
                const double eb = (double)row[blank];
                int start = 2 * (T - t);
                start = (L <= start) ? 0 : L - start;
                double pl = __shfl_up_sync(0xffffffffu, al[P - 1], 1);
                if (lane == 0) pl = 0.0;
                const double sc = pow2_from_field(1023 + kscale);
                double mx = 0.0;
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const int i = lane * P + j;
                    const double el = (lab[j] >= 0) ? (double)row[lab[j]] : 0.0;
                    double b = (ab[j] + pl) * eb;
                    double l = (al[j] + ab[j] + (allow_a[j] ? pl : 0.0)) * el;
                    if (i > nlab) b = 0.0;
                    if (start > 0) {                    // warp-uniform: only the last |l| frames prune
                        if (2 * i < start) b = 0.0;
                        if (2 * i + 1 < start) l = 0.0;
                    }
                    pl = al[j];
                    ab[j] = b * sc;
                    al[j] = l * sc;
                    mx = fmax(mx, fmax(ab[j], al[j]));
                }
                S += kscale;
                const int emax = __reduce_max_sync(0xffffffffu, dexp_field(mx));
                if (emax == 0) { fail = true; break; }      // all mass gone: ZeroDivisionError in :70-76
                kscale = 1023 - emax;
                double2 *wrow = reinterpret_cast<double2 *>(wsu + (int64_t)t * LP) + lane * P;
#pragma unroll
                for (int j = 0; j < P; ++j) wrow[j] = make_double2(ab[j], al[j]);
            }
            __syncwarp();
        }
        if (!fail) {   // p(l|x) = alpha[L-1] + alpha[L-2] at the last frame
            double fs = 0.0;
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const int i = lane * P + j;
                if (i == nlab) fs += ab[j];
                if (i == nlab - 1) fs += al[j];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) fs += __shfl_xor_sync(0xffffffffu, fs, o);
            final_sum = fs;
            if (!(fs > 0.0)) fail = true;
        }
    }

    // ------------------------------------------------------------------ beta sweep + gradient
    float my_Zinv = 1.f;
    if (!fail) {
        double bb[P], bl[P];
#pragma unroll
        for (int j = 0; j < P; ++j) bb[j] = bl[j] = 0.0;
        // virtual state after the last frame: unit mass "beyond" the final blank, so that frame T-1 gets
        // pre[L-1] = pre[L-2] = 1 (ctc_fast.pyx:78-83) from the ordinary recurrence
#pragma unroll
        for (int j = 0; j < P; ++j)
            if (lane * P + j == nlab) bb[j] = 1.0;
        int kscale = 0;
        if (nbuf == 2) issue_tile(a, base, (ntiles - 1) * TT, T, te0 + ((ntiles - 1) & 1) * TT * Kp, lane);
        for (int tile = ntiles - 1; tile >= 0 && !fail; --tile) {
            const int t0 = tile * TT;
            float *te = te0 + ((nbuf == 2) ? (tile & 1) : 0) * TT * Kp;
            if (nbuf == 2 && tile > 0) {
                issue_tile(a, base, t0 - TT, T, te0 + ((tile - 1) & 1) * TT * Kp, lane);
                cp_async_wait<1>();
            } else {
                if (nbuf == 1) issue_tile(a, base, t0, T, te, lane);
                cp_async_wait<0>();
            }
            __syncwarp();
            const float Z = tile_stats(a, t0, T, te, lane);
            my_Zinv = 1.f / Z;
            unsigned *tgu = reinterpret_cast<unsigned *>(tg);
            for (int idx = lane; idx < TT * Kp; idx += 32) tgu[idx] = 0u;
            __syncwarp();
            const int rmax = min(TT, T - t0);
            if (!short_utt) {
                double2 an[P];   // alpha-tilde of the next frame to be processed (prefetched)
                {
                    const double2 *arow = reinterpret_cast<const double2 *>(wsu + (int64_t)(t0 + rmax - 1) * LP) + lane * P;
#pragma unroll
                    for (int j = 0; j < P; ++j) an[j] = arow[j];
                }
                for (int r = rmax - 1; r >= 0; --r) {
                    const int t = t0 + r;
                    const float *row = te + r * Kp;
                    unsigned *grow = tgu + r * Kp;
                    const double eb = (double)row[blank];
                    double2 av[P];
#pragma unroll
                    for (int j = 0; j < P; ++j) av[j] = an[j];
                    {   // pull the alpha-tilde row needed 4 frames from now into L1 (the spill sits in L2/HBM), then
                        // load the row of the next frame (t-1, clamped: the value is unused at t = 0)
                        const double *pf = wsu + (int64_t)max(t - 4, 0) * LP + lane * 2 * P;
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
                        if (P > 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(pf + 16));
                        const int tp = max(t - 1, 0);
                        const double2 *arow = reinterpret_cast<const double2 *>(wsu + (int64_t)tp * LP) + lane * P;
#pragma unroll
                        for (int j = 0; j < P; ++j) an[j] = arow[j];
                    }
                    const int end = min(2 * t + 2, L);
                    double nxb = __shfl_down_sync(0xffffffffu, bb[0], 1);
                    double nxl = __shfl_down_sync(0xffffffffu, bl[0], 1);
                    if (lane == 31) nxb = nxl = 0.0;
                    const double sc = pow2_from_field(1023 + kscale);
                    double nb[P], nl[P], w = 0.0, wb = 0.0, mx = 0.0;
#pragma unroll
                    for (int j = 0; j < P; ++j) {
                        const int i = lane * P + j;
                        const double el = (lab[j] >= 0) ? (double)row[lab[j]] : 0.0;
                        const double b1 = (j + 1 < P) ? bb[(j + 1 < P) ? j + 1 : j] : nxb;
                        const double l1 = (j + 1 < P) ? bl[(j + 1 < P) ? j + 1 : j] : nxl;
                        // pre-emission sums: beta[s,t] = pre[s] * p[lab(s),t]
                        double pb = bb[j] + bl[j];
                        double pll = bl[j] + b1 + (allow_b[j] ? l1 : 0.0);
                        if (i > nlab) pb = 0.0;
                        if (lab[j] < 0) pll = 0.0;
                        if (end < L) {                  // warp-uniform: only the first |l| frames prune
                            if (2 * i >= end) pb = 0.0;
                            if (2 * i + 1 >= end) pll = 0.0;
                        }
                        pb *= sc;
                        pll *= sc;
                        // occupancy numerators alpha*beta/p = alpha * pre  (no division by p, :117-136)
                        const double wbj = av[j].x * pb;
                        nl[j] = av[j].y * pll;          // reuse nl[] for the label numerators
                        wb += wbj;
                        w += wbj + nl[j];
                        nb[j] = pb * eb;
                        const double lnew = pll * el;
                        mx = fmax(mx, fmax(nb[j], lnew));
                        bl[j] = lnew;                   // old bl[j] no longer needed by later j
                    }
#pragma unroll
                    for (int j = 0; j < P; ++j) bb[j] = nb[j];
                    const int emax = __reduce_max_sync(0xffffffffu, dexp_field(mx));
                    kscale = 1023 - emax;
                    if (emax == 0) { fail = true; break; }        // beta mass gone: ZeroDivisionError in :109-114
                    // absum (sum of all numerators) and the blank occupancy without a float64 butterfly: scale the
                    // lane sums by a common power of two (largest lane exponent, one REDUX.MAX), quantise to 2^-24 of
                    // it and add with the integer REDUX.ADD.  alpha*beta can underflow float64 for every state of a
                    // frame (thousands of uninformative frames); the reference then leaves the frame at grad = p
                    // (absum == 0 -> :141-145, no skip), which is what wsum == 0 does here.
                    const int ew = __reduce_max_sync(0xffffffffu, dexp_field(w));
                    const double wscale = (ew >= 24) ? pow2_from_field(2070 - ew) : 0.0;   // largest lane sum -> [2^24, 2^25)
                    const unsigned wq = (unsigned)__double2uint_rz(w * wscale);
                    const unsigned wbq = (unsigned)__double2uint_rz(wb * wscale);
                    const unsigned wsum = __reduce_add_sync(0xffffffffu, wq);
                    const unsigned wbsum = __reduce_add_sync(0xffffffffu, wbq);
                    const float winv = (wsum > 0u) ? 1.f / (float)wsum : 0.f;
#pragma unroll
                    for (int j = 0; j < P; ++j) {
                        const float g = (float)(nl[j] * wscale) * winv;
                        if (g > 0.f) atomicAdd(grow + lab[j], (unsigned)(g * GFIX + 0.5f));
                    }
                    if (lane == 0) atomicAdd(grow + blank, (unsigned)((float)wbsum * winv * GFIX + 0.5f));
                }
            }
            __syncwarp();
            if (fail) break;
            // tile epilogue: grad = p - occupancy (:139-145), coalesced row stores
            for (int r = 0; r < rmax; ++r) {
                const float zinv = __shfl_sync(0xffffffffu, my_Zinv, r);
                float *orow = gbase + (int64_t)(t0 + r) * a.fs;
                for (int k = lane; k < K; k += 32)
                    orow[k] = te[r * Kp + k] * zinv - (float)tgu[r * Kp + k] * (1.f / GFIX);
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
        else nll = (float)(-(log(final_sum) - (double)S * 0.69314718055994530942 - (double)lz));
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
    return (size_t)B * (size_t)Tmax * (size_t)(64 * P) * sizeof(double);
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
    a.B = B; a.Tmax = Tmax; a.K = K; a.blank = blank;
    // 8-byte copies need even K, even strides, an 8-byte aligned base and an even shared row pitch
    a.vec2 = (K % 2 == 0) && (utt_stride % 2 == 0) && (frame_stride % 2 == 0) && (((uintptr_t)acts) % 8 == 0);
    // shared row pitch: odd (4-byte copies) or = 2 mod 32 (8-byte copies) so that the 16 rows of a tile start in
    // different banks for the per-row statistics pass
    a.Kp = a.vec2 ? ((K + 29) / 32 * 32 + 2) : (K | 1);
    a.grad = grad_out; a.nll = nll_out; a.skip = skip_out;
    a.ws = (float *)workspace; a.ws_utt = (int64_t)Tmax * 64 * P;

    // many utterances: favour occupancy (one e tile); few: favour the latency of each warp (double-buffered tiles)
    a.nbuf = (B >= 16 * num_sms()) ? 1 : 2;
    const size_t per_warp = (size_t)(a.nbuf + 1) * TT * a.Kp * sizeof(float);
    int wpb = 8;
    while (wpb > 1 && per_warp * wpb > 100 * 1024) wpb >>= 1;
    // small batches (the training step): spread the utterances over the SMs instead of packing 8 per CTA --
    // each warp is a latency-bound serial chain and gains from an otherwise idle SM
    while (wpb > 1 && (B + wpb - 1) / wpb < 2 * num_sms()) wpb >>= 1;
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
