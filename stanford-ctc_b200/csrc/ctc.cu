// ctc.cu -- batched CTC loss + gradient on sm_100a, softmax fused.
//
// Replaces the CPU/float64 Cython routine /root/reference/ctc_fast/ctc-loss/ctc_fast.pyx:13-152
// (called once per utterance from ctc_fast/nnets/brnnet.py:175) and the six-kernel column softmax
// in front of it (brnnet.py:161-168).  Same algorithm -- per-frame RESCALED alpha/beta in the
// probability domain, window pruning [start,end) (:49-54), repeat-label rule (:64), gradient
// grad = p - occupancy/(p*absum) (:139-145), skip when a normaliser is 0 (:147-149) -- but laid
// out for the GPU:
//
//   * Trellis states are held in registers: a lane owns P consecutive (blank,label) pairs, s = 2i and
//     2i+1; the s-1/s-2 neighbours of the recurrence come from one __shfl_up (alpha) / two
//     __shfl_down (beta) per frame, the per-frame rescaling from an integer REDUX.MAX over the
//     exponents (a power of two, applied one frame later).  No block barrier on the serial chain.
//   * the T x K activations are streamed in TIME TILES of 16 (or 8) frames: a coalesced copy into
//     shared memory, then the softmax statistics of all frames of the tile at once (2 or 4 lanes
//     per frame, off the serial chain); the recurrence then gathers e[label] from the tile.
//   * occupancies are scattered into a shared-memory tile as 2^30 fixed point (integer atomics:
//     order-independent, bit-reproducible); the gradient of a whole tile is then written with
//     coalesced row stores.
//
// Four kernel organisations over the same frame functions (chosen by batch size in
// ctcb_ctc_loss_grad_f32; CTCB_CTC / ctcb_debug_set_ctc_kernel force one; results are identical):
//   ctc_par_kernel   at most one utterance per SM (a training step): alpha on warp 0 and beta on warp 1
//                    over all frames at the same time, both spilled; then all warps turn the two planes
//                    into occupancies and gradient, one time tile each.
//   ctc_pair_kernel  a few utterances per SM: two warps meet in the middle of the trellis.
//   ctc_warp_kernel  large batches: one warp per utterance, alpha-tilde spilled to a [T][64P] fp64
//                    workspace (coalesced double2 per pair), read back by the beta sweep.
//   ctc_ckpt_kernel  the same without the spill: alpha checkpointed every 8 frames and recomputed per
//                    tile in shared memory (less than half the HBM traffic, more instructions; slower).
//
// Scaling is arbitrary per frame (the gradient divides by absum[t], ctc_fast.pyx:133-145), so the
// recurrences run on e = exp(x - max) and the log-partition is added to the loss separately.
#include "common.cuh"
#include <math_constants.h>
#include <stdlib.h>

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

// Issue the copy of the frames t0 .. min(t0+TT, T)-1 into a shared tile (rows beyond T keep whatever they held: their
// statistics are never used).  Coalesced: consecutive lanes fetch consecutive classes of one frame (of two frames when a
// frame is at most 16 copies wide); source and destination pointers advance by a row per pass.
__device__ __forceinline__ void issue_tile(const CtcArgs &a, const float *base, int t0, int T, float *te, int lane,
                                           int th = TT) {
    const int K = a.K, Kp = a.Kp, rows = min(th, T - t0);
    const int64_t fs = a.fs;
    const float *src = base + (int64_t)t0 * fs;
    if (a.vec2) {        // rows and the padded shared rows are 8-byte aligned: half the copies
        const int K2 = K >> 1;
        if (K2 <= 16) {
            const int sub = lane >> 4, kk = lane & 15;
            const float *s = src + sub * fs + 2 * kk;
            float *d = te + sub * Kp + 2 * kk;
            if (kk < K2)
                for (int r = sub; r < rows; r += 2, s += 2 * fs, d += 2 * Kp) cp_async8(d, s);
        } else if (K2 <= 32) {      // one copy per lane and frame: nothing but the copy and two pointer bumps per row
            const float *s = src + 2 * lane;
            float *d = te + 2 * lane;
            if (lane < K2)
                for (int r = 0; r < rows; ++r, s += fs, d += Kp) cp_async8(d, s);
        } else {
            const float *s = src + 2 * lane;
            float *d = te + 2 * lane;
            for (int r = 0; r < rows; ++r, s += fs, d += Kp)
                for (int k = 2 * lane; k < K; k += 64) cp_async8(d + (k - 2 * lane), s + (k - 2 * lane));
        }
    } else {
        const float *s = src + lane;
        float *d = te + lane;
        for (int r = 0; r < rows; ++r, s += fs, d += Kp)
            for (int k = lane; k < K; k += 32) cp_async4(d + (k - lane), s + (k - lane));
    }
    cp_async_commit();
}

// Pull the NEXT tile of a sweep towards the L2 while the current one is worked on (one line per lane: frame lane/2,
// half lane&1 of its K floats); the later cp.async then finds it on chip.
__device__ __forceinline__ void prefetch_tile_l2(const CtcArgs &a, const float *base, int t0, int T, int lane, int th = TT) {
    const int r = lane >> 1, h = lane & 1;
    if (t0 >= 0 && r < th && t0 + r < T) {
        const float *p = base + (int64_t)(t0 + r) * a.fs;
        for (int o = h * 32; o < a.K; o += 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
    }
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Turn a landed tile of TH frames into e = exp(x - rowmax) in place; 32/TH lanes per frame (lane = r + TH h).
// Returns Z_r = sum_k e[r][k] in the lanes that own row r (1 for rows beyond T or probability input).
// exp(x - m) = 2^(x*log2e - m*log2e): one FFMA and one MUFU.EX2 per class; the rounding of m*log2e is common to the whole
// row and cancels in e/Z.
template <int TH = TT>
__device__ __forceinline__ float tile_stats(const CtcArgs &a, int t0, int T, float *te, int lane) {
    constexpr int LPR = 32 / TH;      // lanes per row
    float Z = 1.f;
    if (!a.is_prob) {
        const int K = a.K, Kp = a.Kp;
        const int r = lane & (TH - 1), h = lane / TH;
        float *row = te + r * Kp;
        const float L2E = 1.4426950408889634f;
        float m = -CUDART_INF_F, z = 0.f;
        if (a.vec2 && K <= 64) {        // at most 32/LPR float2 per lane: straight-line, predicated
            constexpr int NI = 32 / LPR;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const int k = 2 * (h + LPR * i);
                if (k < K) {
                    const float2 v = *reinterpret_cast<const float2 *>(row + k);
                    m = fmaxf(m, fmaxf(v.x, v.y));
                }
            }
#pragma unroll
            for (int o = TH; o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            const float ml = m * L2E;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const int k = 2 * (h + LPR * i);
                if (k < K) {
                    float2 v = *reinterpret_cast<const float2 *>(row + k);
                    v.x = ex2_approx(fmaf(v.x, L2E, -ml));
                    v.y = ex2_approx(fmaf(v.y, L2E, -ml));
                    *reinterpret_cast<float2 *>(row + k) = v;
                    z += v.x + v.y;
                }
            }
        } else if (a.vec2) {
            for (int k = 2 * h; k < K; k += 2 * LPR) {
                const float2 v = *reinterpret_cast<const float2 *>(row + k);
                m = fmaxf(m, fmaxf(v.x, v.y));
            }
#pragma unroll
            for (int o = TH; o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            const float ml = m * L2E;
            for (int k = 2 * h; k < K; k += 2 * LPR) {
                float2 v = *reinterpret_cast<const float2 *>(row + k);
                v.x = ex2_approx(fmaf(v.x, L2E, -ml));
                v.y = ex2_approx(fmaf(v.y, L2E, -ml));
                *reinterpret_cast<float2 *>(row + k) = v;
                z += v.x + v.y;
            }
        } else {
            for (int k = h; k < K; k += LPR) m = fmaxf(m, row[k]);
#pragma unroll
            for (int o = TH; o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            const float ml = m * L2E;
            for (int k = h; k < K; k += LPR) {
                const float e = ex2_approx(fmaf(row[k], L2E, -ml));
                row[k] = e;
                z += e;
            }
        }
#pragma unroll
        for (int o = TH; o < 32; o <<= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
        Z = (t0 + r < T) ? z : 1.f;
    }
    __syncwarp();
    return Z;
}

// The e tiles carry one guaranteed ZERO per row, at class index K (the row pitch Kp leaves room): lanes without a label
// read their emission there, which keeps their states at exactly 0 without a select on the serial chain.  The copies and
// the statistics only ever write k < K, so the slot is written once.
__device__ __forceinline__ void init_tile_pads(const CtcArgs &a, float *te, int ntile, int lane, int th = TT) {
    for (int r = lane; r < ntile * th; r += 32)
        for (int k = a.K; k < a.Kp; ++k) te[r * a.Kp + k] = 0.f;
}
__device__ __forceinline__ void zero_occupancy_tile(const CtcArgs &a, float *tg, int lane, int th = TT) {
    unsigned *tgu = reinterpret_cast<unsigned *>(tg);
    for (int idx = lane; idx < th * a.Kp; idx += 32) tgu[idx] = 0u;
}

__device__ __forceinline__ double pow2_from_field(int biased) { return __hiloint2double(biased << 20, 0); }
__device__ __forceinline__ float pow2f_from_int(int k) { return __int_as_float((127 + k) << 23); }   // -126 <= k <= 127

constexpr float GFIX = 1073741824.f;   // occupancies are accumulated as 2^30 fixed point (native ATOMS.ADD)
constexpr int KSCALE_MAX = 126;        // largest per-frame rescale: e * 2^k must stay a float

// alpha/beta live in registers in FLOAT64 (the reference's own arithmetic, ctc_fast.pyx:23-37): float32
// cannot hold the product of the alpha and beta tails, which is what the gradient is made of.  Instead of
// dividing by the frame normaliser every frame (a warp reduction on the serial chain) the state is rescaled
// by a power of two derived from the PREVIOUS frame's largest exponent; the accumulated exponent goes into the loss.
// What the serial chain of a frame carries is kept to the recurrence itself:
//   * the power of two is folded into the EMISSIONS while they are still float (exact; off the chain);
//   * all states are >= 0, so the high word of a double is monotone in its value: the largest exponent of a frame is one
//     integer max per state and one REDUX.MAX, no float64 compare/select;
//   * lanes without a label read their emission from the tile's zero slot (init_tile_pads): their label state is exactly 0,
//     and with it every state beyond the end of the label sequence -- no masks.
//
// Per-lane view of one utterance: pairs i = lane*P + j  (blank s = 2i, label s = 2i+1).
template <int P>
struct LaneLabels {
    int lab[P];         // label of pair i, or K (the zero slot) when the pair has none
    bool allow_a[P], allow_b[P], valid[P];
    double m_a[P <= 2 ? P : 1], m_b[P <= 2 ? P : 1], m_first;   // the same as 0/1 factors (P <= 2: registers to spare)
    int nlab, L;
    __device__ __forceinline__ void load(const CtcArgs &a, int lo, int nl, int lane) {
        nlab = nl; L = 2 * nl + 1;
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int i = lane * P + j;
            const int l = (i < nlab) ? a.labels[lo + i] : -1;
            const int lprev = (i >= 1 && i < nlab) ? a.labels[lo + i - 1] : -1;
            const int lnext = (i + 1 < nlab) ? a.labels[lo + i + 1] : -1;
            allow_a[j] = (i >= 1 && i < nlab && l != lprev);       // ctc_fast.pyx:64-68
            allow_b[j] = (i + 1 < nlab && l != lnext);             // ctc_fast.pyx:104-108
            lab[j] = (l >= 0) ? l : a.K;
            valid[j] = (l >= 0);
            if (P <= 2) {
                m_a[j < 2 ? j : 0] = allow_a[j] ? 1.0 : 0.0;
                m_b[j < 2 ? j : 0] = allow_b[j] ? 1.0 : 0.0;
            }
        }
        m_first = (lane > 0) ? 1.0 : 0.0;
    }
};

// One frame of the alpha recurrence (:49-76) on e-values `row`.  ab/al: blank/label states of the previous
// frame in, of frame t out (scaled by 2^kscale of the previous frame).  Returns false when all mass is gone.
// WIN = false is for frames the caller knows to be outside the pruning window of the last |l| frames (start == 0).
template <int P, bool WIN>
__device__ __forceinline__ bool alpha_frame(const LaneLabels<P> &q, const float *row, int blank, int lane, int T, int t,
                                            double (&ab)[P], double (&al)[P], int &kscale, int &S) {
    const float scf = pow2f_from_int(kscale);
    const double eb = (double)(row[blank] * scf);
    int start = 0;
    if (WIN) {
        start = 2 * (T - t);
        start = (q.L <= start || t == 0) ? 0 : q.L - start;   // the reference sets frame 0 without a window (:42-47)
    }
    double pl = __shfl_up_sync(0xffffffffu, al[P - 1], 1);
    int hmax = 0;
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int i = lane * P + j;
        const double el = (double)(row[q.lab[j]] * scf);
        double b, l = al[j] + ab[j];
        if (P <= 2) {       // the two conditional terms as exact multiply-adds with 0/1 (one DFMA instead of add + select)
            b = fma(pl, (j > 0) ? 1.0 : q.m_first, ab[j]);
            l = fma(pl, q.m_a[j < 2 ? j : 0], l);
        } else {
            b = ab[j];
            if (j > 0 || lane > 0) b += pl;     // lane 0 has no left neighbour (its shuffle returns its own value)
            if (q.allow_a[j]) l += pl;
        }
        b *= eb;
        l *= el;
        if (WIN && start > 0) {             // warp-uniform: only the last |l| frames prune
            if (2 * i < start) b = 0.0;
            if (2 * i + 1 < start) l = 0.0;
        }
        pl = al[j];
        ab[j] = b;
        al[j] = l;
        hmax = max(hmax, max(__double2hiint(b), __double2hiint(l)));
    }
    S += kscale;
    const int emax = __reduce_max_sync(0xffffffffu, hmax) >> 20;
    if (emax == 0) return false;            // all mass gone: ZeroDivisionError in :70-76
    kscale = min(1023 - emax, KSCALE_MAX);
    return true;
}
// does any frame up to t_last fall into alpha's pruning window (the last |l| frames)?
__device__ __forceinline__ bool alpha_window(int L, int T, int t_last) { return L > 2 * (T - t_last); }
// does any frame from t_first on fall into beta's (the first |l| frames)?
__device__ __forceinline__ bool beta_window(int L, int t_first) { return 2 * t_first + 2 < L; }

// One frame of the beta recurrence (:85-114).  bb/bl: states of frame t+1 in, of frame t out.  pb/pll receive the
// PRE-emission sums of frame t (any power-of-two scale): beta[s,t] = pre[s] * p[lab(s),t], so alpha*beta/p = alpha*pre.
// The last pair of lane 31 never has a label (|l| + 1 <= 32P pairs): what its shuffle brings in is multiplied by the zero
// slot, and its pre-emission sum only ever meets an alpha state that is 0.
// WIN = false is for frames the caller knows to be outside the pruning window of the first |l| frames (end == L).
template <int P, bool WIN>
__device__ __forceinline__ bool beta_frame(const LaneLabels<P> &q, const float *row, int blank, int lane, int t,
                                           double (&bb)[P], double (&bl)[P], int &kscale, double (&pb)[P],
                                           double (&pll)[P]) {
    const float scf = pow2f_from_int(kscale);
    const double eb = (double)(row[blank] * scf);
    const int end = WIN ? min(2 * t + 2, q.L) : q.L;
    const double nxb = __shfl_down_sync(0xffffffffu, bb[0], 1);
    const double nxl = __shfl_down_sync(0xffffffffu, bl[0], 1);
    double nb[P];
    int hmax = 0;
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int i = lane * P + j;
        const double el = (double)(row[q.lab[j]] * scf);
        const double b1 = (j + 1 < P) ? bb[(j + 1 < P) ? j + 1 : j] : nxb;
        const double l1 = (j + 1 < P) ? bl[(j + 1 < P) ? j + 1 : j] : nxl;
        double xb = bb[j] + bl[j];
        double xl = bl[j] + b1;
        if (P <= 2) xl = fma(l1, q.m_b[j < 2 ? j : 0], xl);
        else if (q.allow_b[j]) xl += l1;
        if (WIN && end < q.L) {             // warp-uniform: only the first |l| frames prune
            if (2 * i >= end) xb = 0.0;
            if (2 * i + 1 >= end) xl = 0.0;
        }
        pb[j] = xb;
        pll[j] = xl;
        nb[j] = xb * eb;
        const double lnew = xl * el;
        hmax = max(hmax, max(__double2hiint(nb[j]), __double2hiint(lnew)));
        bl[j] = lnew;                       // old bl[j] no longer needed by later j
    }
#pragma unroll
    for (int j = 0; j < P; ++j) bb[j] = nb[j];
    const int emax = __reduce_max_sync(0xffffffffu, hmax) >> 20;
    kscale = min(1023 - emax, KSCALE_MAX);
    return emax != 0;                       // beta mass gone: ZeroDivisionError in :109-114
}

// Occupancies of one frame from alpha-tilde (xb, xl) and the beta pre-emission sums (pb, pll), any power-of-two
// scaling of either: numerators alpha*pre (no division by p, :117-136), normalised by their sum (absum, :133-145)
// and scattered into the frame's row of the shared occupancy tile as 2^30 fixed point.
// absum and the blank occupancy come without a float64 butterfly: scale the lane sums by a common power of two
// (largest lane exponent, one REDUX.MAX), quantise to 2^-24 of it and add with the integer REDUX.ADD.  alpha*beta
// can underflow float64 for every state of a frame (thousands of uninformative frames); the reference then leaves
// the frame at grad = p (absum == 0 -> :141-145, no skip), which is what wsum == 0 does here.
template <int P>
__device__ __forceinline__ void occupancy_frame(const LaneLabels<P> &q, int blank, int lane, const double (&xb)[P],
                                                const double (&xl)[P], const double (&pb)[P], const double (&pll)[P],
                                                unsigned *grow) {
    double nl[P], w = 0.0, wb = 0.0;
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const double wbj = xb[j] * pb[j];
        nl[j] = xl[j] * pll[j];
        wb += wbj;
        w += wbj + nl[j];
    }
    const int ew = __reduce_max_sync(0xffffffffu, __double2hiint(w)) >> 20;
    const double wscale = (ew >= 24) ? pow2_from_field(2070 - ew) : 0.0;   // largest lane sum -> [2^24, 2^25)
    const unsigned wq = (unsigned)__double2uint_rz(w * wscale);
    const unsigned wbq = (unsigned)__double2uint_rz(wb * wscale);
    const unsigned wsum = __reduce_add_sync(0xffffffffu, wq);
    const unsigned wbsum = __reduce_add_sync(0xffffffffu, wbq);
    // 1 / wsum: wsum is 0 or in [2^24, 2^30], so the reciprocal needs none of the range checks of a general division
    // (MUFU.RCP and one Newton step -- the correctly rounded quotient's fast path)
    const float wf = (float)wsum;
    float rinv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(wf));
    rinv = fmaf(rinv, -fmaf(wf, rinv, -1.f), rinv);
    const float winv = (wsum > 0u) ? rinv : 0.f;
    // every pair WITH a label adds its occupancy (a 0 adds nothing; pairs without a label are predicated off rather than
    // branched around); the blank occupancy is the same value in all lanes, lane 0 adds it
    const unsigned gaddr = (unsigned)__cvta_generic_to_shared(grow);
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const float g = (float)(nl[j] * wscale) * winv;
        const unsigned v = (unsigned)(g * GFIX + 0.5f);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}"
                     ::"r"(gaddr + 4u * (unsigned)q.lab[j]), "r"(v), "r"((int)q.valid[j]) : "memory");
    }
    {
        const unsigned v = (unsigned)((float)wbsum * winv * GFIX + 0.5f);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}"
                     ::"r"(gaddr + 4u * (unsigned)blank), "r"(v), "r"(lane) : "memory");
    }
}

// Everything one warp needs to walk over the time tiles of one utterance.
struct WarpCtx {
    const float *base;      // activations of the utterance
    float *gbase;           // its gradient
    double *wsu;            // its trellis spill, [T][64P] doubles
    float *te0, *tg;        // shared: nbuf e tiles, one occupancy tile
    int T, ntiles, nbuf, lane;
};

// Load tile `tile` (double-buffered: the copy of `next` -- or -1 -- is issued first and stays in flight).
template <int TH = TT>
__device__ __forceinline__ float *fetch_tile(const CtcArgs &a, const WarpCtx &c, int tile, int next) {
    float *te = c.te0 + ((c.nbuf == 2) ? (tile & 1) : 0) * TH * a.Kp;
    if (c.nbuf == 2 && next >= 0) {
        issue_tile(a, c.base, next * TH, c.T, c.te0 + (next & 1) * TH * a.Kp, c.lane, TH);
        cp_async_wait<1>();
    } else {
        if (c.nbuf == 1) issue_tile(a, c.base, tile * TH, c.T, te, c.lane, TH);
        cp_async_wait<0>();
    }
    __syncwarp();
    return te;
}

// grad = p - occupancy (:139-145) for the rmax frames of a tile, coalesced row stores.  The occupancy entries are set
// back to zero as they are consumed, so the tile is ready for the next one without a separate pass.
__device__ __forceinline__ void tile_epilogue(const CtcArgs &a, const WarpCtx &c, int t0, int rmax, const float *te,
                                              unsigned *tgu, float my_Zinv) {
    const float gs = 1.f / GFIX;
    const int K = a.K, Kp = a.Kp;
    const int64_t fs = a.fs;
    if (a.vec2) {
        float *orow = c.gbase + (int64_t)t0 * fs + 2 * c.lane;
        const float *erow = te + 2 * c.lane;
        unsigned *grow = tgu + 2 * c.lane;
        if (K <= 64) {          // one float2 per lane and frame
            const bool on = 2 * c.lane < K;
            for (int r = 0; r < rmax; ++r, orow += fs, erow += Kp, grow += Kp) {
                const float zinv = __shfl_sync(0xffffffffu, my_Zinv, r);
                if (on) {
                    const float2 e = *reinterpret_cast<const float2 *>(erow);
                    const uint2 g = *reinterpret_cast<const uint2 *>(grow);
                    *reinterpret_cast<uint2 *>(grow) = make_uint2(0u, 0u);
                    float2 v;
                    v.x = fmaf(-(float)g.x, gs, e.x * zinv);
                    v.y = fmaf(-(float)g.y, gs, e.y * zinv);
                    *reinterpret_cast<float2 *>(orow) = v;
                }
            }
        } else
        for (int r = 0; r < rmax; ++r, orow += fs, erow += Kp, grow += Kp) {
            const float zinv = __shfl_sync(0xffffffffu, my_Zinv, r);
            for (int k = 2 * c.lane; k < K; k += 64) {
                const int o = k - 2 * c.lane;
                const float2 e = *reinterpret_cast<const float2 *>(erow + o);
                const uint2 g = *reinterpret_cast<const uint2 *>(grow + o);
                *reinterpret_cast<uint2 *>(grow + o) = make_uint2(0u, 0u);
                float2 v;
                v.x = fmaf(-(float)g.x, gs, e.x * zinv);
                v.y = fmaf(-(float)g.y, gs, e.y * zinv);
                *reinterpret_cast<float2 *>(orow + o) = v;
            }
        }
    } else {
        float *orow = c.gbase + (int64_t)t0 * fs + c.lane;
        const float *erow = te + c.lane;
        unsigned *grow = tgu + c.lane;
        for (int r = 0; r < rmax; ++r, orow += fs, erow += Kp, grow += Kp) {
            const float zinv = __shfl_sync(0xffffffffu, my_Zinv, r);
            for (int k = c.lane; k < K; k += 32) {
                const int o = k - c.lane;
                const unsigned g = grow[o];
                grow[o] = 0u;
                orow[o] = fmaf(-(float)g, gs, erow[o] * zinv);
            }
        }
    }
    __syncwarp();
}

// alpha over tiles [tile_from, tile_to), ascending (:42-76).  COMBINE = false: the scaled states of every frame are
// spilled to the workspace for a later beta sweep.  COMBINE = true: the workspace already holds the beta
// pre-emission sums of these frames (stored by beta_tiles<P,false>); occupancies and gradient are produced here.
// The workspace row of a frame is loaded at the top of its iteration (its latency hides behind the recurrence, which does
// not need it) from a line that was pulled into L1 four frames earlier; all addressing is pointer increments.
template <int P, bool COMBINE, int TH = TT>
__device__ __forceinline__ bool alpha_tiles(const CtcArgs &a, const WarpCtx &c, const LaneLabels<P> &q, int tile_from,
                                            int tile_to, bool recur, double (&ab)[P], double (&al)[P], int &kscale,
                                            int &S, float &logZ) {
    constexpr int LP = 64 * P;
    const int lane = c.lane, T = c.T;
    if (c.nbuf == 2 && tile_from < tile_to) issue_tile(a, c.base, tile_from * TH, T, c.te0 + (tile_from & 1) * TH * a.Kp, lane, TH);
    for (int tile = tile_from; tile < tile_to; ++tile) {
        const int t0 = tile * TH;
        const float *te = fetch_tile<TH>(a, c, tile, (tile + 1 < tile_to) ? tile + 1 : -1);
        const float Z = tile_stats<TH>(a, t0, T, const_cast<float *>(te), lane);
        if (lane < TH) logZ += logf(Z);
        const int rmax = min(TH, T - t0);
        const bool win = alpha_window(q.L, T, t0 + rmax - 1);
        unsigned *tgu = reinterpret_cast<unsigned *>(c.tg);
        if (recur) {
            double2 *wrow = reinterpret_cast<double2 *>(c.wsu + (int64_t)t0 * LP) + lane * P;
            const float *row = te;
            unsigned *grow = tgu;
            for (int r = 0; r < rmax; ++r, wrow += 32 * P, row += a.Kp, grow += a.Kp) {
                const int t = t0 + r;
                double pb[P], pll[P];
                if (COMBINE) {
                    if (t + 4 < T) {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(wrow + 4 * 32 * P));
                        if (P > 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(wrow + 4 * 32 * P + 8));
                    }
#pragma unroll
                    for (int j = 0; j < P; ++j) { const double2 v = wrow[j]; pb[j] = v.x; pll[j] = v.y; }
                }
                const bool ok = win ? alpha_frame<P, true>(q, row, a.blank, lane, T, t, ab, al, kscale, S)
                                    : alpha_frame<P, false>(q, row, a.blank, lane, T, t, ab, al, kscale, S);
                if (!ok) return false;
                if (COMBINE) {
                    occupancy_frame<P>(q, a.blank, lane, ab, al, pb, pll, grow);
                } else {
#pragma unroll
                    for (int j = 0; j < P; ++j) wrow[j] = make_double2(ab[j], al[j]);
                }
            }
        }
        __syncwarp();
        if (COMBINE) tile_epilogue(a, c, t0, rmax, te, tgu, 1.f / Z);
    }
    return true;
}

// beta over tiles tile_from, tile_from-1, ..., tile_to (descending, :78-114).  COMBINE = true: the workspace holds
// alpha-tilde of these frames; occupancies and gradient are produced here.  COMBINE = false: the pre-emission sums
// of every frame are spilled for a later alpha_tiles<P,true>.
template <int P, bool COMBINE, int TH = TT>
__device__ __forceinline__ bool beta_tiles(const CtcArgs &a, const WarpCtx &c, const LaneLabels<P> &q, int tile_from,
                                           int tile_to, bool recur, double (&bb)[P], double (&bl)[P], int &kscale) {
    constexpr int LP = 64 * P;
    const int lane = c.lane, T = c.T;
    if (c.nbuf == 2 && tile_from >= tile_to) issue_tile(a, c.base, tile_from * TH, T, c.te0 + (tile_from & 1) * TH * a.Kp, lane, TH);
    for (int tile = tile_from; tile >= tile_to; --tile) {
        const int t0 = tile * TH;
        const float *te = fetch_tile<TH>(a, c, tile, (tile > tile_to) ? tile - 1 : -1);
        const float Z = tile_stats<TH>(a, t0, T, const_cast<float *>(te), lane);
        unsigned *tgu = reinterpret_cast<unsigned *>(c.tg);
        const int rmax = min(TH, T - t0);
        const bool win = beta_window(q.L, t0);
        if (recur) {
            double2 *wrow = reinterpret_cast<double2 *>(c.wsu + (int64_t)(t0 + rmax - 1) * LP) + lane * P;
            const float *row = te + (rmax - 1) * a.Kp;
            unsigned *grow = tgu + (rmax - 1) * a.Kp;
            for (int r = rmax - 1; r >= 0; --r, wrow -= 32 * P, row -= a.Kp, grow -= a.Kp) {
                const int t = t0 + r;
                double xb[P], xl[P];
                if (COMBINE) {
                    // pull the alpha-tilde row needed 4 frames from now into L1 (the spill sits in L2/HBM)
                    if (t >= 4) {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(wrow - 4 * 32 * P));
                        if (P > 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(wrow - 4 * 32 * P + 8));
                    }
#pragma unroll
                    for (int j = 0; j < P; ++j) { const double2 v = wrow[j]; xb[j] = v.x; xl[j] = v.y; }
                }
                double pb[P], pll[P];
                const bool ok = win ? beta_frame<P, true>(q, row, a.blank, lane, t, bb, bl, kscale, pb, pll)
                                    : beta_frame<P, false>(q, row, a.blank, lane, t, bb, bl, kscale, pb, pll);
                if (!ok) return false;
                if (COMBINE) {
                    occupancy_frame<P>(q, a.blank, lane, xb, xl, pb, pll, grow);
                } else {
#pragma unroll
                    for (int j = 0; j < P; ++j) wrow[j] = make_double2(pb[j], pll[j]);
                }
            }
        }
        __syncwarp();
        if (COMBINE) tile_epilogue(a, c, t0, rmax, te, tgu, 1.f / Z);
    }
    return true;
}

// p(l|x) at the last frame = alpha[L-1] + alpha[L-2] (:76), in the scaled domain
// A one-frame utterance never leaves the reference's un-windowed initialisation, whose normaliser is
// alpha[0,0] + alpha[1,0] (:42-47): that sum is what it returns then.
template <int P>
__device__ __forceinline__ double final_mass(const LaneLabels<P> &q, int lane, int T, const double (&ab)[P],
                                             const double (&al)[P]) {
    double fs = 0.0;
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int i = lane * P + j;
        if (T == 1) {
            if (i == 0) fs += ab[j] + al[j];
        } else {
            if (i == q.nlab) fs += ab[j];
            if (i == q.nlab - 1) fs += al[j];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fs += __shfl_xor_sync(0xffffffffu, fs, o);
    return fs;
}

__device__ __forceinline__ void zero_rows(const CtcArgs &a, float *gbase, int t_from, int t_to, int lane, int nlanes) {
    for (int t = t_from; t < t_to; ++t) {
        float *orow = gbase + (int64_t)t * a.fs;
        for (int k = lane; k < a.K; k += nlanes) orow[k] = 0.f;
    }
}

__device__ __forceinline__ void write_loss(const CtcArgs &a, int u, bool short_utt, bool fail, double final_sum, int S,
                                           float lz) {
    float nll;
    if (short_utt && !fail) nll = CUDART_INF_F;    // every window empty: the reference returns (inf, p, False)
    else nll = (float)(-(log(final_sum) - (double)S * 0.69314718055994530942 - (double)lz));
    a.nll[u] = nll;
    a.skip[u] = fail ? 1 : 0;
}

// Throughput shape: ONE WARP per utterance -- alpha over all frames (spilled), then beta with the gradient.
// TH = frames per tile: 16, or 8 for large batches of short label sequences (half the shared memory and a 64-register
// build: 32 instead of 24 resident warps per SM -- the kernel is bound by the latency of its serial chains).
template <int P, int TH>
__global__ void __launch_bounds__(256, P <= 2 ? (TH == 8 ? 4 : 3) : 1) ctc_warp_kernel(CtcArgs a) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int u = blockIdx.x * (blockDim.x >> 5) + wib;
    if (u >= a.B) return;
    // nbuf = 2: two e tiles (the copy of tile n+1 overlaps the work on tile n; used for small batches, where one
    // warp has an SM almost to itself); nbuf = 1: one e tile, 1/3 less shared memory -> 24 instead of 16 warps per
    // SM, the other warps hide the copy (large batches).  Then the occupancy tile.
    WarpCtx c;
    c.nbuf = a.nbuf; c.lane = lane;
    c.te0 = smem + (size_t)wib * (a.nbuf + 1) * TH * a.Kp;
    c.tg = c.te0 + a.nbuf * TH * a.Kp;
    c.T = min(a.Tlen[u], a.Tmax);
    c.ntiles = (c.T + TH - 1) / TH;
    c.base = a.acts + (int64_t)u * a.us;
    c.gbase = a.grad + (int64_t)u * a.us;
    c.wsu = reinterpret_cast<double *>(a.ws) + (int64_t)u * a.ws_utt;
    const int lo = a.loff[u];
    LaneLabels<P> q;
    q.load(a, lo, a.loff[u + 1] - lo, lane);
    init_tile_pads(a, c.te0, a.nbuf, lane, TH);
    zero_occupancy_tile(a, c.tg, lane, TH);
    __syncwarp();

    const bool short_utt = (c.T < q.nlab);
    bool fail = (c.T <= 0);
    float logZ = 0.f;     // lanes < TH accumulate log Z of the rows they own
    int S = 0;            // accumulated power-of-two scaling of alpha
    double final_sum = 1.0;

    if (!short_utt && !fail) {
        double ab[P], al[P];
#pragma unroll
        for (int j = 0; j < P; ++j) ab[j] = al[j] = 0.0;
        // virtual state before frame 0: a unit mass on the first blank makes frame 0 an ordinary frame
        // (alpha[0,0] = p_blank, alpha[1,0] = p_label0, ctc_fast.pyx:42-47) -- no special case in the loop
        if (lane == 0) ab[0] = 1.0;
        int kscale = 0;   // power of two applied to the next frame
        fail = !alpha_tiles<P, false, TH>(a, c, q, 0, c.ntiles, true, ab, al, kscale, S, logZ);
        if (!fail) {
            final_sum = final_mass<P>(q, lane, c.T, ab, al);
            if (!(final_sum > 0.0)) fail = true;
        }
    }
    if (!fail) {
        double bb[P], bl[P];
#pragma unroll
        for (int j = 0; j < P; ++j) bb[j] = bl[j] = 0.0;
        // virtual state after the last frame: unit mass "beyond" the final blank, so that frame T-1 gets
        // pre[L-1] = pre[L-2] = 1 (ctc_fast.pyx:78-83) from the ordinary recurrence
#pragma unroll
        for (int j = 0; j < P; ++j)
            if (lane * P + j == q.nlab) bb[j] = 1.0;
        int kscale = 0;
        fail = !beta_tiles<P, true, TH>(a, c, q, c.ntiles - 1, 0, !short_utt, bb, bl, kscale);
    }
    if (fail) zero_rows(a, c.gbase, 0, c.T, lane, 32);   // the reference returns the zero-initialised grad on its failure path
    zero_rows(a, c.gbase, max(c.T, 0), a.Tmax, lane, 32);  // padded frames carry no gradient
    const float lz = warp_sum(logZ);
    if (lane == 0) write_loss(a, u, short_utt, fail, final_sum, S, lz);
}

// Throughput shape without the spill (CTCB_CTC=ckpt): ONE WARP per utterance, and the trellis never leaves the SM.  The
// spill of the kernel above costs 2 x T x 64P doubles of HBM traffic per utterance -- more than the activations and the
// gradient together (ncu, C1 shape, B = 8192: 2.78 GB moved for 0.81 GB of algorithmic bytes).  Here pass A runs alpha over
// all frames and keeps only a CHECKPOINT per 8-frame tile (the scaled states entering the tile and the pending power of
// two; the last tile's stays in registers).  Pass B walks the tiles backwards: alpha of the tile's frames is recomputed
// from its checkpoint into shared memory (bit-identical: same operations on the same values), then beta runs backwards
// over the tile, combining each frame with its alpha row into occupancies, and the gradient rows of the tile are written.
// HBM sees the activations twice (less the last tile, still in shared memory when pass B starts), the gradient once and
// 64P doubles per 8 frames: 1.29 GB at the shape above (1.58 x algorithmic).  The price is one more alpha frame per frame
// (519 M instead of 470 M warp instructions) and an 8 KB alpha tile per warp (26 resident warps instead of 32); the kernel
// is bound by instruction issue and chain latency, not by HBM, so it is the slower of the two and not the default.
constexpr int TC = 8;   // frames per tile of the checkpoint kernel: 8 KB of shared memory per warp -> 26 warps per SM

template <int P>
__global__ void __launch_bounds__(64, P == 1 ? 13 : 4) ctc_ckpt_kernel(CtcArgs a) {
    extern __shared__ double2 smem_ck[];
    constexpr int LP = 64 * P;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int u = blockIdx.x * (blockDim.x >> 5) + wib;
    if (u >= a.B) return;
    const size_t tile_f = (size_t)TC * a.Kp;
    float *wbase = reinterpret_cast<float *>(smem_ck) + (size_t)wib * ((size_t)TC * LP * 2 + 2 * tile_f);
    double2 *atile = reinterpret_cast<double2 *>(wbase);      // alpha-tilde of one tile: [TT][32 lanes][P] (blank, label)
    WarpCtx c;
    c.nbuf = 1; c.lane = lane;
    c.te0 = wbase + (size_t)TC * LP * 2;
    c.tg = c.te0 + tile_f;
    c.T = min(a.Tlen[u], a.Tmax);
    c.ntiles = (c.T + TC - 1) / TC;
    c.base = a.acts + (int64_t)u * a.us;
    c.gbase = a.grad + (int64_t)u * a.us;
    c.wsu = reinterpret_cast<double *>(a.ws) + (int64_t)u * a.ws_utt;
    int *ckk = reinterpret_cast<int *>(c.wsu + (int64_t)a.Tmax * LP);   // power of two pending at the start of each tile
    const int lo = a.loff[u];
    LaneLabels<P> q;
    q.load(a, lo, a.loff[u + 1] - lo, lane);
    init_tile_pads(a, c.te0, 1, lane, TC);
    zero_occupancy_tile(a, c.tg, lane, TC);
    __syncwarp();
    float *te = c.te0;
    unsigned *tgu = reinterpret_cast<unsigned *>(c.tg);
    const int T = c.T;

    const bool short_utt = (T < q.nlab);
    bool fail = (T <= 0);
    float logZ = 0.f, Zlast = 1.f;
    int S = 0, have = -1;       // have: the tile whose e-values are in shared memory
    double final_sum = 1.0;
    double ck_b[P], ck_l[P];    // checkpoint of the last tile
    int ck_k = 0;
#pragma unroll
    for (int j = 0; j < P; ++j) ck_b[j] = ck_l[j] = 0.0;

    // ---- pass A: alpha over all frames; loss and checkpoints
    if (!short_utt && !fail) {
        double ab[P], al[P];
#pragma unroll
        for (int j = 0; j < P; ++j) ab[j] = al[j] = 0.0;
        if (lane == 0) ab[0] = 1.0;     // virtual state before frame 0 (see ctc_warp_kernel)
        int kscale = 0;
        for (int tile = 0; tile < c.ntiles && !fail; ++tile) {
            const int t0 = tile * TC;
            prefetch_tile_l2(a, c.base, t0 + TC, T, lane, TC);
            issue_tile(a, c.base, t0, T, te, lane, TC);
            cp_async_wait<0>();
            __syncwarp();
            const float Z = tile_stats<TC>(a, t0, T, te, lane);
            if (lane < TC) logZ += logf(Z);
            Zlast = Z;
            have = tile;
            ck_k = kscale;
#pragma unroll
            for (int j = 0; j < P; ++j) { ck_b[j] = ab[j]; ck_l[j] = al[j]; }
            if (tile + 1 < c.ntiles) {
                double2 *crow = reinterpret_cast<double2 *>(c.wsu + (int64_t)tile * LP) + lane * P;
#pragma unroll
                for (int j = 0; j < P; ++j) crow[j] = make_double2(ab[j], al[j]);
                if (lane == 0) ckk[tile] = kscale;
            }
            const int rmax = min(TC, T - t0);
            const float *row = te;
            if (alpha_window(q.L, T, t0 + rmax - 1)) {
                for (int r = 0; r < rmax; ++r, row += a.Kp)
                    if (!alpha_frame<P, true>(q, row, a.blank, lane, T, t0 + r, ab, al, kscale, S)) { fail = true; break; }
            } else {
                for (int r = 0; r < rmax; ++r, row += a.Kp)
                    if (!alpha_frame<P, false>(q, row, a.blank, lane, T, t0 + r, ab, al, kscale, S)) { fail = true; break; }
            }
        }
        if (!fail) {
            final_sum = final_mass<P>(q, lane, T, ab, al);
            if (!(final_sum > 0.0)) fail = true;
        }
    }
    // ---- pass B: tiles backwards -- alpha again from the checkpoint, beta, occupancies, gradient
    if (!fail) {
        double bb[P], bl[P];
#pragma unroll
        for (int j = 0; j < P; ++j) bb[j] = bl[j] = 0.0;
#pragma unroll
        for (int j = 0; j < P; ++j)
            if (lane * P + j == q.nlab) bb[j] = 1.0;     // virtual state after the last frame (see ctc_warp_kernel)
        int kb = 0;
        for (int tile = c.ntiles - 1; tile >= 0 && !fail; --tile) {
            const int t0 = tile * TC;
            const int rmax = min(TC, T - t0);
            float Z = Zlast;
            prefetch_tile_l2(a, c.base, t0 - TC, T, lane, TC);
            if (tile != have) {
                issue_tile(a, c.base, t0, T, te, lane, TC);
                cp_async_wait<0>();
                __syncwarp();
                Z = tile_stats<TC>(a, t0, T, te, lane);
            }
            if (!short_utt) {
                double ab[P], al[P];
                int ka = ck_k, Sd = 0;
                if (tile == c.ntiles - 1) {
#pragma unroll
                    for (int j = 0; j < P; ++j) { ab[j] = ck_b[j]; al[j] = ck_l[j]; }
                } else {
                    const double2 *crow = reinterpret_cast<const double2 *>(c.wsu + (int64_t)tile * LP) + lane * P;
#pragma unroll
                    for (int j = 0; j < P; ++j) { const double2 v = crow[j]; ab[j] = v.x; al[j] = v.y; }
                    ka = ckk[tile];
                }
                const bool awin = alpha_window(q.L, T, t0 + rmax - 1), bwin = beta_window(q.L, t0);
                {
                    const float *row = te;
                    double2 *arow = atile + (size_t)lane * P;
                    for (int r = 0; r < rmax; ++r, row += a.Kp, arow += 32 * P) {
                        if (awin) alpha_frame<P, true>(q, row, a.blank, lane, T, t0 + r, ab, al, ka, Sd);
                        else alpha_frame<P, false>(q, row, a.blank, lane, T, t0 + r, ab, al, ka, Sd);
#pragma unroll
                        for (int j = 0; j < P; ++j) arow[j] = make_double2(ab[j], al[j]);
                    }
                }
                {
                    const float *row = te + (rmax - 1) * a.Kp;
                    unsigned *grow = tgu + (rmax - 1) * a.Kp;
                    const double2 *arow = atile + ((size_t)(rmax - 1) * 32 + lane) * P;
                    for (int r = rmax - 1; r >= 0; --r, row -= a.Kp, grow -= a.Kp, arow -= 32 * P) {
                        double pb[P], pll[P], xb[P], xl[P];
                        const bool ok = bwin ? beta_frame<P, true>(q, row, a.blank, lane, t0 + r, bb, bl, kb, pb, pll)
                                             : beta_frame<P, false>(q, row, a.blank, lane, t0 + r, bb, bl, kb, pb, pll);
                        if (!ok) { fail = true; break; }
#pragma unroll
                        for (int j = 0; j < P; ++j) { const double2 v = arow[j]; xb[j] = v.x; xl[j] = v.y; }
                        occupancy_frame<P>(q, a.blank, lane, xb, xl, pb, pll, grow);
                    }
                }
            }
            __syncwarp();
            if (!fail) tile_epilogue(a, c, t0, rmax, te, tgu, 1.f / Z);
        }
    }
    if (fail) zero_rows(a, c.gbase, 0, T, lane, 32);     // the reference returns the zero-initialised grad on its failure path
    zero_rows(a, c.gbase, max(T, 0), a.Tmax, lane, 32);  // padded frames carry no gradient
    const float lz = warp_sum(logZ);
    if (lane == 0) write_loss(a, u, short_utt, fail, final_sum, S, lz);
}

// Latency shape (few utterances, e.g. the 32 of a training step): TWO WARPS per utterance meet in the middle.
// Warp 0 runs alpha forward, warp 1 beta backward, at the same time, each spilling its half of the trellis; after
// one barrier warp 0 continues alpha through the second half combining with the stored beta sums, warp 1 continues
// beta through the first half combining with the stored alpha: T dependent frame steps instead of 2T.
template <int P>
__global__ void __launch_bounds__(64) ctc_pair_kernel(CtcArgs a) {
    extern __shared__ float smem[];
    __shared__ int fail_flag;
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    const int u = blockIdx.x;
    WarpCtx c;
    c.nbuf = 2; c.lane = lane;
    c.te0 = smem + (size_t)role * 3 * TT * a.Kp;
    c.tg = c.te0 + 2 * TT * a.Kp;
    c.T = min(a.Tlen[u], a.Tmax);
    c.ntiles = (c.T + TT - 1) / TT;
    c.base = a.acts + (int64_t)u * a.us;
    c.gbase = a.grad + (int64_t)u * a.us;
    c.wsu = reinterpret_cast<double *>(a.ws) + (int64_t)u * a.ws_utt;
    const int lo = a.loff[u];
    LaneLabels<P> q;
    q.load(a, lo, a.loff[u + 1] - lo, lane);
    init_tile_pads(a, c.te0, 2, lane);
    zero_occupancy_tile(a, c.tg, lane);
    const bool short_utt = (c.T < q.nlab);
    const bool recur = !short_utt && c.T > 0;
    const int mid = c.ntiles / 2;          // tiles [0, mid): alpha spilled, beta combines; [mid, ntiles): the reverse
    if (threadIdx.x == 0) fail_flag = (c.T <= 0) ? 1 : 0;
    __syncthreads();

    float logZ = 0.f;
    int S = 0, kscale = 0;
    double x0[P], x1[P];     // (ab, al) in warp 0, (bb, bl) in warp 1
#pragma unroll
    for (int j = 0; j < P; ++j) x0[j] = x1[j] = 0.0;
    bool ok = true;
    if (role == 0) {
        if (lane == 0) x0[0] = 1.0;
        if (recur) ok = alpha_tiles<P, false>(a, c, q, 0, mid, true, x0, x1, kscale, S, logZ);
    } else {
#pragma unroll
        for (int j = 0; j < P; ++j)
            if (lane * P + j == q.nlab) x0[j] = 1.0;
        if (recur) ok = beta_tiles<P, false>(a, c, q, c.ntiles - 1, mid, true, x0, x1, kscale);
    }
    if (!ok && lane == 0) fail_flag = 1;
    __threadfence_block();
    __syncthreads();                       // both halves of the trellis are in the workspace
    double final_sum = 1.0;
    if (!fail_flag) {
        if (role == 0) {
            ok = alpha_tiles<P, true>(a, c, q, mid, c.ntiles, recur, x0, x1, kscale, S, logZ);
            if (ok && recur) {
                final_sum = final_mass<P>(q, lane, c.T, x0, x1);
                if (!(final_sum > 0.0)) ok = false;
            }
        } else {
            ok = beta_tiles<P, true>(a, c, q, mid - 1, 0, recur, x0, x1, kscale);
        }
        if (!ok && lane == 0) fail_flag = 1;
    }
    __syncthreads();
    const bool fail = fail_flag != 0;
    if (fail) zero_rows(a, c.gbase, 0, c.T, threadIdx.x, 64);
    zero_rows(a, c.gbase, max(c.T, 0), a.Tmax, threadIdx.x, 64);
    if (role == 0) {
        // log Z of the first-half rows was only seen by warp 1's tiles: warp 0 walked tiles [0, mid) in its first
        // phase and [mid, ntiles) in its second, i.e. every row exactly once
        const float lz = warp_sum(logZ);
        if (lane == 0) write_loss(a, u, short_utt, fail, final_sum, S, lz);
    }
}

// Gradient of one time tile from the spilled trellis: alpha-tilde (plane A) and the beta pre-emission sums (plane B) of
// every frame are in the workspace, so the frames are independent -- any warp can take any tile.
template <int P>
__device__ __forceinline__ void combine_tile(const CtcArgs &a, const WarpCtx &c, const LaneLabels<P> &q, int tile, bool recur,
                                             const double *wsA, const double *wsB) {
    constexpr int LP = 64 * P;
    const int lane = c.lane, T = c.T, t0 = tile * TT;
    float *te = c.te0;
    issue_tile(a, c.base, t0, T, te, lane);
    cp_async_wait<0>();
    __syncwarp();
    const float Z = tile_stats(a, t0, T, te, lane);
    unsigned *tgu = reinterpret_cast<unsigned *>(c.tg);
    const int rmax = min(TT, T - t0);
    if (recur) {
        for (int r = 0; r < rmax; ++r) {
            const int t = t0 + r;
            const double2 *arow = reinterpret_cast<const double2 *>(wsA + (int64_t)t * LP) + lane * P;
            const double2 *prow = reinterpret_cast<const double2 *>(wsB + (int64_t)t * LP) + lane * P;
            double xb[P], xl[P], pb[P], pll[P];
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const double2 av = arow[j], pv = prow[j];
                xb[j] = av.x; xl[j] = av.y; pb[j] = pv.x; pll[j] = pv.y;
            }
            occupancy_frame<P>(q, a.blank, lane, xb, xl, pb, pll, tgu + r * a.Kp);
        }
    }
    __syncwarp();
    tile_epilogue(a, c, t0, rmax, te, tgu, 1.f / Z);
}

// Latency shape, round 2 (a training step's minibatch: at most one utterance per SM): the serial chain carries ONLY the
// recurrences.  Warp 0 runs alpha forward over all frames and warp 1 beta backward over all frames, at the same time,
// each spilling its scaled states (two workspace planes); then the frames are independent and ALL warps of the CTA turn
// the two planes into occupancies and the gradient, one time tile each.  Against the meet-in-the-middle kernel above
// (which computes occupancies and gradient on the two serial warps) the chain loses ~40 % of its instructions:
// 0.153 -> 0.1 ms for the 32 utterances of the C2 step.
template <int P>
__global__ void __launch_bounds__(256) ctc_par_kernel(CtcArgs a, int NW) {
    extern __shared__ float smem[];
    __shared__ int fail_flag, S_s;
    __shared__ float lz_s;
    __shared__ double fin_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int u = blockIdx.x;
    WarpCtx c;
    c.lane = lane;
    c.T = min(a.Tlen[u], a.Tmax);
    c.ntiles = (c.T + TT - 1) / TT;
    c.base = a.acts + (int64_t)u * a.us;
    c.gbase = a.grad + (int64_t)u * a.us;
    double *wsA = reinterpret_cast<double *>(a.ws) + (int64_t)u * 2 * a.ws_utt;
    double *wsB = wsA + a.ws_utt;
    const int lo = a.loff[u];
    LaneLabels<P> q;
    q.load(a, lo, a.loff[u + 1] - lo, lane);
    const bool short_utt = (c.T < q.nlab);
    const bool recur = !short_utt && c.T > 0;
    if (threadIdx.x == 0) { fail_flag = (c.T <= 0) ? 1 : 0; S_s = 0; lz_s = 0.f; fin_s = 1.0; }
    __syncthreads();
    const size_t tile_f = (size_t)TT * a.Kp;
    if (warp < 2) {
        // ---- phase 1: the two recurrences, nothing else on the chain
        c.nbuf = 2;
        c.te0 = smem + (size_t)warp * 2 * tile_f;
        c.tg = nullptr;
        c.wsu = (warp == 0) ? wsA : wsB;
        init_tile_pads(a, c.te0, 2, lane);
        __syncwarp();
        double x0[P], x1[P];
#pragma unroll
        for (int j = 0; j < P; ++j) x0[j] = x1[j] = 0.0;
        int kscale = 0, S = 0;
        float logZ = 0.f;
        bool ok = true;
        if (warp == 0) {
            if (lane == 0) x0[0] = 1.0;
            ok = alpha_tiles<P, false>(a, c, q, 0, c.ntiles, recur, x0, x1, kscale, S, logZ);
            double final_sum = 1.0;
            if (ok && recur) {
                final_sum = final_mass<P>(q, lane, c.T, x0, x1);
                if (!(final_sum > 0.0)) ok = false;
            }
            const float lz = warp_sum(logZ);
            if (lane == 0) { lz_s = lz; fin_s = final_sum; S_s = S; }
        } else {
#pragma unroll
            for (int j = 0; j < P; ++j)
                if (lane * P + j == q.nlab) x0[j] = 1.0;
            if (recur) ok = beta_tiles<P, false>(a, c, q, c.ntiles - 1, 0, true, x0, x1, kscale);
        }
        if (!ok && lane == 0) fail_flag = 1;
    }
    __threadfence_block();
    __syncthreads();                       // both planes of the trellis are in the workspace
    // ---- phase 2: frames are independent now -- every warp takes time tiles
    if (!fail_flag && warp < NW) {
        c.nbuf = 1;
        c.te0 = smem + (size_t)warp * 2 * tile_f;
        c.tg = c.te0 + tile_f;
        zero_occupancy_tile(a, c.tg, lane);
        __syncwarp();
        for (int tile = warp; tile < c.ntiles; tile += NW) combine_tile<P>(a, c, q, tile, recur, wsA, wsB);
    }
    __syncthreads();
    const bool fail = fail_flag != 0;
    if (fail) zero_rows(a, c.gbase, 0, c.T, threadIdx.x, blockDim.x);
    zero_rows(a, c.gbase, max(c.T, 0), a.Tmax, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) write_loss(a, u, short_utt, fail, fin_s, S_s, lz_s);
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

static int row_pitch(int K, bool vec2) {
    if (!vec2) return (K + 1) | 1;
    int kp = K + 1;
    while (kp % 4 != 2) ++kp;
    return kp;
}

// trellis rows [Tmax][64P] + one 32-bit word per 8-frame tile (checkpoint kernel), kept even: rows are read as double2
static int64_t ws_doubles_per_utt(int Tmax, int P) {
    const int64_t tiles = (Tmax + TT - 1) / TT;
    return (int64_t)Tmax * 64 * P + ((tiles + 1) & ~(int64_t)1);
}

static int pairs_per_lane(int max_labels) {
    const int npairs = max_labels + 1;
    for (int p = 1; p <= 16; p *= 2)
        if (npairs <= 32 * p) return p;
    return 0;
}

}  // namespace ctcb

using namespace ctcb;

static int g_ctc_shape = -1;    // 0 automatic, 1 warp, 2 pair, 3 par, 4 ckpt; -1: read CTCB_CTC first

extern "C" int ctcb_debug_set_ctc_kernel(int shape) {
    if (shape < 0 || shape > 4) return set_error(CTCB_EINVAL, "ctcb_debug_set_ctc_kernel: shape %d not in 0..4", shape);
    g_ctc_shape = shape;
    return CTCB_OK;
}

extern "C" size_t ctcb_ctc_workspace_bytes(int B, int Tmax, int max_labels) {
    const int P = pairs_per_lane(max_labels);
    if (P == 0 || B <= 0 || Tmax <= 0) return 0;
    // small batches take the three-phase latency kernel, which spills alpha AND beta for every frame (two planes)
    // + one word per 16-frame tile for the checkpoint kernel (the pending power of two)
    return (size_t)B * (size_t)ws_doubles_per_utt(Tmax, P) * sizeof(double) * (B <= 256 ? 2 : 1);
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
    // 8-byte copies and stores need even K, even strides, 8-byte aligned bases and an even shared row pitch
    a.vec2 = (K % 2 == 0) && (utt_stride % 2 == 0) && (frame_stride % 2 == 0) && (((uintptr_t)acts) % 8 == 0) &&
             (((uintptr_t)grad_out) % 8 == 0);
    // shared row pitch: at least K + 1 (class index K of every row is the zero slot of init_tile_pads), and odd (4-byte
    // copies) or = 2 mod 4 (8-byte copies) so that the 16 rows of a tile start in different banks for the statistics pass
    a.Kp = row_pitch(K, a.vec2 != 0);
    a.grad = grad_out; a.nll = nll_out; a.skip = skip_out;
    a.ws = (float *)workspace; a.ws_utt = ws_doubles_per_utt(Tmax, P);

    cudaStream_t st = (cudaStream_t)stream;
    // which kernel: chosen from the batch below; CTCB_CTC=warp|pair|par|ckpt or ctcb_debug_set_ctc_kernel force one
    // (tests, measurements)
    if (g_ctc_shape < 0) {
        const char *e = getenv("CTCB_CTC");
        g_ctc_shape = !e ? 0 : (e[0] == 'w' ? 1 : (e[0] == 'c' ? 4 : (e[0] == 'p' && e[1] == 'a' && e[2] == 'r' ? 3 : (e[0] == 'p' ? 2 : 0))));
    }
    const int shape_env = g_ctc_shape;
    {   // at most one utterance per SM (a training step): recurrences on two warps, gradient on all eight
        const size_t tile_b = (size_t)TT * a.Kp * sizeof(float);
        int NW = 8;
        while (NW > 2 && (size_t)2 * NW * tile_b > 160 * 1024) NW >>= 1;
        const size_t smem = (size_t)2 * NW * tile_b;
        const bool par = (shape_env == 3 && B <= 256) || (shape_env == 0 && B <= num_sms());
        if (par && smem <= 200 * 1024) {
            a.nbuf = 2;
#define LAUNCH_PAR(PP)                                                                                 \
    case PP: {                                                                                         \
        CTCB_CUDA_CHECK(cudaFuncSetAttribute(ctc_par_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        ctc_par_kernel<PP><<<B, 256, smem, st>>>(a, NW);                                               \
        break;                                                                                         \
    }
            switch (P) {
                LAUNCH_PAR(1) LAUNCH_PAR(2) LAUNCH_PAR(4) LAUNCH_PAR(8) LAUNCH_PAR(16)
                default: return set_error(CTCB_EINVAL, "bad P");
            }
#undef LAUNCH_PAR
            CTCB_LAUNCH_CHECK();
            return CTCB_OK;
        }
    }
    // the two-warp shape wins as long as its CTAs (one per utterance) fit the SMs in one wave: below that the
    // one-warp shape leaves the schedulers short of warps (C5 sweep: 1.6 vs 3.1 ms at T=2000, B=292 vs 604)
    const size_t smem_pair = (size_t)2 * 3 * TT * a.Kp * sizeof(float);
    int pair_per_sm = (int)((200 * 1024) / (smem_pair ? smem_pair : 1));
    if (pair_per_sm > 8) pair_per_sm = 8;
    const bool pair = (shape_env == 2) || (shape_env == 0 && pair_per_sm >= 1 && B <= pair_per_sm * num_sms());
    if (pair) {
        const size_t smem = smem_pair;
        if (smem > 200 * 1024)
            return set_error(CTCB_EINVAL, "ctcb_ctc_loss_grad_f32: K=%d too large for the shared-memory tile", K);
        a.nbuf = 2;
#define LAUNCH_PAIR(PP)                                                                                \
    case PP: {                                                                                         \
        CTCB_CUDA_CHECK(cudaFuncSetAttribute(ctc_pair_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        ctc_pair_kernel<PP><<<B, 64, smem, st>>>(a);                                                   \
        break;                                                                                         \
    }
        switch (P) {
            LAUNCH_PAIR(1) LAUNCH_PAIR(2) LAUNCH_PAIR(4) LAUNCH_PAIR(8) LAUNCH_PAIR(16)
            default: return set_error(CTCB_EINVAL, "bad P");
        }
#undef LAUNCH_PAIR
        CTCB_LAUNCH_CHECK();
        return CTCB_OK;
    }
    // CTCB_CTC=ckpt: one warp per utterance, trellis checkpointed on chip instead of spilled (1.6x instead of 3.4x the
    // algorithmic HBM bytes, but ~10 % more instructions on a kernel that is bound by instruction issue, not by HBM:
    // 0.77 vs 0.62 ms at B = 8192 x C1 -- so it is not the default).  One warp needs an 8-frame alpha tile and two
    // activation tiles of shared memory; the block size is the smallest that reaches the most resident warps per SM.
    {
        const size_t per_warp = (size_t)TC * 64 * P * sizeof(double) + (size_t)2 * TC * a.Kp * sizeof(float);
        int best_wpb = 0, best_warps = 0;
        for (int w = 1; w <= 2; ++w) {      // (the kernel is compiled for blocks of at most 2 warps)
            const size_t blk = per_warp * w;
            if (blk > 200 * 1024) break;
            int nblk = (int)((size_t)(228 * 1024) / (blk + 1024));
            if (nblk > 32) nblk = 32;
            if (nblk * w > 64) nblk = 64 / w;
            if (nblk * w > best_warps) { best_warps = nblk * w; best_wpb = w; }
        }
        const bool ckpt = (shape_env == 4 && P <= 4);
        if (ckpt && best_warps >= 8) {
            const int wpb = best_wpb;
            const size_t smem = per_warp * wpb;
            a.nbuf = 1;
#define LAUNCH_CK(PP)                                                                                  \
    case PP: {                                                                                         \
        CTCB_CUDA_CHECK(cudaFuncSetAttribute(ctc_ckpt_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        ctc_ckpt_kernel<PP><<<(B + wpb - 1) / wpb, wpb * 32, smem, st>>>(a);                           \
        break;                                                                                         \
    }
            switch (P) {
                LAUNCH_CK(1) LAUNCH_CK(2) LAUNCH_CK(4)
                default: return set_error(CTCB_EINVAL, "CTCB_CTC=ckpt: label sequences longer than 127 take the spilling kernel");
            }
#undef LAUNCH_CK
            CTCB_LAUNCH_CHECK();
            return CTCB_OK;
        }
    }
    // many utterances: one warp each; favour occupancy (one e tile) when the SMs are full, else the latency of
    // each warp (double-buffered tiles)
    a.nbuf = (B >= 16 * num_sms()) ? 1 : 2;
    const int th = (a.nbuf == 1 && P <= 2) ? 8 : TT;
    const size_t per_warp = (size_t)(a.nbuf + 1) * th * a.Kp * sizeof(float);
    int wpb = 8;
    while (wpb > 1 && per_warp * wpb > 100 * 1024) wpb >>= 1;
    while (wpb > 1 && (B + wpb - 1) / wpb < 2 * num_sms()) wpb >>= 1;
    const size_t smem = per_warp * wpb;
    if (smem > 200 * 1024)
        return set_error(CTCB_EINVAL, "ctcb_ctc_loss_grad_f32: K=%d too large for the shared-memory tile", K);
    const int grid = (B + wpb - 1) / wpb;
#define LAUNCH_PT(PP, TH_)                                                                             \
    {                                                                                                  \
        CTCB_CUDA_CHECK(cudaFuncSetAttribute(ctc_warp_kernel<PP, TH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        ctc_warp_kernel<PP, TH_><<<grid, wpb * 32, smem, st>>>(a);                                     \
    }
#define LAUNCH_P(PP)                                                                                   \
    case PP: {                                                                                         \
        if (PP <= 2 && th == 8) LAUNCH_PT((PP <= 2 ? PP : 1), 8) else LAUNCH_PT(PP, TT)                \
        break;                                                                                         \
    }
    switch (P) {
        LAUNCH_P(1) LAUNCH_P(2) LAUNCH_P(4) LAUNCH_P(8) LAUNCH_P(16)
        default: return set_error(CTCB_EINVAL, "bad P");
    }
#undef LAUNCH_PT
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
