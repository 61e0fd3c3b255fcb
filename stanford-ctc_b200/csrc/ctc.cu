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

// alpha/beta live in registers in FLOAT64 (the reference's own arithmetic, ctc_fast.pyx:23-37): float32
// cannot hold the product of the alpha and beta tails, which is what the gradient is made of.  Instead of
// dividing by the frame normaliser every frame (a warp reduction on the serial chain) the state is rescaled
// by a power of two derived from the PREVIOUS frame's largest exponent (one REDUX.MAX off the chain); the
// accumulated exponent goes into the loss.
//
// Per-lane view of one utterance: pairs i = lane*P + j  (blank s = 2i, label s = 2i+1).
template <int P>
struct LaneLabels {
    int lab[P];
    bool allow_a[P], allow_b[P];
    int nlab, L;
    __device__ __forceinline__ void load(const CtcArgs &a, int lo, int nl, int lane) {
        nlab = nl; L = 2 * nl + 1;
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int i = lane * P + j;
            lab[j] = (i < nlab) ? a.labels[lo + i] : -1;
            const int lprev = (i >= 1 && i < nlab) ? a.labels[lo + i - 1] : -1;
            const int lnext = (i + 1 < nlab) ? a.labels[lo + i + 1] : -1;
            allow_a[j] = (i >= 1 && i < nlab && lab[j] != lprev);       // ctc_fast.pyx:64-68
            allow_b[j] = (i + 1 < nlab && lab[j] != lnext);             // ctc_fast.pyx:104-108
        }
    }
};

// One frame of the alpha recurrence (:49-76) on e-values `row`.  ab/al: blank/label states of the previous
// frame in, of frame t out (scaled by 2^kscale of the previous frame).  Returns false when all mass is gone.
template <int P>
__device__ __forceinline__ bool alpha_frame(const LaneLabels<P> &q, const float *row, int blank, int lane, int T, int t,
                                            double (&ab)[P], double (&al)[P], int &kscale, int &S) {
    const double eb = (double)row[blank];
    int start = 2 * (T - t);
    start = (q.L <= start || t == 0) ? 0 : q.L - start;   // the reference sets frame 0 without a window (:42-47)
    double pl = __shfl_up_sync(0xffffffffu, al[P - 1], 1);
    if (lane == 0) pl = 0.0;
    const double sc = pow2_from_field(1023 + kscale);
    double mx = 0.0;
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int i = lane * P + j;
        const double el = (q.lab[j] >= 0) ? (double)row[q.lab[j]] : 0.0;
        double b = (ab[j] + pl) * eb;
        double l = (al[j] + ab[j] + (q.allow_a[j] ? pl : 0.0)) * el;
        if (i > q.nlab) b = 0.0;
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
    if (emax == 0) return false;            // all mass gone: ZeroDivisionError in :70-76
    kscale = 1023 - emax;
    return true;
}

// One frame of the beta recurrence (:85-114).  bb/bl: states of frame t+1 in, of frame t out.  pb/pll receive the
// (scaled) PRE-emission sums of frame t: beta[s,t] = pre[s] * p[lab(s),t], so alpha*beta/p = alpha*pre.
template <int P>
__device__ __forceinline__ bool beta_frame(const LaneLabels<P> &q, const float *row, int blank, int lane, int t,
                                           double (&bb)[P], double (&bl)[P], int &kscale, double (&pb)[P],
                                           double (&pll)[P]) {
    const double eb = (double)row[blank];
    const int end = min(2 * t + 2, q.L);
    double nxb = __shfl_down_sync(0xffffffffu, bb[0], 1);
    double nxl = __shfl_down_sync(0xffffffffu, bl[0], 1);
    if (lane == 31) nxb = nxl = 0.0;
    const double sc = pow2_from_field(1023 + kscale);
    double nb[P], mx = 0.0;
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int i = lane * P + j;
        const double el = (q.lab[j] >= 0) ? (double)row[q.lab[j]] : 0.0;
        const double b1 = (j + 1 < P) ? bb[(j + 1 < P) ? j + 1 : j] : nxb;
        const double l1 = (j + 1 < P) ? bl[(j + 1 < P) ? j + 1 : j] : nxl;
        double xb = bb[j] + bl[j];
        double xl = bl[j] + b1 + (q.allow_b[j] ? l1 : 0.0);
        if (i > q.nlab) xb = 0.0;
        if (q.lab[j] < 0) xl = 0.0;
        if (end < q.L) {                    // warp-uniform: only the first |l| frames prune
            if (2 * i >= end) xb = 0.0;
            if (2 * i + 1 >= end) xl = 0.0;
        }
        xb *= sc;
        xl *= sc;
        pb[j] = xb;
        pll[j] = xl;
        nb[j] = xb * eb;
        const double lnew = xl * el;
        mx = fmax(mx, fmax(nb[j], lnew));
        bl[j] = lnew;                       // old bl[j] no longer needed by later j
    }
#pragma unroll
    for (int j = 0; j < P; ++j) bb[j] = nb[j];
    const int emax = __reduce_max_sync(0xffffffffu, dexp_field(mx));
    kscale = 1023 - emax;
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
        if (g > 0.f) atomicAdd(grow + q.lab[j], (unsigned)(g * GFIX + 0.5f));
    }
    if (lane == 0) atomicAdd(grow + blank, (unsigned)((float)wbsum * winv * GFIX + 0.5f));
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
__device__ __forceinline__ float *fetch_tile(const CtcArgs &a, const WarpCtx &c, int tile, int next) {
    float *te = c.te0 + ((c.nbuf == 2) ? (tile & 1) : 0) * TT * a.Kp;
    if (c.nbuf == 2 && next >= 0) {
        issue_tile(a, c.base, next * TT, c.T, c.te0 + (next & 1) * TT * a.Kp, c.lane);
        cp_async_wait<1>();
    } else {
        if (c.nbuf == 1) issue_tile(a, c.base, tile * TT, c.T, te, c.lane);
        cp_async_wait<0>();
    }
    __syncwarp();
    return te;
}

// grad = p - occupancy (:139-145) for the rmax frames of a tile, coalesced row stores
__device__ __forceinline__ void tile_epilogue(const CtcArgs &a, const WarpCtx &c, int t0, int rmax, const float *te,
                                              const unsigned *tgu, float my_Zinv) {
    for (int r = 0; r < rmax; ++r) {
        const float zinv = __shfl_sync(0xffffffffu, my_Zinv, r);
        float *orow = c.gbase + (int64_t)(t0 + r) * a.fs;
        for (int k = c.lane; k < a.K; k += 32)
            orow[k] = te[r * a.Kp + k] * zinv - (float)tgu[r * a.Kp + k] * (1.f / GFIX);
    }
    __syncwarp();
}

// alpha over tiles [tile_from, tile_to), ascending (:42-76).  COMBINE = false: the scaled states of every frame are
// spilled to the workspace for a later beta sweep.  COMBINE = true: the workspace already holds the beta
// pre-emission sums of these frames (stored by beta_tiles<P,false>); occupancies and gradient are produced here.
template <int P, bool COMBINE>
__device__ __forceinline__ bool alpha_tiles(const CtcArgs &a, const WarpCtx &c, const LaneLabels<P> &q, int tile_from,
                                            int tile_to, bool recur, double (&ab)[P], double (&al)[P], int &kscale,
                                            int &S, float &logZ) {
    constexpr int LP = 64 * P;
    const int lane = c.lane, T = c.T;
    if (c.nbuf == 2 && tile_from < tile_to) issue_tile(a, c.base, tile_from * TT, T, c.te0 + (tile_from & 1) * TT * a.Kp, lane);
    for (int tile = tile_from; tile < tile_to; ++tile) {
        const int t0 = tile * TT;
        const float *te = fetch_tile(a, c, tile, (tile + 1 < tile_to) ? tile + 1 : -1);
        const float Z = tile_stats(a, t0, T, const_cast<float *>(te), lane);
        if (lane < TT) logZ += logf(Z);
        const int rmax = min(TT, T - t0);
        unsigned *tgu = reinterpret_cast<unsigned *>(c.tg);
        if (COMBINE) {
            for (int idx = lane; idx < TT * a.Kp; idx += 32) tgu[idx] = 0u;
            __syncwarp();
        }
        if (recur) {
            double2 pn[P];      // beta pre-emission sums of the next frame to be processed (prefetched)
            if (COMBINE) {
                const double2 *prow = reinterpret_cast<const double2 *>(c.wsu + (int64_t)t0 * LP) + lane * P;
#pragma unroll
                for (int j = 0; j < P; ++j) pn[j] = prow[j];
            }
            for (int r = 0; r < rmax; ++r) {
                const int t = t0 + r;
                double2 pv[P];
                if (COMBINE) {
#pragma unroll
                    for (int j = 0; j < P; ++j) pv[j] = pn[j];
                    const double *pf = c.wsu + (int64_t)min(t + 4, T - 1) * LP + lane * 2 * P;
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
                    if (P > 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(pf + 16));
                    const double2 *prow = reinterpret_cast<const double2 *>(c.wsu + (int64_t)min(t + 1, T - 1) * LP) + lane * P;
#pragma unroll
                    for (int j = 0; j < P; ++j) pn[j] = prow[j];
                }
                if (!alpha_frame<P>(q, te + r * a.Kp, a.blank, lane, T, t, ab, al, kscale, S)) return false;
                if (COMBINE) {
                    double pb[P], pll[P];
#pragma unroll
                    for (int j = 0; j < P; ++j) { pb[j] = pv[j].x; pll[j] = pv[j].y; }
                    occupancy_frame<P>(q, a.blank, lane, ab, al, pb, pll, tgu + r * a.Kp);
                } else {
                    double2 *wrow = reinterpret_cast<double2 *>(c.wsu + (int64_t)t * LP) + lane * P;
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
template <int P, bool COMBINE>
__device__ __forceinline__ bool beta_tiles(const CtcArgs &a, const WarpCtx &c, const LaneLabels<P> &q, int tile_from,
                                           int tile_to, bool recur, double (&bb)[P], double (&bl)[P], int &kscale) {
    constexpr int LP = 64 * P;
    const int lane = c.lane, T = c.T;
    if (c.nbuf == 2 && tile_from >= tile_to) issue_tile(a, c.base, tile_from * TT, T, c.te0 + (tile_from & 1) * TT * a.Kp, lane);
    for (int tile = tile_from; tile >= tile_to; --tile) {
        const int t0 = tile * TT;
        const float *te = fetch_tile(a, c, tile, (tile > tile_to) ? tile - 1 : -1);
        const float Z = tile_stats(a, t0, T, const_cast<float *>(te), lane);
        unsigned *tgu = reinterpret_cast<unsigned *>(c.tg);
        if (COMBINE) {
            for (int idx = lane; idx < TT * a.Kp; idx += 32) tgu[idx] = 0u;
            __syncwarp();
        }
        const int rmax = min(TT, T - t0);
        if (recur) {
            double2 an[P];   // alpha-tilde of the next frame to be processed (prefetched)
            if (COMBINE) {
                const double2 *arow = reinterpret_cast<const double2 *>(c.wsu + (int64_t)(t0 + rmax - 1) * LP) + lane * P;
#pragma unroll
                for (int j = 0; j < P; ++j) an[j] = arow[j];
            }
            for (int r = rmax - 1; r >= 0; --r) {
                const int t = t0 + r;
                double2 av[P];
                if (COMBINE) {
#pragma unroll
                    for (int j = 0; j < P; ++j) av[j] = an[j];
                    // pull the alpha-tilde row needed 4 frames from now into L1 (the spill sits in L2/HBM), then
                    // load the row of the next frame (t-1, clamped: the value is unused at t = 0)
                    const double *pf = c.wsu + (int64_t)max(t - 4, 0) * LP + lane * 2 * P;
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
                    if (P > 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(pf + 16));
                    const double2 *arow = reinterpret_cast<const double2 *>(c.wsu + (int64_t)max(t - 1, 0) * LP) + lane * P;
#pragma unroll
                    for (int j = 0; j < P; ++j) an[j] = arow[j];
                }
                double pb[P], pll[P];
                if (!beta_frame<P>(q, te + r * a.Kp, a.blank, lane, t, bb, bl, kscale, pb, pll)) return false;
                if (COMBINE) {
                    double xb[P], xl[P];
#pragma unroll
                    for (int j = 0; j < P; ++j) { xb[j] = av[j].x; xl[j] = av[j].y; }
                    occupancy_frame<P>(q, a.blank, lane, xb, xl, pb, pll, tgu + r * a.Kp);
                } else {
                    double2 *wrow = reinterpret_cast<double2 *>(c.wsu + (int64_t)t * LP) + lane * P;
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
template <int P>
__global__ void __launch_bounds__(256) ctc_warp_kernel(CtcArgs a) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int u = blockIdx.x * (blockDim.x >> 5) + wib;
    if (u >= a.B) return;
    // nbuf = 2: two e tiles (the copy of tile n+1 overlaps the work on tile n; used for small batches, where one
    // warp has an SM almost to itself); nbuf = 1: one e tile, 1/3 less shared memory -> 24 instead of 16 warps per
    // SM, the other warps hide the copy (large batches).  Then the occupancy tile.
    WarpCtx c;
    c.nbuf = a.nbuf; c.lane = lane;
    c.te0 = smem + (size_t)wib * (a.nbuf + 1) * TT * a.Kp;
    c.tg = c.te0 + a.nbuf * TT * a.Kp;
    c.T = min(a.Tlen[u], a.Tmax);
    c.ntiles = (c.T + TT - 1) / TT;
    c.base = a.acts + (int64_t)u * a.us;
    c.gbase = a.grad + (int64_t)u * a.us;
    c.wsu = reinterpret_cast<double *>(a.ws) + (int64_t)u * a.ws_utt;
    const int lo = a.loff[u];
    LaneLabels<P> q;
    q.load(a, lo, a.loff[u + 1] - lo, lane);

    const bool short_utt = (c.T < q.nlab);
    bool fail = (c.T <= 0);
    float logZ = 0.f;     // lanes < TT accumulate log Z of the rows they own
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
        fail = !alpha_tiles<P, false>(a, c, q, 0, c.ntiles, true, ab, al, kscale, S, logZ);
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
        fail = !beta_tiles<P, true>(a, c, q, c.ntiles - 1, 0, !short_utt, bb, bl, kscale);
    }
    if (fail) zero_rows(a, c.gbase, 0, c.T, lane, 32);   // the reference returns the zero-initialised grad on its failure path
    zero_rows(a, c.gbase, max(c.T, 0), a.Tmax, lane, 32);  // padded frames carry no gradient
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
    for (int idx = lane; idx < TT * a.Kp; idx += 32) tgu[idx] = 0u;
    __syncwarp();
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
    // small batches take the three-phase latency kernel, which spills alpha AND beta for every frame (two planes)
    return (size_t)B * (size_t)Tmax * (size_t)(64 * P) * sizeof(double) * (B <= 256 ? 2 : 1);
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

    cudaStream_t st = (cudaStream_t)stream;
    // few utterances (a training step): two warps per utterance meet in the middle of the trellis, one CTA each;
    // CTCB_CTC=warp|pair forces a shape (tests)
    static int shape_env = -1;
    if (shape_env < 0) {
        const char *e = getenv("CTCB_CTC");
        shape_env = !e ? 0 : (e[0] == 'w' ? 1 : (e[0] == 'p' && e[1] == 'a' && e[2] == 'r' ? 3 : (e[0] == 'p' ? 2 : 0)));
    }
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
    // many utterances: one warp each; favour occupancy (one e tile) when the SMs are full, else the latency of
    // each warp (double-buffered tiles)
    a.nbuf = (B >= 16 * num_sms()) ? 1 : 2;
    const size_t per_warp = (size_t)(a.nbuf + 1) * TT * a.Kp * sizeof(float);
    int wpb = 8;
    while (wpb > 1 && per_warp * wpb > 100 * 1024) wpb >>= 1;
    while (wpb > 1 && (B + wpb - 1) / wpb < 2 * num_sms()) wpb >>= 1;
    const size_t smem = per_warp * wpb;
    if (smem > 200 * 1024)
        return set_error(CTCB_EINVAL, "ctcb_ctc_loss_grad_f32: K=%d too large for the shared-memory tile", K);
    const int grid = (B + wpb - 1) / wpb;
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
