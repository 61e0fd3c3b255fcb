/*
 * ctcb200.h -- C ABI of libctcb200.so, the B200 (sm_100a) compute backend for the
 * stanford-ctc training hot path (BRNN forward/backward + CTC alpha/beta + Nesterov SGD).
 *
 * Every entry point takes plain pointers and sizes (no torch types).  Unless a parameter is
 * documented as a HOST pointer it is a DEVICE pointer; all work is enqueued asynchronously on
 * the caller-supplied cudaStream_t (passed as void*).  Return value: 0 on success, a negative
 * CTCB_E* code otherwise, with a message retrievable from ctcb_last_error().  Numerical
 * infeasibility of one utterance (the reference's `skip=True`) is NOT an error: it is reported
 * per utterance in `skip_out`, as /root/reference/ctc_fast/ctc-loss/ctc_fast.pyx:147-149 does.
 *
 * The reference interface each entry point replaces is cited as (file:line) relative to
 * /root/reference/.
 */
#ifndef CTCB200_H
#define CTCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTCB_OK 0
#define CTCB_EINVAL (-1)   /* bad argument / unsupported shape */
#define CTCB_ECUDA (-2)    /* CUDA runtime error (see ctcb_last_error) */
#define CTCB_ENOMEM (-3)   /* caller workspace too small */

/* Longest label sequence per utterance the CTC kernel accepts (16 state pairs per lane, ctc.cu). */
#define CTCB_CTC_MAX_LABELS 511

/* ---- library ------------------------------------------------------------------------------ */
int ctcb_version(void);
const char *ctcb_last_error(void);
/* Number of kernels this library has launched since it was loaded (bench.py: gpu_launches). */
uint64_t ctcb_launch_count(void);
/* Per-phase device timing with CUDA events on the launching stream (bench.py roofline). When enabled,
 * every phase of ctcb_brnn_cost_and_grad and the optimiser entry points is bracketed by events;
 * ctcb_profile_report synchronises the device and writes a JSON object
 * {"phase": {"count": n, "total_ms": t}, ...} into buf, then clears the records. */
void ctcb_profile_enable(int on);
int ctcb_profile_report(char *buf, size_t cap);

/* ---- CTC loss + gradient (replaces ctc_fast.ctc_loss, ctc_fast/ctc-loss/ctc_fast.pyx:13-152;
 *      call site ctc_fast/nnets/brnnet.py:175, with the softmax of brnnet.py:161-168 fused) ----
 * acts      : activations of B utterances; element (u, t, k) at acts[u*utt_stride + t*frame_stride + k]
 *             (strides in elements).  is_prob=0: pre-softmax logits, the softmax is fused;
 *             is_prob=1: probabilities (the contract of ctc_fast.ctc_loss's `params`).
 * labels    : concatenated int32 label sequences; utterance u owns labels[label_off[u] .. label_off[u+1])
 * T_per_utt : int32[B] frame counts (<= Tmax).  Frames t >= T_per_utt[u] get zero gradient.
 * grad_out  : same indexing as acts; d(nll)/d(logit) = softmax - posterior occupancy
 *             (ctc_fast.pyx:139-145).  Zero for skipped utterances.
 * nll_out   : float32[B]  -log p(labels | acts)     (ctc_fast.pyx:152)
 * skip_out  : int32[B]    1 where a frame normaliser was 0 (ctc_fast.pyx:147-149)
 * workspace : ctcb_ctc_workspace_bytes(B, Tmax, max_labels) bytes of device scratch
 */
size_t ctcb_ctc_workspace_bytes(int B, int Tmax, int max_labels);
int ctcb_ctc_loss_grad_f32(const float *acts, int is_prob, int64_t utt_stride, int64_t frame_stride,
                           const int32_t *labels, const int32_t *label_off, const int32_t *T_per_utt,
                           int B, int Tmax, int K, int max_labels, int blank,
                           float *grad_out, float *nll_out, int32_t *skip_out,
                           void *workspace, size_t ws_bytes, void *stream);

/* ---- best-path decode (replaces ctc_fast.decode_best_path, ctc_fast.pyx:154-187) -----------
 * hyp_out/align_out: int32[B*Tmax] (row u holds hyp_len_out[u] valid entries).
 * drop_swbd_noise!=0 reproduces the reference's removal of labels 1, 2 and 8 (ctc_fast.pyx:176-179). */
int ctcb_ctc_best_path_f32(const float *acts, int64_t utt_stride, int64_t frame_stride,
                           const int32_t *T_per_utt, int B, int Tmax, int K, int blank,
                           int drop_swbd_noise, int32_t *hyp_out, int32_t *align_out,
                           int32_t *hyp_len_out, void *stream);

/* ---- blank-forced CTC (replaces ctc_fast_blankforce.ctc_loss, ctc-loss/ctc_fast_blankforce.pyx:13-113) ----
 * Same buffers as ctcb_ctc_loss_grad_f32, but `seq` already contains the blanks (:24-26): utterance u has
 * seq_off[u+1]-seq_off[u] trellis states (at most 1024), moves s->s and s->s+1 only, one start and one end
 * state, no window; state 0 is propagated with row 0 of the probabilities (:49).  grad_out must not alias
 * acts.  workspace: ctcb_ctc_blankforce_workspace_bytes(B, Tmax, max_states) bytes (the alpha trellis).
 * The best path of that module (:115-142) is ctcb_ctc_best_path_f32 with drop_swbd_noise=0; its align
 * output is simply not used. */
size_t ctcb_ctc_blankforce_workspace_bytes(int B, int Tmax, int max_states);
int ctcb_ctc_blankforce_loss_grad_f32(const float *acts, int is_prob, int64_t utt_stride, int64_t frame_stride,
                                      const int32_t *seq, const int32_t *seq_off, const int32_t *T_per_utt,
                                      int B, int Tmax, int K, int max_states,
                                      float *grad_out, float *nll_out, int32_t *skip_out,
                                      void *workspace, size_t ws_bytes, void *stream);

/* ---- dense fp32 contraction (replaces cudamat cm.dot call sites brnnet.py:140,196,204,227-230) ----
 * Row-major C[M x N] = alpha * op(A) * op(B) + beta * C, op(X) = X or X^T per transA/transB.
 * A is M x K (or K x M if transA), leading dimensions in elements.
 * bias (may be NULL): added per output column n.  relu!=0: C = max(C, 0) after bias.
 * mask_src (may be NULL, ldc-strided like C): C *= (mask_src > 0).
 * workspace is used for split-K partial sums (ctcb_gemm_workspace_bytes). */
size_t ctcb_gemm_workspace_bytes(int M, int N, int K);
int ctcb_gemm_f32(int transA, int transB, int M, int N, int K, float alpha,
                  const float *A, int64_t lda, const float *B, int64_t ldb, float beta,
                  float *C, int64_t ldc, const float *bias, int relu, const float *mask_src,
                  void *workspace, size_t ws_bytes, void *stream);

/* Diagnostic / measurement: force the CTC kernel organisation used by ctcb_ctc_loss_grad_f32 (and the BRNN step).
 * 0 = automatic (by batch size), 1 = one warp per utterance with the trellis spilled to the workspace, 2 = two warps per
 * utterance meeting in the middle, 3 = recurrences on two warps + frame-parallel gradient (at most 256 utterances),
 * 4 = one warp per utterance with on-chip checkpoints (at most 127 labels).  Same as the environment variable
 * CTCB_CTC=warp|pair|par|ckpt, which is read on first use.  Results are identical whichever is chosen. */
int ctcb_debug_set_ctc_kernel(int shape);

/* Diagnostic (CTCB_GEMM_TRACE=1): SM clock stamps of CTA (0,0,0) of the last tensor-core GEMM, HOST buffer of 64 x 8
 * uint64 = per k-block {TMA issued, tile landed, low halves written, MMA thread saw it, MMAs issued, -, -, -}. */
int ctcb_debug_gemm_trace(unsigned long long *host_out);

/* ---- BRNN (replaces nnets.brnnet.NNet, ctc_fast/nnets/brnnet.py:10-277) --------------------- */
typedef struct ctcb_brnn_config {
    int32_t inputDim;      /* brnnet.py:10 */
    int32_t outputDim;
    int32_t layerSize;
    int32_t numLayers;
    int32_t temporalLayer; /* <=0: none.  1..numLayers (numLayers itself is an extension) */
    int32_t maxT;          /* frames per utterance the buffers are sized for (reference: maxBatch) */
    int32_t maxB;          /* utterances per step */
    int32_t maxLabels;     /* longest label sequence */
    float reg;             /* L2 coefficient (brnnet.py:11,177-183,197-198,244-247) */
    float maxAct;          /* clip of the temporal layer, 20.0 (brnnet.py:32) */
    int32_t unidirectional; /* 0: nnets/brnnet.py (Wtf and Wtb, output For+Back).  1: nnets/rnnet.py:91-191 -- one
                             * recurrent matrix, forward in time only, stack = [...layers, [Wt, dummy]] (rnnet.py:57-65).
                             * temporalLayer <= 0 with either value is the plain DNN of nnets/nnet.py:57-113. */
} ctcb_brnn_config;

typedef struct ctcb_brnn ctcb_brnn; /* opaque; owns no device memory */

/* Flat parameter vector layout: the reference's `stack` order (brnnet.py:38-41,66-72):
 * [W1,b1] ... [W_{N+1},b_{N+1}] [Wtf,dummy] [Wtb,dummy]; W row-major (out x in), dummy = 1 float.
 * unidirectional: [W1,b1] ... [W_{N+1},b_{N+1}] [Wt,dummy]. */
int64_t ctcb_brnn_param_count(const ctcb_brnn_config *cfg);
int ctcb_brnn_num_tensors(const ctcb_brnn_config *cfg);                    /* 2 * len(stack) */
int ctcb_brnn_tensor_info(const ctcb_brnn_config *cfg, int idx, int64_t *offset, int32_t *rows, int32_t *cols);
size_t ctcb_brnn_workspace_bytes(const ctcb_brnn_config *cfg);
/* Byte offset, inside the workspace, of a uint32 that the recurrent-sweep kernels set non-zero when an
 * inter-CTA wait timed out (results are then invalid; never a hang).  Callers that synchronise anyway
 * (e.g. for the per-step log line) should read it. */
size_t ctcb_brnn_error_flag_offset(const ctcb_brnn_config *cfg);

/* Diagnostic hook (used by the parity tests): byte offset inside the workspace of the activations the LAST
 * ctcb_brnn_cost_and_grad call left behind, time-major [Tmax][B][*width] with that call's B and Tmax.
 * what = 0: output of affine map `layer` (1..numLayers+1; For+Back at the temporal layer, logits at the last),
 * what = 1 / 2: For / Back of the temporal layer (brnnet.py:143-153). */
int ctcb_brnn_activation_offset(const ctcb_brnn_config *cfg, int what, int layer, size_t *offset, int32_t *width);

int ctcb_brnn_create(const ctcb_brnn_config *cfg, ctcb_brnn **out);
void ctcb_brnn_destroy(ctcb_brnn *h);

/* costAndGrad over a batch (brnnet.py:117-249 per utterance; gradients SUMMED over utterances).
 * feats     : time-major [Tmax][B][inputDim] fp32 (frame (t,u) contiguous)
 * params    : flat parameter vector; grads: same layout, overwritten
 * cost_out  : float32[B] per-utterance nll (L2 term is returned separately in *regcost_out, device float)
 * probs_out : optional [Tmax][B][outputDim] softmax output (forward-only mode when grads==NULL,
 *             brnnet.py:171-173)
 * stats_out : optional device float[4] = {number of non-skipped utterances, sum of their nll,
 *             number skipped, sweep error flag (see ctcb_brnn_error_flag_offset)}.  Callers place it directly behind the flat gradient so that the
 *             data-parallel all-reduce of the gradient carries it along.
 */
int ctcb_brnn_cost_and_grad(ctcb_brnn *h, const float *feats, const int32_t *T_per_utt,
                            const int32_t *labels, const int32_t *label_off, int B, int Tmax,
                            const float *params, float *grads, float *cost_out, int32_t *skip_out,
                            float *regcost_out, float *probs_out, float *stats_out,
                            void *workspace, size_t ws_bytes, void *stream);

/* L2 regularisation under data parallelism.  By default ctcb_brnn_cost_and_grad adds reg*W to the weight
 * gradients itself (brnnet.py:197-198,244-247).  A data-parallel caller SUMS the gradient over ranks and must
 * add reg*W exactly once: it switches the handle to deferred mode (the call then only reports the L2 cost) and
 * calls ctcb_brnn_apply_l2_f32 on the reduced gradient. */
int ctcb_brnn_set_deferred_l2(ctcb_brnn *h, int deferred);
int ctcb_brnn_apply_l2_f32(ctcb_brnn *h, const float *params, float *grads, void *stream);

/* ---- the time recurrences of the temporal layer on their own (replaces the per-frame loops
 *      brnnet.py:144-152 [mode 0] and brnnet.py:208-224 [mode 1]) -------------------------------
 * All arrays time-major [T][B][H] fp32.  mode 0: outF[t] = clip(pre[t] + outF[t-1].Wf^T, 0, maxAct),
 * outB mirrored in time with Wb; utterance u is zero beyond T_per_utt[u].
 * mode 1: outF[t] = within(actF[t]) * (pre[t] + outF[t+1].Wf), outB mirrored.
 * Wb == NULL (with outB/actB ignored) runs the forward-in-time direction alone: the uni-directional layer of
 * nnets/rnnet.py:112-116 [mode 0] and :162-177 [mode 1].
 * scratch: 256-byte aligned device memory, >= 4096 bytes (word 0 = error flag, cleared by the call); the tensor-core
 * kernel (H >= 1024, sweep_tc.cu) additionally needs room for the (hi, lo) stacks of the recurrent matrices and the
 * ring of state low halves: ctcb_brnn_sweep_workspace_bytes(H, B) in total, else the FFMA kernels run. */
size_t ctcb_brnn_sweep_workspace_bytes(int H, int B);
int ctcb_brnn_sweep_f32(int mode, int T, int B, int H, const int32_t *T_per_utt, const float *pre,
                        const float *Wf, const float *Wb, float *outF, float *outB, const float *actF,
                        const float *actB, float maxAct, void *scratch, size_t scratch_bytes, void *stream);
/* 1 when the recurrences of a (layerSize H, B utterances, bi-directional) step run on the tensor-core kernel. */
int ctcb_sweep_uses_tensor_cores(int H, int B);

/* ---- optimiser (replaces sgd.SGD.run arithmetic, ctc_fast/sgd.py:91-161, and
 *      NNet.updateParams, brnnet.py:251-256) --------------------------------------------------- */
/* w += scale * u over n floats (updateParams) */
int ctcb_axpy_f32(float *w, const float *u, float scale, int64_t n, void *stream);
/* gnorm2_out[0] = sum(g^2) (sgd.py:103-107), deterministic two-stage reduction; scratch >= 4 KB (512 doubles) */
int ctcb_sumsq_f32(const float *g, int64_t n, float *gnorm2_out, void *scratch, void *stream);
/* Fused sgd.py:97-100,130-140,161 given w at the look-ahead point w + mom*v:
 *   w -= mom*v;  alph = alpha*min(1, maxGNorm/sqrt(*gnorm2));  v = mom*v - alph*g;  w += v
 * gnorm2 is read on the device (no host sync).  n_valid (optional device float): when it reads 0
 * every utterance of the step was skipped and only the undo `w -= mom*v` is applied, as the
 * reference's `if skip: continue` does (sgd.py:109-111). */
int ctcb_sgd_nesterov_step_f32(float *w, float *v, const float *g, int64_t n, float mom, float alpha,
                               float max_gnorm, const float *gnorm2, const float *n_valid, void *stream);

/* ---- CUDA graphs: replay one optimisation step (the reference issues ~16(N+3)+8(T-1) launches per utterance,
 *      sgd.py:91-161 + brnnet.py:117-249; this library ~40 per minibatch) as ONE graph launch -----------------------
 * Between begin and end, calls into this library on `stream` are captured instead of executed (including the internal
 * side stream and the NCCL exchange); every pointer, size and scalar argument is frozen into the graph, so a graph is
 * valid for one (buffers, B, Tmax, momentum, step size) combination.  Warm the same calls up once before capturing
 * (first-use initialisation is not capturable). */
typedef struct ctcb_graph ctcb_graph;
int ctcb_graph_capture_begin(void *stream);
int ctcb_graph_capture_end(void *stream, ctcb_graph **out);
int ctcb_graph_launch(ctcb_graph *g, void *stream);
void ctcb_graph_destroy(ctcb_graph *g);

/* ---- data-parallel exchange (the reference's only device hook is CUDA_DEVICE,
 *      ctc_fast/runNNet.py:117-120, and it has no multi-GPU path) -------------------------------------
 * One process per GPU.  Rank 0 obtains an id (HOST buffer of CTCB_COMM_ID_BYTES) and hands it to the other ranks by
 * whatever channel the host program has (a file, MPI, torch.distributed ...); every rank then creates its
 * communicator with the device it computes on current.  NCCL is loaded at run time (libnccl.so.2). */
#define CTCB_COMM_ID_BYTES 128
typedef struct ctcb_comm ctcb_comm;
int ctcb_comm_get_unique_id(void *id_out /* HOST */);
int ctcb_comm_create(const void *id /* HOST */, int rank, int world, ctcb_comm **out);
void ctcb_comm_destroy(ctcb_comm *comm);
int ctcb_comm_rank(const ctcb_comm *comm);
int ctcb_comm_world(const ctcb_comm *comm);
/* grads[0..n) <- sum over ranks, in place, asynchronous on `stream` (one ncclAllReduce over NVLink). */
int ctcb_allreduce_grads(ctcb_comm *comm, float *grads, int64_t n, void *stream);
/* Attach a communicator to a net: ctcb_brnn_cost_and_grad then performs the exchange itself -- the gradients of the
 * layers at and above the temporal layer are summed over ranks on an internal side stream WHILE the BPTT sweep runs,
 * the remaining tensors and the statistics tail (stats_out, normally grads + param_count) in one grouped launch at
 * the end -- and adds the L2 term once, after the sum.  On return (stream order) `grads` holds the global gradient. */
int ctcb_brnn_set_comm(ctcb_brnn *h, ctcb_comm *comm);
/* For a rank whose shard of the step is empty: the same sequence of collectives on a zero-filled gradient. */
int ctcb_brnn_exchange_only(ctcb_brnn *h, const float *params, float *grads, float *stats_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CTCB200_H */
