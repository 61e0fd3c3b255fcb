"""ctc_fast -- drop-in for the reference's Cython module of the same name
(/root/reference/ctc_fast/ctc-loss/ctc_fast.pyx), computed on the B200 by libctcb200.

    ctc_loss(params, seq, blank=0)        -> (nll, grad, skip)      ctc_fast.pyx:13-152
    decode_best_path(probs, blank=0)      -> (hyp, align)           ctc_fast.pyx:154-187

Both keep the reference's argument contract (params: K x T float64 Fortran-contiguous ndarray of
per-frame probability distributions, seq: int32 labels) and error behaviour (ValueError for a
non-Fortran array; numerical infeasibility is reported through skip=True, never raised).
The batched, zero-copy entry `ctc_loss_batch` is what nnets.brnnet uses internally through the
C ABI; it is exported for callers that already hold activations on the device.
"""
import numpy as np

import _ctcb
from _ctcb import lib, check, ptr


def _check_params(params):
    if params is None:
        raise TypeError("Argument 'params' must not be None")
    if not isinstance(params, np.ndarray) or params.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2)")
    if params.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'double' but got '%s'" % params.dtype)
    if not params.flags.f_contiguous:
        raise ValueError("ndarray is not Fortran contiguous")


def ctc_loss_batch(acts, T_per_utt, labels, label_off, max_labels, blank=0, is_prob=False,
                   utt_stride=None, frame_stride=None, grad=None, workspace=None):
    """Batched CTC on device tensors.  acts: float32 CUDA tensor holding B utterances of up to Tmax
    frames x K classes; default layout [B][Tmax][K] (pass strides, in elements, for time-major data).
    T_per_utt/labels/label_off: int32 CUDA tensors.  Returns (nll[B], grad like acts, skip[B]) on the
    device without synchronising."""
    torch = _ctcb.require_cuda()
    assert acts.is_cuda and acts.dtype == torch.float32 and acts.is_contiguous()
    B = T_per_utt.numel()
    K = acts.shape[-1]
    if utt_stride is None:
        Tmax = acts.shape[1]
        utt_stride, frame_stride = Tmax * K, K
    else:
        Tmax = acts.numel() // (B * K)
    if grad is None:
        grad = torch.empty_like(acts)
    nbytes = lib.ctcb_ctc_workspace_bytes(B, Tmax, int(max_labels))
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=acts.device)
    nll = torch.empty(B, dtype=torch.float32, device=acts.device)
    skip = torch.empty(B, dtype=torch.int32, device=acts.device)
    check(lib.ctcb_ctc_loss_grad_f32(ptr(acts), int(bool(is_prob)), int(utt_stride), int(frame_stride), ptr(labels),
                                     ptr(label_off), ptr(T_per_utt), B, Tmax, K, int(max_labels), int(blank),
                                     ptr(grad), ptr(nll), ptr(skip), ptr(workspace), workspace.numel(),
                                     _ctcb.current_stream()))
    return nll, grad, skip


def ctc_loss(params, seq, blank=0):
    """CTC loss function (reference signature).  params - n x m matrix of n-D probability
    distributions over m frames, Fortran order; seq - label ids.  Returns (objective, gradient with
    respect to the unnormalised inputs, skip)."""
    _check_params(params)
    if seq is None:
        raise TypeError("Argument 'seq' must not be None")
    seq = np.ascontiguousarray(seq)
    if seq.dtype != np.int32 or seq.ndim != 1:
        raise ValueError("Buffer dtype mismatch, expected 'int' but got '%s'" % seq.dtype)
    torch = _ctcb.require_cuda()
    K, T = params.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    # K x T Fortran == T x K row-major: the frame-contiguous layout the kernel streams
    acts = torch.from_numpy(np.ascontiguousarray(params.T, dtype=np.float32)).to(dev).view(1, T, K)
    lab = torch.from_numpy(seq if seq.size else np.zeros(1, np.int32)).to(dev)
    off = torch.tensor([0, seq.shape[0]], dtype=torch.int32, device=dev)
    tl = torch.tensor([T], dtype=torch.int32, device=dev)
    nll, grad, skip = ctc_loss_batch(acts, tl, lab, off, seq.shape[0], blank=blank, is_prob=True)
    g = grad.view(T, K).cpu().numpy().astype(np.float64)
    return float(nll.item()), np.asfortranarray(g.T), bool(skip.item())


def decode_best_path(probs, blank=0):
    """Best path: most likely label per frame, collapse repeats, drop blanks and (as the reference
    does, ctc_fast.pyx:176-179) the SWBD noise labels 1, 2 and 8.  Returns (hyp, align)."""
    _check_params(probs)
    torch = _ctcb.require_cuda()
    K, T = probs.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    acts = torch.from_numpy(np.ascontiguousarray(probs.T, dtype=np.float32)).to(dev)
    tl = torch.tensor([T], dtype=torch.int32, device=dev)
    hyp = torch.empty(T, dtype=torch.int32, device=dev)
    ali = torch.empty(T, dtype=torch.int32, device=dev)
    n = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib.ctcb_ctc_best_path_f32(ptr(acts), T * K, K, ptr(tl), 1, T, K, int(blank), 1, ptr(hyp), ptr(ali),
                                     ptr(n), _ctcb.current_stream()))
    n = int(n.item())
    return hyp[:n].cpu().tolist(), ali[:n].cpu().tolist()
