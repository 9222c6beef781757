"""Device pipeline behind the reference-named host classes: batched GP setup,
posterior + acquisition sweeps, knowledge gradient, reification and k-medoids,
all through the C ABI in include/barefoot_sm100.h.

One call here covers what the reference fans out as H (or m*N*H) Python work
items (barefoot.py:1070-1073, 1271-1283, 2001-2013).
"""
import os

import numpy as np
import torch

from . import _lib
from ._lib import (ACQ_EI, ACQ_GREEDY, ACQ_PI, ACQ_UCB, KERN, NSLOT_IDX, NSLOT_VAL, SLOT_EI,
                   SLOT_GREEDY, SLOT_PI, SLOT_TOP2, SLOT_UCB, call, ptr, stream)

F64 = torch.float64
SN_KG = 0.1      # util.py:260,619
XI = 0.01

ACQ_BITS = {"EI": ACQ_EI, "PI": ACQ_PI, "UCB": ACQ_UCB, "Greedy": ACQ_GREEDY, "KG": ACQ_GREEDY}
ACQ_SLOT = {"EI": SLOT_EI, "PI": SLOT_PI, "UCB": SLOT_UCB, "Greedy": SLOT_GREEDY}


def device():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


_copy_stream = None


def to_dev_async(a, dtype=F64):
    """Host -> device copy of a (pinned) host tensor / array on a side stream, so it overlaps the GP
    set-up kernels of the same work item.  Returns (device tensor, event to wait on); tensors that
    already live on the device come back with event None."""
    global _copy_stream
    if isinstance(a, torch.Tensor) and a.is_cuda:
        return a.to(dtype=dtype).contiguous(), None
    dev = device()
    t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a))
    t = t.to(dtype=dtype).contiguous()
    if _copy_stream is None:
        _copy_stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(_copy_stream):
        out = t.to(dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(_copy_stream)
    return out, ev


_read_stream = None


class AsyncRead:
    """Device -> host read of small result tensors that does NOT queue behind later work on the compute stream:
    the copies run on a read-back stream after an event recorded now, into pinned buffers; result() waits for
    them only.  `prepare` (optional) is called on the read-back stream first and returns the tensors to read
    (e.g. the wait + final kernel of an asynchronous cross-rank reduction).  A caller that pipelines work items
    (launch item s + 1, then read item s) keeps the GPU busy: with a plain .cpu() the read of item s would sit
    behind the sweep of item s + 1 in stream order and the queue would drain every step."""

    def __init__(self, tensors=(), prepare=None):
        global _read_stream
        dev = device()
        if _read_stream is None:
            _read_stream = torch.cuda.Stream(device=dev)
        ev = torch.cuda.Event()
        ev.record()
        self._keep = list(tensors)
        with torch.cuda.stream(_read_stream):
            _read_stream.wait_event(ev)
            src = list(prepare()) if prepare is not None else []
            src += list(tensors)
            self._keep += src
            self.host = []
            for t in src:
                h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                h.copy_(t, non_blocking=True)
                self.host.append(h)
            self.done = torch.cuda.Event()
            self.done.record(_read_stream)

    def result(self):
        self.done.synchronize()
        self._keep = None
        return self.host


_PIN_BELOW = 1 << 20        # bytes


def to_dev(a, dtype=F64):
    """Host array / tensor -> device tensor on the current stream.  Small host arrays (the per-call
    hyper-parameter rows, training data, scalars) go through pinned staging with an asynchronous copy: a copy
    from pageable memory makes the host wait until the stream has drained, which would serialise a caller that
    pipelines work items (the next item's set-up could not even be queued before the previous sweep ends)."""
    dev = device()                               # raises without a CUDA device: there is no CPU path
    if isinstance(a, torch.Tensor):
        if a.is_cuda or a.numel() * a.element_size() >= _PIN_BELOW or a.is_pinned():
            return a.to(device=dev, dtype=dtype, non_blocking=a.is_pinned() if not a.is_cuda else False).contiguous()
        t = a.to(dtype=dtype).contiguous()
    else:
        t = torch.as_tensor(np.ascontiguousarray(a), dtype=dtype)
        if t.numel() * t.element_size() >= _PIN_BELOW:
            return t.to(dev)
    if t.numel() == 0:
        return t.to(dev)
    return t.pin_memory().to(dev, non_blocking=True)


def amplitude(sigma_f, ndim):
    """george: ``c * kernel`` -> ConstantKernel(log(c / ndim)) (oracle/george_shim.py)."""
    return np.asarray(sigma_f, dtype=np.float64) / ndim


def inv_metric(l_sq):
    """george keeps log(metric) and evaluates exp(-log M)."""
    return np.exp(-np.log(np.asarray(l_sq, dtype=np.float64)))


def ucb_kt(ndim, iteration):
    """util.py:515-517,607-608."""
    beta = np.abs((2 * np.log(ndim * (iteration ** 2) * (np.pi ** 2) / (6 / 0.1))) / 5)
    return float(np.sqrt(0.2 * beta))


def linv_packed_doubles(n):
    """Doubles of the packed tensor-core operand of one n x n factor (bf_linv_packed_doubles)."""
    return int(_lib.load().bf_linv_packed_doubles(int(n)))


# Measurement hook (bench.py, tools/): set `timeline` to a list and every section below appends
# (name, start event, end event) recorded on the current stream; None = no events, no overhead.
timeline = None


class _Section:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if timeline is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if timeline is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            timeline.append((self.name, self.e0, e1))
        return False


def section_ms(name):
    """Sum of the recorded sections called `name` (synchronise first)."""
    return sum(a.elapsed_time(b) for nm, a, b in (timeline or []) if nm == name)


# Factorisation kernel choice.  "rl" (default) = bf_kernel_matrix_lower + the right-looking shared-memory-panel
# kernel bf_cholesky_batched; "ll" = the fused left-looking kernel bf_cholesky_fused (kernel tiles generated into
# the accumulators, no K in HBM, two sets per SM).  Measured on B200 at n = 500, S = 1000 (tools/bench_setup.py,
# tools/prof_chol_ll.cu, profiles/r2_cholesky_variants.json): rl 0.93 + 3.78 ms, ll 8.4 ms fused / 4.45 ms
# in place - the left-looking form is bound by the single-warp chains of its diagonal warp (L2-latency-bound
# accumulation through a two-slot ring + the 32 x 32 factorisation) and by FP64 kernel evaluation at 8 warps per
# SM, so it stays an option (it also covers padded sizes up to 1024, the right-looking panel 736).
CHOLESKY = "rl"
POISON_SCRATCH = False      # tests: fill the factor scratch with NaN first (no consumer may read an unwritten block)


def _scratch(shape, dev):
    if POISON_SCRATCH:
        return torch.full(shape, float("nan"), dtype=F64, device=dev)
    return torch.empty(shape, dtype=F64, device=dev)


def _use_fused(n, d, S):
    return CHOLESKY == "ll" and bool(_lib.load().bf_cholesky_fused_supported(int(n), int(d)))


_sm_count = None


def _whole_waves(chunk, S):
    """Sets per chunk of a chunked set-up: the batched Cholesky, inverse and solve kernels run one CTA (or a
    fixed number of CTAs) per set with one CTA per SM, so a memory-limited chunk of, say, 963 sets costs
    SEVEN waves of 148 - rounded down to a multiple of the SM count every chunk but the last fills its waves
    (888 sets: six)."""
    global _sm_count
    if chunk >= S:
        return chunk
    if _sm_count is None:
        _sm_count = torch.cuda.get_device_properties(device()).multi_processor_count
    return chunk // _sm_count * _sm_count if chunk >= _sm_count else chunk


class GPFactors:
    """Factorised GPs sharing the training inputs X: one "set" per hyper-parameter
    set (and, in the ROM stage, per fantasy point)."""

    def __init__(self, kern, X, inv_l2, amp, set_hp, linv_packed, alpha, S, L=None, Linv=None):
        self.kern, self.X, self.inv_l2, self.amp, self.set_hp = kern, X, inv_l2, amp, set_hp
        self.linv_packed, self.alpha, self.S = linv_packed, alpha, S
        self.n, self.d = X.shape
        self.L, self.Linv = L, Linv
        self.info = None

    def raise_if_not_positive_definite(self):
        """george / SciPy raise LinAlgError from the Cholesky factorisation (gpModel.py:41); here the per-set
        status of bf_cholesky_batched is read back (one device synchronisation).  build_factors(check=False)
        lets a caller run the sweep first and call this at the point where it reads its results anyway."""
        bad = int(self.info.max().item()) if (self.info is not None and self.S) else 0
        if bad:
            raise np.linalg.LinAlgError("%d-th leading minor of a GP kernel matrix is not positive definite" % bad)


def build_factors(kern, X, y, l_sq, sigma_f, yerr, ndim=None, set_hp=None, n_sets=None,
                  chunk_bytes=2 << 30, keep_dense=False, check=True):
    """gp_model.__init__ for a batch of sets (gpModel.py:17-42).

    X [n, d]; l_sq [H, d] (already squared where the reference squares, gpModel.py:20);
    sigma_f [H]; y [n] shared or [S, n]; yerr scalar, [n] shared or [S, n].
    set_hp [S] maps sets to rows of (l_sq, sigma_f); default identity with S = H."""
    dev = device()
    X = to_dev(X).reshape(len(X), -1)
    n, d = X.shape
    ndim = d if ndim is None else ndim
    l_sq = np.atleast_2d(np.asarray(l_sq, dtype=np.float64))
    H = l_sq.shape[0]
    inv_l2 = to_dev(inv_metric(l_sq))
    amp = to_dev(np.atleast_1d(amplitude(sigma_f, ndim)))
    if set_hp is None and n_sets is not None and n_sets != H:
        assert H == 1, "n_sets without set_hp needs a single hyper-parameter row"
        set_hp = np.zeros(n_sets, dtype=np.int32)
    if set_hp is None:
        S = H
        hp_t = None
    else:
        hp_t = to_dev(set_hp, torch.int32)
        S = hp_t.numel()
    y = to_dev(y)
    y_stride = n if y.dim() == 2 else 0
    yerr_t = to_dev(np.atleast_1d(yerr) if not isinstance(yerr, torch.Tensor) else yerr)
    if yerr_t.dim() == 2:
        yerr_n, yerr_stride = n, n
    else:
        yerr_n, yerr_stride = yerr_t.numel(), 0
    assert yerr_n in (1, n)
    lib = _lib.load()
    np_ = lib.bf_padded_n(n)
    pdoubles = lib.bf_linv_packed_doubles(n)
    linv_packed = torch.empty((S, pdoubles), dtype=F64, device=dev)
    alpha = torch.empty((S, n), dtype=F64, device=dev)
    info = torch.zeros((S,), dtype=torch.int32, device=dev)
    per_set = 2 * np_ * np_ * 8
    chunk = max(1, S) if keep_dense else _whole_waves(max(1, min(S, 65535, chunk_bytes // per_set)), S)
    Lbuf = _scratch((chunk, np_, np_), dev)
    Libuf = torch.empty((chunk, np_, np_), dtype=F64, device=dev)
    st = stream()
    with _Section("setup"):
        for s0 in range(0, S, chunk):
            sc = min(chunk, S - s0)
            hp_p = None if hp_t is None else hp_t[s0:].data_ptr()
            inv_p, amp_p = inv_l2, amp
            if hp_t is None:    # identity map: shift the HP rows instead
                inv_p, amp_p = inv_l2[s0:], amp[s0:]
            yerr_p = yerr_t[s0:].data_ptr() if yerr_stride else ptr(yerr_t)
            if _use_fused(n, d, sc):
                call("bf_cholesky_fused", KERN[kern], n, d, sc, ptr(X), inv_p.data_ptr(), amp_p.data_ptr(), hp_p, yerr_p,
                     yerr_n, yerr_stride, ptr(Lbuf), info[s0:].data_ptr(), None, int(bool(keep_dense)), st)
            else:
                call("bf_kernel_matrix_lower", KERN[kern], n, d, sc, ptr(X), inv_p.data_ptr(), amp_p.data_ptr(), hp_p,
                     yerr_p, yerr_n, yerr_stride, ptr(Lbuf), st)
                call("bf_cholesky_batched", n, sc, ptr(Lbuf), info[s0:].data_ptr(), st)
            call("bf_factor_finish", n, sc, ptr(Lbuf), y[s0:].data_ptr() if y_stride else ptr(y), y_stride,
                 ptr(Libuf), linv_packed[s0:].data_ptr(), alpha[s0:].data_ptr(), st)
    f = GPFactors(kern, X, inv_l2, amp, hp_t, linv_packed, alpha, S,
                  L=Lbuf if keep_dense else None, Linv=Libuf if keep_dense else None)
    f.info = info
    # the sweep's training-side operand: X pre-scaled by the metric of every hyper-parameter row
    f.Xw = torch.empty((H, n, d), dtype=F64, device=dev)
    call("bf_prescale_train", KERN[kern], n, d, H, ptr(X), ptr(inv_l2), ptr(f.Xw), st)
    # host copies of the small per-set scalars, so per-set launches (predict_cov) never read the device back
    f.amp_host = np.atleast_1d(amplitude(sigma_f, ndim)).astype(np.float64)
    f.set_hp_host = None if set_hp is None else np.asarray(set_hp, dtype=np.int64)
    if check:
        f.raise_if_not_positive_definite()
    return f


_setup_stream = None


def build_factors_overlapped(*args, **kwargs):
    """build_factors on a high-priority side stream: the set-up of work item s + 1 (a few small kernels and ONE
    Cholesky CTA that needs a whole SM) then runs WHILE the sweep of item s occupies the compute stream - it
    takes the next SM a sweep CTA frees instead of 0.6 ms of an otherwise idle GPU between two sweeps.  The
    compute stream waits for the factors before it uses them; every tensor of the result is marked as used by the
    compute stream so the allocator does not recycle it early.  Same arguments and result as build_factors
    (pass check=False and read the status later, or the call synchronises)."""
    global _setup_stream
    dev = device()
    main = torch.cuda.current_stream()
    if _setup_stream is None:
        _setup_stream = torch.cuda.Stream(device=dev, priority=-1)
    with torch.cuda.stream(_setup_stream):
        f = build_factors(*args, **kwargs)
        ready = torch.cuda.Event()
        ready.record(_setup_stream)
    main.wait_event(ready)
    for t in (f.X, f.Xw, f.inv_l2, f.amp, f.set_hp, f.linv_packed, f.alpha, f.info, f.L, f.Linv):
        if t is not None:
            t.record_stream(main)
    return f


class SolveSpec:
    """GPs that are swept over FEW candidates per set (the literal ROM stage, N ~ 50 << n): the inputs of
    build_factors, kept unfactorised.  posterior_sweep factorises them chunk by chunk and solves
    L v = [k* | y] directly (bf_gp_posterior_solve) - no L^-1, no alpha, nothing kept per set."""

    def __init__(self, kern, X, y, l_sq, sigma_f, yerr, ndim=None, set_hp=None, chunk_bytes=2 << 30):
        self.kern = kern
        self.X = to_dev(X).reshape(len(X), -1)
        self.n, self.d = self.X.shape
        self.ndim = self.d if ndim is None else ndim
        l_sq = np.atleast_2d(np.asarray(l_sq, dtype=np.float64))
        self.inv_l2 = to_dev(inv_metric(l_sq))
        self.amp = to_dev(np.atleast_1d(amplitude(sigma_f, self.ndim)))
        self.set_hp = None if set_hp is None else to_dev(set_hp, torch.int32)
        self.S = l_sq.shape[0] if set_hp is None else self.set_hp.numel()
        self.y = to_dev(y)
        self.y_stride = self.n if self.y.dim() == 2 else 0
        self.yerr = to_dev(np.atleast_1d(yerr) if not isinstance(yerr, torch.Tensor) else yerr)
        if self.yerr.dim() == 2:
            self.yerr_n, self.yerr_stride = self.n, self.n
        else:
            self.yerr_n, self.yerr_stride = self.yerr.numel(), 0
        assert self.yerr_n in (1, self.n)
        self.chunk_bytes = chunk_bytes

    @staticmethod
    def supported(n, n_candidates):
        """The solve path pays off while the candidates of a set are fewer than about half the
        training points (above that the explicit inverse is amortised)."""
        cols = _lib.load().bf_posterior_solve_cols(int(n))
        return cols > 0 and 0 < n_candidates <= max(cols, int(n) // 2)


def solve_sweep(f, Xs, scale=None, shift=None, want_mu=False, want_var=False, acq_mask=0,
                spread_is_sqrt=False, curr_max=0.0, xi=XI, kt=0.0, check=True):
    """posterior_sweep for a SolveSpec: kernel matrix -> Cholesky (+ diagonal-block inverses) ->
    forward substitution with the cross-covariances and y as right-hand sides, chunk of sets by chunk."""
    dev = device()
    Xs = to_dev(Xs)
    Xs = Xs.reshape(Xs.shape[0], f.d)
    N, S, n, d = Xs.shape[0], f.S, f.n, f.d
    lib = _lib.load()
    out = SweepResult()
    if want_mu:
        out.mu = torch.empty((S, N), dtype=F64, device=dev)
    if want_var:
        out.var = torch.empty((S, N), dtype=F64, device=dev)
    if N == 0:
        return out
    if acq_mask:
        out.best_val = torch.empty((S, NSLOT_VAL), dtype=F64, device=dev)
        out.best_idx = torch.empty((S, NSLOT_IDX), dtype=torch.int64, device=dev)
    np_ = lib.bf_padded_n(n)
    nb = np_ // 32
    parts = lib.bf_posterior_solve_parts(n, N)
    per_set = 8 * (np_ * np_ + nb * 1024)
    chunk = _whole_waves(max(1, min(S, 65535, f.chunk_bytes // per_set)), S)
    Lbuf = _scratch((chunk, np_, np_), dev)
    dinv = torch.empty((chunk, nb, 32, 32), dtype=F64, device=dev)
    info = torch.zeros((S,), dtype=torch.int32, device=dev)
    pv = pi = None
    if acq_mask:
        pv = torch.empty((chunk, parts, NSLOT_VAL), dtype=F64, device=dev)
        pi = torch.empty((chunk, parts, NSLOT_IDX), dtype=torch.int64, device=dev)
    st = stream()
    for s0 in range(0, S, chunk):
        sc = min(chunk, S - s0)
        hp_p = None if f.set_hp is None else f.set_hp[s0:].data_ptr()
        inv_p = f.inv_l2 if f.set_hp is not None else f.inv_l2[s0:]
        amp_p = f.amp if f.set_hp is not None else f.amp[s0:]
        yerr_p = f.yerr[s0:].data_ptr() if f.yerr_stride else ptr(f.yerr)
        if _use_fused(n, d, sc):
            with _Section("solve_cholesky_fused"):
                call("bf_cholesky_fused", KERN[f.kern], n, d, sc, ptr(f.X), inv_p.data_ptr(), amp_p.data_ptr(), hp_p,
                     yerr_p, f.yerr_n, f.yerr_stride, ptr(Lbuf), info[s0:].data_ptr(), ptr(dinv), 0, st)
        else:
            with _Section("solve_kernel_matrix"):
                call("bf_kernel_matrix_lower", KERN[f.kern], n, d, sc, ptr(f.X), inv_p.data_ptr(), amp_p.data_ptr(),
                     hp_p, yerr_p, f.yerr_n, f.yerr_stride, ptr(Lbuf), st)
            with _Section("solve_cholesky"):
                call("bf_cholesky_batched_dinv", n, sc, ptr(Lbuf), info[s0:].data_ptr(), ptr(dinv), st)
        e_ps = _Section("solve_posterior").__enter__()
        call("bf_gp_posterior_solve", KERN[f.kern], N, n, d, sc, ptr(Xs), ptr(f.X), inv_p.data_ptr(), amp_p.data_ptr(),
             hp_p, ptr(Lbuf), ptr(dinv), f.y[s0:].data_ptr() if f.y_stride else ptr(f.y), f.y_stride,
             None if scale is None else scale[s0:].data_ptr(), None if shift is None else shift[s0:].data_ptr(),
             None if out.mu is None else out.mu[s0:].data_ptr(), None if out.var is None else out.var[s0:].data_ptr(),
             acq_mask, int(bool(spread_is_sqrt)), float(curr_max), float(xi), float(kt), ptr(pv), ptr(pi), st)
        e_ps.__exit__()
        if acq_mask:
            call("bf_acq_finalize", sc, parts, acq_mask, ptr(pv), ptr(pi), out.best_val[s0:].data_ptr(),
                 out.best_idx[s0:].data_ptr(), st)
    bad = int(info.max().item()) if (S and check) else 0
    if bad:
        raise np.linalg.LinAlgError("%d-th leading minor of a GP kernel matrix is not positive definite" % bad)
    return out


class SweepResult:
    def __init__(self):
        self.mu = self.var = self.best_val = self.best_idx = None


# bench.py sets this to a list to receive one (start, end) CUDA-event pair per bf_gp_posterior launch
posterior_events = None
# True: the static split of the candidate tiles over the CTAs of a set (measurement aid; same results)
STATIC_TILE_SPLIT = bool(int(os.environ.get("BF_STATIC_TILE_SPLIT", "0")))


def posterior_sweep(f, Xs, scale=None, shift=None, want_mu=False, want_var=False, acq_mask=0,
                    spread_is_sqrt=False, curr_max=0.0, xi=XI, kt=0.0):
    """gp_model.predict_var (gpModel.py:50-54) for every set of `f` at the candidates
    Xs [N, d], with the de-normalisation of predict_fused_GP / predict_low_order
    (reificationFusion.py:189-206) and the elementwise acquisitions fused in."""
    if isinstance(f, SolveSpec):
        return solve_sweep(f, Xs, scale, shift, want_mu, want_var, acq_mask, spread_is_sqrt, curr_max, xi, kt)
    dev = device()
    Xs = to_dev(Xs)
    Xs = Xs.reshape(Xs.shape[0], f.d)
    N = Xs.shape[0]
    lib = _lib.load()
    out = SweepResult()
    S = f.S
    if N == 0:                  # empty candidate batch: nothing to launch (the reference's np.max would raise)
        out.mu = torch.empty((S, 0), dtype=F64, device=dev) if want_mu else None
        out.var = torch.empty((S, 0), dtype=F64, device=dev) if want_var else None
        return out
    if want_mu:
        out.mu = torch.empty((S, N), dtype=F64, device=dev)
    if want_var:
        out.var = torch.empty((S, N), dtype=F64, device=dev)
    if acq_mask:
        out.best_val = torch.empty((S, NSLOT_VAL), dtype=F64, device=dev)
        out.best_idx = torch.empty((S, NSLOT_IDX), dtype=torch.int64, device=dev)
    st = stream()
    for s0 in range(0, S, 65535):
        sc = min(65535, S - s0)
        sl = slice(s0, s0 + sc)
        pv = pi = None
        if acq_mask:
            parts = lib.bf_posterior_parts(f.n, N, sc)
            pv = torch.empty((sc, parts, NSLOT_VAL), dtype=F64, device=dev)
            pi = torch.empty((sc, parts, NSLOT_IDX), dtype=torch.int64, device=dev)
        hp_p = None if f.set_hp is None else f.set_hp[s0:].data_ptr()
        inv_p = f.inv_l2 if f.set_hp is not None else f.inv_l2[s0:]
        amp_p = f.amp if f.set_hp is not None else f.amp[s0:]
        xw_p = f.Xw if f.set_hp is not None else f.Xw[s0:]
        # per-set work counters: the candidate tiles of a set go to its CTAs dynamically (cleared by the call)
        ctr = None if STATIC_TILE_SPLIT else torch.empty((sc,), dtype=torch.int32, device=dev)
        if posterior_events is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        with _Section("posterior_kernel"):
            call("bf_gp_posterior", KERN[f.kern], N, f.n, f.d, sc, ptr(Xs), xw_p.data_ptr(), inv_p.data_ptr(),
                 amp_p.data_ptr(), hp_p, f.linv_packed[s0:].data_ptr(), f.alpha[s0:].data_ptr(),
                 None if scale is None else scale[s0:].data_ptr(),
                 None if shift is None else shift[s0:].data_ptr(),
                 None if out.mu is None else out.mu[s0:].data_ptr(),
                 None if out.var is None else out.var[s0:].data_ptr(),
                 acq_mask, int(bool(spread_is_sqrt)), float(curr_max), float(xi), float(kt),
                 ptr(pv), ptr(pi), ptr(ctr), st)
        if posterior_events is not None:
            ev1.record()
            posterior_events.append((ev0, ev1))
        if acq_mask:
            call("bf_acq_finalize", sc, parts, acq_mask, ptr(pv), ptr(pi),
                 out.best_val[sl].data_ptr(), out.best_idx[sl].data_ptr(), st)
    return out


def solve_rows(f, Xa, set_index=0, want_mean=True):
    """Vt = k(Xa, X) L^-T (and the posterior mean) for one set of `f` (needs keep_dense factors)."""
    if f.Linv is None:
        raise ValueError("solve_rows needs factors built with keep_dense=True")
    Xa = to_dev(Xa).reshape(len(Xa), -1)
    A = Xa.shape[0]
    lib = _lib.load()
    np_ = lib.bf_padded_n(f.n)
    Ap = (A + 31) // 32 * 32
    Ks = torch.empty((Ap, np_), dtype=F64, device=Xa.device)
    Vt = torch.empty((Ap, np_), dtype=F64, device=Xa.device)
    mean = torch.empty((A,), dtype=F64, device=Xa.device) if want_mean else None
    hp_host = getattr(f, "set_hp_host", None)
    if f.set_hp is None:
        hp = set_index
    else:
        hp = int(hp_host[set_index]) if hp_host is not None else int(f.set_hp[set_index].item())
    amp_host = getattr(f, "amp_host", None)
    amp = float(amp_host[hp]) if amp_host is not None else float(f.amp[hp].item())
    call("bf_gp_solve_rows", KERN[f.kern], A, f.n, f.d, ptr(Xa), ptr(f.X), f.inv_l2[hp].data_ptr(), amp,
         f.Linv[set_index].data_ptr(), f.alpha[set_index].data_ptr(), ptr(Ks), ptr(Vt), ptr(mean), stream())
    return Xa, Vt, mean, hp, amp


def posterior_cov(f, Xa, Xb=None, set_index=0, out=None):
    """gp_model.predict_cov (gpModel.py:44-48): posterior mean at Xa and the full covariance
    between Xa and Xb (default Xb = Xa) for one set (written into `out` [A, B] when given)."""
    Xa, Vta, mean, hp, amp = solve_rows(f, Xa, set_index)
    if Xb is None:
        Xb, Vtb = Xa, Vta
    else:
        Xb, Vtb, _, _, _ = solve_rows(f, Xb, set_index, want_mean=False)
    A, B = Xa.shape[0], Xb.shape[0]
    C = torch.empty((A, B), dtype=F64, device=Xa.device) if out is None else out
    call("bf_gp_cov", KERN[f.kern], A, B, f.n, f.d, ptr(Xa), ptr(Xb), f.inv_l2[hp].data_ptr(), amp,
         ptr(Vta), ptr(Vtb), ptr(C), B, stream())
    return mean, C


def kg_full(mu, sigma, M=None, sn=SN_KG):
    """knowledge_gradient(M, sn, mu, sigma) with a dense covariance (acquisitionFunc.py:12-96):
    mu [N] or [S, N], sigma [N, N] or [S, N, N]; ONE launch for all S sets.  N <= 4096 sorts in shared memory,
    larger N (to 65536) on a device scratch.  Returns device tensors (kg_val [S], kg_idx [S], nu [S, M])."""
    mu = to_dev(mu)
    mu2 = mu.reshape(1, -1) if mu.dim() == 1 else mu
    S, N = mu2.shape
    sg = to_dev(sigma).reshape(S, N, N)
    M = N if M is None else int(M)
    nu = torch.empty((S, max(M, 1)), dtype=F64, device=mu2.device)
    kg_val = torch.empty((S,), dtype=F64, device=mu2.device)
    kg_idx = torch.empty((S,), dtype=torch.int64, device=mu2.device)
    st = stream()
    lib = _lib.load()
    for s0 in range(0, S, 65535):
        sc = min(65535, S - s0)
        nbytes = lib.bf_kg_full_workspace(N, M, sc)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=mu2.device) if nbytes else None
        call("bf_kg_full", N, M, sc, mu2[s0:].data_ptr(), sg[s0:].data_ptr(), float(sn), nu[s0:].data_ptr(),
             kg_val[s0:].data_ptr(), kg_idx[s0:].data_ptr(), ptr(ws), st)
    return kg_val, kg_idx, nu[:, :M]


def kg_full_from_factors(f, x_test, shift=0.0, sn=SN_KG, max_bytes=8 << 30):
    """Batch-only KG (util.py:952,1078-1081): the full posterior covariance at x_test (predict_cov) of every set,
    then the dense knowledge gradient - the covariances of a chunk of sets (at most max_bytes) are built by
    per-set launches without any host read-back and scored by ONE bf_kg_full launch.  Returns host arrays
    (values [S], indices [S])."""
    xs = to_dev(x_test).reshape(len(x_test), -1)
    N = xs.shape[0]
    per_chunk = max(1, min(f.S, int(max_bytes // max(1, 8 * N * N))))
    vals, idxs = [], []
    for s0 in range(0, f.S, per_chunk):
        sc = min(per_chunk, f.S - s0)
        means = torch.empty((sc, N), dtype=F64, device=xs.device)
        covs = torch.empty((sc, N, N), dtype=F64, device=xs.device)
        for s in range(sc):
            mean, _ = posterior_cov(f, xs, set_index=s0 + s, out=covs[s])
            means[s] = mean if not shift else mean + float(shift)
        v, i, _ = kg_full(means, covs, sn=sn)
        vals.append(v)
        idxs.append(i)
    return torch.cat(vals).cpu().numpy(), torch.cat(idxs).cpu().numpy()


def kg_from_sweep(res, M=None, sn=SN_KG, want_nu=False):
    """knowledge_gradient on np.diag(var) (acquisitionFunc.py:12-96 via util.py:259-262):
    needs a sweep run with want_mu, want_var and ACQ_GREEDY."""
    dev = device()
    S, N = res.mu.shape
    M = N if M is None else M
    lib = _lib.load()
    tiles = lib.bf_kg_tiles(N)
    pv = torch.empty((S, tiles), dtype=F64, device=dev)
    pi = torch.empty((S, tiles), dtype=torch.int64, device=dev)
    kg_val = torch.empty((S,), dtype=F64, device=dev)
    kg_idx = torch.empty((S,), dtype=torch.int64, device=dev)
    nu = torch.full((S, N), float("-inf"), dtype=F64, device=dev) if want_nu else None
    st = stream()
    with _Section("kg"):
        for s0 in range(0, S, 65535):
            sc = min(65535, S - s0)
            call("bf_kg_diag", N, M, sc, res.mu[s0:].data_ptr(), res.var[s0:].data_ptr(),
                 res.best_val[s0:].data_ptr(), res.best_idx[s0:].data_ptr(), float(sn),
                 None if nu is None else nu[s0:].data_ptr(), pv[s0:].data_ptr(), pi[s0:].data_ptr(),
                 kg_val[s0:].data_ptr(), kg_idx[s0:].data_ptr(), st)
    return kg_val, kg_idx, nu


def acquisition_sweep(f, Xs, name, stage, curr_max, iteration, scale=None, shift=None, xi=XI, M=None):
    """One acquisition for every set: returns device tensors (values [S], indices [S]).
    stage "ROM": PI/UCB see sqrt(var) (util.py:451,519); "TM"/"BATCH": var (util.py:598-612)."""
    kt = ucb_kt(f.d, iteration) if name == "UCB" else 0.0
    kg = name == "KG"
    res = posterior_sweep(f, Xs, scale, shift, want_mu=kg, want_var=kg, acq_mask=ACQ_BITS[name],
                          spread_is_sqrt=(stage == "ROM"), curr_max=curr_max, xi=xi, kt=kt)
    if kg:
        v, i, _ = kg_from_sweep(res, M=M)
        return v, i, res
    slot = ACQ_SLOT[name]
    return res.best_val[:, slot], res.best_idx[:, slot], res


def acq_eval(name, y, spread, curr_max=0.0, xi=XI, kt=0.0, want_all=True):
    """Stand-alone EI / PI / UCB / Greedy on given vectors (acquisitionFunc.py:98-214):
    y, spread [N] or [S, N].  Returns device tensors (best_val [S], best_idx [S], values [S, N])."""
    y = to_dev(y)
    y2 = y.reshape(1, -1) if y.dim() == 1 else y
    S, N = y2.shape
    sp = None
    if spread is not None:
        sp = to_dev(spread).reshape(S, N)
    lib = _lib.load()
    tiles = lib.bf_acq_tiles(N)
    dev = y2.device
    pv = torch.empty((S, tiles), dtype=F64, device=dev)
    pi = torch.empty((S, tiles), dtype=torch.int64, device=dev)
    bv = torch.empty((S,), dtype=F64, device=dev)
    bi = torch.empty((S,), dtype=torch.int64, device=dev)
    out = torch.empty((S, N), dtype=F64, device=dev) if want_all else None
    st = stream()
    for s0 in range(0, S, 65535):
        sc = min(65535, S - s0)
        call("bf_acq_eval", ACQ_BITS[name], N, sc, y2[s0:].data_ptr(),
             None if sp is None else sp[s0:].data_ptr(), float(curr_max), float(xi), float(kt),
             None if out is None else out[s0:].data_ptr(), pv[s0:].data_ptr(), pi[s0:].data_ptr(),
             bv[s0:].data_ptr(), bi[s0:].data_ptr(), st)
    return bv, bi, out


def kg_diag(mu, var, M=None, sn=SN_KG, want_nu=True):
    """knowledge_gradient(M, sn, mu, np.diag(var)) on given vectors [N] or [S, N]: a greedy pass
    for the top-2 of the mean, then the closed form (bf_kg_diag)."""
    mu = to_dev(mu)
    mu2 = mu.reshape(1, -1) if mu.dim() == 1 else mu
    S, N = mu2.shape
    res = SweepResult()
    res.mu, res.var = mu2, to_dev(var).reshape(S, N)
    res.best_val = torch.full((S, NSLOT_VAL), float("-inf"), dtype=F64, device=mu2.device)
    res.best_idx = torch.zeros((S, NSLOT_IDX), dtype=torch.int64, device=mu2.device)
    call("bf_top2", N, S, ptr(mu2), ptr(res.best_val), ptr(res.best_idx), stream())
    return kg_from_sweep(res, M=M, sn=sn, want_nu=want_nu)


def reification(y, var):
    """reification(y, sig) (reificationFusion.py:26-57) on device tensors [m, P]."""
    y = to_dev(y)
    var = to_dev(var)
    m, P = y.shape
    mean_f = torch.empty((P,), dtype=F64, device=y.device)
    var_f = torch.empty((P,), dtype=F64, device=y.device)
    call("bf_reification", m, P, ptr(y), ptr(var), ptr(mean_f), ptr(var_f), stream())
    return mean_f, var_f


def affine(a, scale, shift):
    out = torch.empty_like(a)
    call("bf_affine", a.numel(), ptr(a), float(scale), float(shift), ptr(out), stream())
    return out


def total_variance(err_mu, err_std, err_mean, var, model_std):
    out = torch.empty_like(var)
    call("bf_total_variance", var.numel(), ptr(err_mu), float(err_std), float(err_mean), ptr(var),
         float(model_std), ptr(out), stream())
    return out


def zscore_targets(f_mean, f_var):
    """reificationFusion.py:154-159 -> (z, zerr, stats[mean, std]) on the device."""
    n = f_mean.numel()
    z = torch.empty_like(f_mean)
    zerr = torch.empty_like(f_mean)
    stats = torch.empty((2,), dtype=F64, device=f_mean.device)
    call("bf_zscore_targets", n, ptr(f_mean), ptr(f_var), ptr(z), ptr(zerr), ptr(stats), stream())
    return z, zerr, stats


def zscore_targets_batched(f_mean, f_var):
    """zscore_targets for G groups at once: f_mean, f_var [G, n] -> (z [G, n], zerr [G, n], stats [G, 2])."""
    G, n = f_mean.shape
    z = torch.empty_like(f_mean)
    zerr = torch.empty_like(f_mean)
    stats = torch.empty((G, 2), dtype=F64, device=f_mean.device)
    call("bf_zscore_targets_batched", G, n, ptr(f_mean), ptr(f_var), ptr(z), ptr(zerr), ptr(stats), stream())
    return z, zerr, stats


def fantasy_rows(mu_f, var_f, C, mu_t, var_t, y_new, noise_var, gp_mean, model_std, model_mean, err_term):
    """bf_fantasy_rows: the de-normalised mean / total-variance rows [K, F] of one ROM after appending each of K
    fantasy points in turn (bordered form of gp_model.update, gpModel.py:56-64)."""
    F, K = mu_f.numel(), mu_t.numel()
    mean_rows = torch.empty((K, F), dtype=F64, device=mu_f.device)
    var_rows = torch.empty((K, F), dtype=F64, device=mu_f.device)
    call("bf_fantasy_rows", F, K, ptr(mu_f), ptr(var_f), ptr(C), C.stride(0), ptr(mu_t), ptr(var_t), ptr(y_new),
         float(noise_var), float(gp_mean), float(model_std), float(model_mean), ptr(err_term), ptr(mean_rows),
         ptr(var_rows), stream())
    return mean_rows, var_rows


# ---- k-medoids (util.py:38-49) -----------------------------------------------------------------
def kmedoids_rowsum(X, row0=0, rows=None):
    Xd = to_dev(X)
    N, f = Xd.shape
    rows = N - row0 if rows is None else rows
    out = torch.empty((rows,), dtype=F64, device=Xd.device)
    call("bf_kmedoids_rowsum", N, f, ptr(Xd), row0, rows, ptr(out), stream())
    return out


def kmedoids_rowsum_sharded(X):
    """Row sums of the implicit N x N distance matrix with the row blocks split over the ranks of a
    torchrun job ("row blocks of the implicit D"); one all-gather of 8 N bytes.  Every
    row sum is computed by exactly one rank with the same kernel, so the result does not depend on
    the number of ranks."""
    from . import dist as bdist
    Xd = to_dev(X)
    lo, hi = bdist.shard_range(Xd.shape[0])
    return bdist.gather_rows(kmedoids_rowsum(Xd, lo, hi - lo), Xd.shape[0])


def kmedoids_fit(X, k, max_iter=300, rowsum=None):
    """sklearn_extra KMedoids(k, random_state=0) defaults (oracle/kmedoids_shim.py).  All O(N^2) and
    O(N k) work runs on the device: row sums, label assignment, the stable label sort
    (bf_kmedoids_sort_labels), in-cluster costs and the medoid update; the host sees the N row sums
    once (np.argpartition defines the order of the initial medoids) and one `changed` flag per
    iteration.  Under torchrun the row sums and the in-cluster cost chunks are split over the ranks:
    one all-gather of the row sums, then per iteration one all-gather of k (cost, position) records."""
    from . import dist as bdist
    Xd = to_dev(X)
    N, f = Xd.shape
    if k > N:
        raise ValueError("The number of medoids (%d) must be less than the number of samples %d." % (k, N))
    rank, world = bdist.world()
    if rowsum is None:
        rowsum = kmedoids_rowsum_sharded(Xd) if world > 1 else kmedoids_rowsum(Xd)
    med = np.argpartition(rowsum.cpu().numpy(), k - 1)[:k].astype(np.int64)
    med_d = to_dev(med, torch.int64)
    dev = Xd.device
    labels = torch.empty((N,), dtype=torch.int32, device=dev)
    cost = torch.empty((N,), dtype=F64, device=dev)
    perm = torch.empty((N,), dtype=torch.int64, device=dev)
    seg = torch.empty((k + 1,), dtype=torch.int64, device=dev)
    ws = torch.empty((_lib.load().bf_kmedoids_sort_workspace(N, k),), dtype=torch.uint8, device=dev)
    # one `changed` counter per iteration, read back every `poll` iterations: an iteration that moves no
    # medoid is a fixed point of the loop, so the iterations run past it change nothing and the reported
    # count is the first iteration whose counter is zero (KMedoids.n_iter_)
    changed = torch.zeros((max_iter,), dtype=torch.int32, device=dev)
    st = stream()
    # one host call per iteration (bf_kmedoids_iteration: assign, label sort, this rank's cost chunks, update);
    # under torchrun each rank reduces its chunks to k (cost, position, current-medoid cost) records, the ranks
    # all-gather 24 k bytes each and bf_kmedoids_update_final applies the update on every rank
    recs = gathered = None
    if world > 1:
        recs = torch.empty((k, 3), dtype=F64, device=dev)
        gathered = torch.empty((world, k, 3), dtype=F64, device=dev)
    poll = (8 if world > 1 else 4) if N >= 20000 else 1
    n_iter = max_iter
    for it in range(max_iter):
        call("bf_kmedoids_iteration", N, f, k, ptr(Xd), ptr(med_d), ptr(labels), ptr(perm), ptr(seg), ptr(ws),
             rank, world, ptr(cost), ptr(recs), changed[it:].data_ptr(), st)
        if world > 1:
            bdist.all_gather_flat(gathered, recs)
            call("bf_kmedoids_update_final", k, world, ptr(gathered), ptr(perm), ptr(med_d), changed[it:].data_ptr(), st)
        if (it + 1) % poll == 0 or it + 1 == max_iter:
            done = (changed[: it + 1] == 0).nonzero()
            if done.numel():
                n_iter = int(done[0].item()) + 1
                break
    call("bf_kmedoids_assign", N, f, k, ptr(Xd), ptr(med_d), ptr(labels), st)
    return med_d.cpu().numpy(), labels.cpu().numpy().astype(np.int64), n_iter


def group_best(keys, values, n_groups):
    """barefoot.__kg_calc_clustering's dictionary loop (barefoot.py:2102-2140) for all records at once:
    per key the reference's in-order "first strict maximum".  keys [R] int64 in [0, n_groups),
    values [R].  Returns device tensors (first_row [G], best_val [G], best_row [G]); empty groups have
    first_row == R and best_row == -1."""
    dev = device()
    k = keys if isinstance(keys, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(keys, dtype=np.int64))
    k = k.to(device=dev, dtype=torch.int64).contiguous()
    v = to_dev(values)
    R, G = k.numel(), int(n_groups)
    first = torch.empty((G,), dtype=torch.int64, device=dev)
    best_val = torch.empty((G,), dtype=F64, device=dev)
    best_row = torch.empty((G,), dtype=torch.int64, device=dev)
    bad = torch.zeros((1,), dtype=torch.int32, device=dev)
    call("bf_group_best", R, G, ptr(k) if R else None, ptr(v) if R else None, ptr(first), ptr(best_val),
         ptr(best_row), ptr(bad), stream())
    if int(bad.item()):
        raise ValueError("group_best: %d key(s) outside [0, %d)" % (int(bad.item()), G))
    return first, best_val, best_row
