"""ctypes binding of libai4xde_b200.so (C ABI declared in include/ai4xde_b200.h).

There is no CPU fallback: every compute entry point raises if the shared library is missing
or if it is handed anything but CUDA tensors.  torch is used here only for device memory and
the current stream.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libai4xde_b200.so")

MAX_LINEAR, MAX_FIRST, MAX_SECOND = 16, 4, 10
ACT = {"tanh": 0, "sin": 1, "sigmoid": 2}
ENGINE = {"auto": 0, "fp32": 1, "tf32": 2}


class NetDesc(C.Structure):
    _fields_ = [
        ("n_linear", C.c_int32),
        ("widths", C.c_int32 * (MAX_LINEAR + 1)),
        ("activation", C.c_int32),
        ("n_nets", C.c_int32),
        ("first_activation", C.c_int32),
    ]


class JetDesc(C.Structure):
    _fields_ = [
        ("n1", C.c_int32),
        ("dir1", C.c_int32 * MAX_FIRST),
        ("n2", C.c_int32),
        ("pa", C.c_int32 * MAX_SECOND),
        ("pb", C.c_int32 * MAX_SECOND),
    ]


_P = C.c_void_p


class ResidualProgram(C.Structure):
    _fields_ = [
        ("code", _P), ("segments", _P), ("consts", _P), ("aux", _P), ("aux_meta", _P), ("scalars", _P),
        ("term_scale", _P),
        ("n_instr", C.c_int32), ("n_seg", C.c_int32), ("n_reg", C.c_int32), ("n_terms", C.c_int32),
        ("n_streams", C.c_int32), ("n_const", C.c_int32), ("n_scalars", C.c_int32),
    ]


_SIGNATURES = {
    "ax_version": (C.c_int, []),
    "ax_last_error": (C.c_char_p, []),
    "ax_launch_count": (C.c_int64, []),
    "ax_timing_enable": (None, [C.c_int]),
    "ax_timing_report": (C.c_int64, [C.c_char_p, C.c_int64]),
    "ax_param_count": (C.c_int64, [C.POINTER(NetDesc)]),
    "ax_jet_supported": (C.c_int, [C.POINTER(JetDesc)]),
    "ax_engine_supported": (C.c_int, [C.POINTER(NetDesc), C.POINTER(JetDesc), C.c_int]),
    "ax_narrow_supported": (C.c_int, [C.POINTER(NetDesc), C.POINTER(JetDesc)]),
    "ax_stash_bytes": (C.c_size_t, [C.POINTER(NetDesc), C.POINTER(JetDesc), C.c_int64, C.c_int]),
    "ax_stash_bytes_engine": (C.c_size_t, [C.POINTER(NetDesc), C.POINTER(JetDesc), C.c_int64, C.c_int, C.c_int]),
    "ax_forward_ws_bytes": (C.c_size_t, [C.POINTER(NetDesc), C.POINTER(JetDesc), C.c_int64]),
    "ax_backward_ws_bytes": (C.c_size_t, [C.POINTER(NetDesc), C.POINTER(JetDesc), C.c_int64]),
    "ax_jet_mlp_forward": (
        C.c_int,
        [C.POINTER(NetDesc), C.POINTER(JetDesc), _P, _P, C.c_int64, C.c_int64, _P, _P, C.c_int, _P,
         C.c_size_t, C.c_int, _P],
    ),
    "ax_jet_mlp_backward": (
        C.c_int,
        [C.POINTER(NetDesc), C.POINTER(JetDesc), _P, _P, C.c_int64, C.c_int64, _P, _P, _P, _P,
         C.c_size_t, C.c_int, _P],
    ),
    "ax_fbpinn_window_supported": (C.c_int, [C.POINTER(NetDesc), C.POINTER(JetDesc), C.c_int]),
    "ax_fbpinn_window_ws_bytes": (C.c_size_t, [C.POINTER(NetDesc), C.POINTER(JetDesc), C.c_int64]),
    "ax_fbpinn_window_forward": (
        C.c_int,
        [C.POINTER(NetDesc), C.POINTER(JetDesc), _P, _P, C.c_int64, C.c_int64, _P, _P, _P, C.c_int32, _P, _P,
         C.c_size_t, C.c_int, _P],
    ),
    "ax_resample_set_fast": (None, [C.c_int]),
    "ax_resample_get_fast": (C.c_int, []),
    "ax_resample_ws_bytes": (C.c_size_t, [C.c_int64]),
    "ax_pairwise_sum_f32": (C.c_int, [_P, C.c_int64, _P, _P, C.c_size_t, _P]),
    "ax_pairwise_sum_f64": (C.c_int, [_P, C.c_int64, _P, _P, C.c_size_t, _P]),
    "ax_rad_probabilities": (
        C.c_int,
        [_P, C.c_int32, C.c_int64, C.c_int64, C.c_double, C.c_double, _P, _P, C.c_size_t, _P],
    ),
    "ax_choice_round": (
        C.c_int,
        [_P, C.c_int64, _P, C.c_int64, _P, C.c_int64, _P, _P, C.c_size_t, _P],
    ),
    "ax_r3_retain": (
        C.c_int,
        [_P, C.c_int32, C.c_int64, _P, C.c_int32, C.c_int64, C.c_int64, _P, C.c_int64, C.c_int32, _P,
         _P, _P, _P, C.c_size_t, _P],
    ),
    "ax_r3_abs_sum": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int64, _P, _P]),
    "ax_r3_retain_thr": (
        C.c_int,
        [_P, _P, _P, C.c_int32, C.c_int64, C.c_int64, _P, C.c_int64, C.c_int32, _P, _P, C.c_size_t, _P],
    ),
    "ax_gather_points": (C.c_int, [_P, C.c_int32, C.c_int64, _P, C.c_int64, _P, C.c_int64, _P]),
    "ax_uniform_box_points": (
        C.c_int,
        [_P, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint64,
         C.c_uint64, _P],
    ),
    "ax_lbfgs_ws_floats": (C.c_size_t, [C.c_int64, C.c_int32]),
    "ax_lbfgs_iterate": (
        C.c_int,
        [_P, _P, _P, C.c_int64, _P, _P, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32, _P],
    ),
    "ax_adam_step": (
        C.c_int,
        [_P, _P, _P, _P, C.c_int64, _P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _P],
    ),
    "ax_residual_ws_bytes": (C.c_size_t, [C.POINTER(ResidualProgram), C.c_int64]),
    "ax_residual_forward": (C.c_int, [C.POINTER(ResidualProgram), _P, _P, C.c_int64, C.c_int64, _P, _P, C.c_size_t, _P]),
    "ax_residual_backward": (C.c_int, [C.POINTER(ResidualProgram), _P, _P, C.c_int64, C.c_int64, _P, _P, _P, _P,
                                       C.c_size_t, _P]),
    "ax_residual_eval": (C.c_int, [C.POINTER(ResidualProgram), _P, _P, C.c_int64, C.c_int64, _P, _P]),
    "ax_tc_set_option": (C.c_int, [C.c_char_p, C.c_int32]),
    "ax_debug_tc_wgrad": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, _P, _P]),
    "ax_debug_fma_peak": (C.c_double, [_P, C.c_int32, C.c_int32, _P]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def exported_symbols():
    return sorted(_SIGNATURES)


def load():
    """dlopen the in-tree library; raise loudly if it is not there (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m ai4xde_b200.build` "
            "(ai4xde_b200 has no CPU or PyTorch fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().ax_last_error().decode()


def _check(rc, what):
    if rc != 0:
        raise NativeLibraryError(f"{what} failed: {last_error()}")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NativeLibraryError(
                "ai4xde_b200 kernels need CUDA tensors (sm_100a); there is no CPU path"
            )


def make_net_desc(widths, activation, n_nets=1, first_activation=None):
    if len(widths) - 1 > MAX_LINEAR or len(widths) < 3:
        raise ValueError(f"layer_size {widths}: need 1..{MAX_LINEAR - 1} hidden layers")
    if activation not in ACT:
        raise NotImplementedError(
            f"activation {activation!r} is not in the fused path (tanh, sin, sigmoid are)"
        )
    d = NetDesc()
    d.n_linear = len(widths) - 1
    for i, w in enumerate(widths):
        d.widths[i] = int(w)
    d.activation = ACT[activation]
    d.n_nets = int(n_nets)
    if first_activation is not None and first_activation != activation:
        if first_activation not in ACT:
            raise NotImplementedError(f"first-layer activation {first_activation!r} is not in the fused path")
        d.first_activation = ACT[first_activation] + 1
    return d


def make_jet_desc(dirs, pairs):
    """dirs: input coordinate of each first-order stream; pairs: (a, b) indices into dirs."""
    d = JetDesc()
    d.n1 = len(dirs)
    for i, v in enumerate(dirs):
        d.dir1[i] = int(v)
    d.n2 = len(pairs)
    for q, (a, b) in enumerate(pairs):
        d.pa[q], d.pb[q] = int(a), int(b)
    return d


def param_count(net):
    return int(load().ax_param_count(C.byref(net)))


def jet_supported(jet):
    return bool(load().ax_jet_supported(C.byref(jet)))


def engine_supported(net, jet, engine):
    return bool(load().ax_engine_supported(C.byref(net), C.byref(jet), ENGINE[engine]))


def stash_bytes(net, jet, ldn, keep_all, engine="auto"):
    return int(load().ax_stash_bytes_engine(C.byref(net), C.byref(jet), int(ldn), int(bool(keep_all)), ENGINE[engine]))


def narrow_supported(net, jet):
    return bool(load().ax_narrow_supported(C.byref(net), C.byref(jet)))


def launch_count():
    return int(load().ax_launch_count())


def timing_enable(on):
    load().ax_timing_enable(int(bool(on)))


def timing_report():
    """{kernel: {launches, ms, flops, bytes}} for the launches since the last report."""
    import json

    lib = load()
    buf = C.create_string_buffer(1 << 16)  # one call: the report clears the records
    need = lib.ax_timing_report(buf, len(buf))
    if need > len(buf):
        raise NativeLibraryError("timing report truncated")
    return json.loads(buf.value.decode())


def pad32(n):
    return max(32, (int(n) + 31) // 32 * 32)


def _bytes_buf(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def jet_forward(net, jet, params, x_soa, n, keep_all, engine="auto"):
    """params [K, P] f32, x_soa [d, ldn] f32 -> (jets [K, S, C, ldn], stash)."""
    lib = load()
    _need_cuda(params, x_soa)
    ldn = x_soa.shape[1]
    S = 1 + jet.n1 + jet.n2
    C_out = net.widths[net.n_linear]
    dev = x_soa.device
    jets = torch.empty((net.n_nets, S, C_out, ldn), dtype=torch.float32, device=dev)
    stash = _bytes_buf(lib.ax_stash_bytes_engine(C.byref(net), C.byref(jet), ldn, int(keep_all), ENGINE[engine]), dev)
    wsb = lib.ax_forward_ws_bytes(C.byref(net), C.byref(jet), ldn)
    ws = _bytes_buf(wsb, dev)
    rc = lib.ax_jet_mlp_forward(
        C.byref(net), C.byref(jet), _ptr(params), _ptr(x_soa), int(n), int(ldn), _ptr(jets),
        _ptr(stash), int(keep_all), _ptr(ws), int(wsb), ENGINE[engine], _stream()
    )
    _check(rc, "ax_jet_mlp_forward")
    return jets, stash


def jet_backward(net, jet, params, x_soa, n, jets_bar, stash, engine="auto"):
    """-> grad [K, P] f32 of sum(jets * jets_bar) w.r.t. params."""
    lib = load()
    _need_cuda(params, x_soa, jets_bar, stash)
    ldn = x_soa.shape[1]
    dev = x_soa.device
    grad = torch.empty_like(params)
    wsb = lib.ax_backward_ws_bytes(C.byref(net), C.byref(jet), ldn)
    ws = _bytes_buf(wsb, dev)
    rc = lib.ax_jet_mlp_backward(
        C.byref(net), C.byref(jet), _ptr(params), _ptr(x_soa), int(n), int(ldn), _ptr(jets_bar),
        _ptr(stash), _ptr(grad), _ptr(ws), int(wsb), ENGINE[engine], _stream()
    )
    _check(rc, "ax_jet_mlp_backward")
    return grad


class ResidualTables:
    """Device tables of one residual.Program (code, segments, constants, array / scalar pointers, term scales) and
    the C struct that points at them.  Holds every tensor the program reads, so nothing it points at can be freed."""

    def __init__(self, program, device):
        import numpy as np

        i32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32).reshape(-1), device=device)
        code, segs, off = [], [], 0
        for beg, end, c in program.segments:
            code.append(c)
            segs.append((beg, end, off, off + len(c)))
            off += len(c)
        self.code = i32(np.concatenate(code, axis=0))
        self.segments = i32(np.array(segs))
        self.consts = torch.as_tensor(program.consts, dtype=torch.float32, device=device)
        self.term_scale = torch.as_tensor(program.scale if len(program.scale) else [0.0], dtype=torch.float32, device=device)
        self.aux = [t.detach().to(device=device, dtype=torch.float32).contiguous() for t, _ in program.aux]
        self.scalars = list(program.scalars)       # read through their storage at every step (never copied)
        for t in self.scalars:
            if not (t.is_cuda and t.dtype == torch.float32):
                raise NativeLibraryError("residual program scalars must be float32 CUDA tensors")
        ptrs = lambda ts: torch.tensor([t.data_ptr() for t in ts] or [0], dtype=torch.int64, device=device)
        self.aux_ptr, self.scalar_ptr = ptrs(self.aux), ptrs(self.scalars)
        self.aux_meta = i32(np.array([(first, t.shape[1]) for t, (_, first) in zip(self.aux, program.aux)] or [(0, 1)]))
        self.n_terms, self.n_streams = program.n_terms, program.S * program.C
        self.struct = ResidualProgram(
            _ptr(self.code), _ptr(self.segments), _ptr(self.consts), _ptr(self.aux_ptr), _ptr(self.aux_meta),
            _ptr(self.scalar_ptr), _ptr(self.term_scale), off, len(segs), program.n_reg, program.n_terms,
            self.n_streams, int(self.consts.numel()), len(self.scalars))
        self.n_scalars, self.scalar_grad = len(self.scalars), list(program.scalar_grad)


def residual_forward(tables, x_soa, jets, n):
    """per-term losses [T] of a residual program (csrc/residual_vm.cu); jets [1, S, C, ldn]"""
    lib = load()
    _need_cuda(x_soa, jets)
    ldn = x_soa.shape[1]
    losses = torch.empty(tables.n_terms, dtype=torch.float32, device=x_soa.device)
    wsb = lib.ax_residual_ws_bytes(C.byref(tables.struct), int(n))
    ws = _bytes_buf(wsb, x_soa.device)
    rc = lib.ax_residual_forward(C.byref(tables.struct), _ptr(x_soa), _ptr(jets), int(n), int(ldn), _ptr(losses),
                                 _ptr(ws), int(wsb), _stream())
    _check(rc, "ax_residual_forward")
    return losses


def residual_eval(tables, x_soa, jets, n, n_cols):
    """point-wise outputs [n_cols, ldn] of an operator program (rows n..ldn are left untouched: zero them if read)"""
    lib = load()
    _need_cuda(x_soa, jets)
    ldn = x_soa.shape[1]
    out = torch.zeros((n_cols, ldn), dtype=torch.float32, device=x_soa.device)
    rc = lib.ax_residual_eval(C.byref(tables.struct), _ptr(x_soa), _ptr(jets), int(n), int(ldn), _ptr(out), _stream())
    _check(rc, "ax_residual_eval")
    return out


def residual_backward(tables, x_soa, jets, n, term_grads):
    """(jets_bar [1, S, C, ldn] = sum_t term_grads[t] * d losses[t] / d jets, gradients w.r.t. the program's run-time
    scalars [n_scalars] or None when none of them is trainable)"""
    lib = load()
    term_grads = term_grads.to(torch.float32).contiguous()
    _need_cuda(x_soa, jets, term_grads)
    ldn = x_soa.shape[1]
    jets_bar = torch.empty_like(jets)
    sgrad = ws = None
    wsb = 0
    if tables.scalar_grad:
        sgrad = torch.empty(tables.n_scalars, dtype=torch.float32, device=x_soa.device)
        wsb = lib.ax_residual_ws_bytes(C.byref(tables.struct), int(ldn))
        ws = _bytes_buf(wsb, x_soa.device)
    rc = lib.ax_residual_backward(C.byref(tables.struct), _ptr(x_soa), _ptr(jets), int(n), int(ldn),
                                  _ptr(term_grads), _ptr(jets_bar), _ptr(sgrad), _ptr(ws), int(wsb), _stream())
    _check(rc, "ax_residual_backward")
    return jets_bar, sgrad


def tc_set_option(name, value):
    """run-time switch of the tcgen05 engine ("l2chain", "l2chain_group", "l2chain_hints", "l2chain_defer", "l2chain_wres")"""
    _check(load().ax_tc_set_option(name.encode(), int(value)), "ax_tc_set_option")


def fbpinn_window_supported(net, jet, engine):
    return bool(load().ax_fbpinn_window_supported(C.byref(net), C.byref(jet), ENGINE[engine]))


def fbpinn_window_forward(net, jet, params, x_soa, n, win_lo, win_hi, net_list, engine="auto"):
    """Streams [S, ldn] of sum_k sigmoid(x - lo_k) sigmoid(hi_k - x) N_k(x) over the nets in `net_list`
    (int32 device tensor of ids into the [K, P] bank), one launch (csrc/tc_bank.cu)."""
    lib = load()
    _need_cuda(params, x_soa, win_lo, win_hi, net_list)
    ldn = x_soa.shape[1]
    S = 1 + jet.n1 + jet.n2
    out = torch.empty((S, ldn), dtype=torch.float32, device=x_soa.device)
    wsb = lib.ax_fbpinn_window_ws_bytes(C.byref(net), C.byref(jet), ldn)
    ws = _bytes_buf(wsb, x_soa.device)
    rc = lib.ax_fbpinn_window_forward(C.byref(net), C.byref(jet), _ptr(params), _ptr(x_soa), int(n), int(ldn),
                                      _ptr(win_lo), _ptr(win_hi), _ptr(net_list), int(net_list.numel()), _ptr(out),
                                      _ptr(ws), int(wsb), ENGINE[engine], _stream())
    _check(rc, "ax_fbpinn_window_forward")
    return out


# ---- resampling ---------------------------------------------------------------------------
def resample_set_fast(on):
    """Parallel float64 sum / cumsum in the RAD selection instead of NumPy's sequential order (see the header)."""
    load().ax_resample_set_fast(int(bool(on)))


def resample_get_fast():
    return bool(load().ax_resample_get_fast())


def resample_ws(n, device):
    nbytes = load().ax_resample_ws_bytes(int(n))
    return _bytes_buf(nbytes, device), int(nbytes)


def pairwise_sum(a):
    """NumPy-order pairwise sum of a 1-D float32/float64 CUDA tensor -> 0-d tensor."""
    lib = load()
    _need_cuda(a)
    a = a.contiguous()
    out = torch.empty((1,), dtype=a.dtype, device=a.device)
    ws, wsb = resample_ws(a.numel(), a.device)
    fn = lib.ax_pairwise_sum_f32 if a.dtype == torch.float32 else lib.ax_pairwise_sum_f64
    _check(fn(_ptr(a), a.numel(), _ptr(out), _ptr(ws), wsb, _stream()), "ax_pairwise_sum")
    return out[0]


def rad_probabilities(res, n, k, c):
    """res [n_res, ldr] f32 residuals -> p [n] f64 (RAD.py:29-32 rounding order)."""
    lib = load()
    _need_cuda(res)
    p = torch.empty((n,), dtype=torch.float64, device=res.device)
    ws, wsb = resample_ws(n, res.device)
    rc = lib.ax_rad_probabilities(_ptr(res), res.shape[0], res.shape[1], int(n), float(k), float(c),
                                  _ptr(p), _ptr(ws), wsb, _stream())
    _check(rc, "ax_rad_probabilities")
    return p


def choice_round(p, u, found, n_found, n_found_out, ws, wsb):
    lib = load()
    _need_cuda(p, u, found, n_found_out)
    rc = lib.ax_choice_round(_ptr(p), p.numel(), _ptr(u), u.numel(), _ptr(found), int(n_found),
                             _ptr(n_found_out), _ptr(ws), wsb, _stream())
    _check(rc, "ax_choice_round")


def r3_retain(res, x_soa, n, place_at_end=True):
    """res [n_res, ldr] f32, x_soa [d, ldn] -> (x_out [d, ldn], n_keep (device i64), thr, res_sum)."""
    lib = load()
    _need_cuda(res, x_soa)
    dev = x_soa.device
    d, ldn = x_soa.shape
    x_out = torch.empty_like(x_soa)
    n_keep = torch.zeros((1,), dtype=torch.int64, device=dev)
    thr = torch.zeros((1,), dtype=torch.float32, device=dev)
    res_sum = torch.empty((n,), dtype=torch.float32, device=dev)
    ws, wsb = resample_ws(n, dev)
    rc = lib.ax_r3_retain(_ptr(res), res.shape[0], res.shape[1], _ptr(x_soa), d, int(n), ldn,
                          _ptr(x_out), ldn, int(bool(place_at_end)), _ptr(n_keep), _ptr(thr),
                          _ptr(res_sum), _ptr(ws), wsb, _stream())
    _check(rc, "ax_r3_retain")
    return x_out, n_keep, thr, res_sum


def r3_abs_sum(res, n):
    """res [n_res, ldr] f32 -> sum_r |res_r| [n] f32 (first half of the R3 retain step)."""
    lib = load()
    _need_cuda(res)
    out = torch.empty((n,), dtype=torch.float32, device=res.device)
    _check(lib.ax_r3_abs_sum(_ptr(res), res.shape[0], res.shape[1], int(n), _ptr(out), _stream()), "ax_r3_abs_sum")
    return out


def r3_retain_thr(res_sum, thr, x_soa, n, place_at_end=True):
    """Compaction half of the R3 retain step against a given threshold (0-d / 1-element f32 device tensor)."""
    lib = load()
    _need_cuda(res_sum, thr, x_soa)
    dev = x_soa.device
    d, ldn = x_soa.shape
    x_out = torch.empty_like(x_soa)
    n_keep = torch.zeros((1,), dtype=torch.int64, device=dev)
    thr = thr.reshape(1).to(torch.float32).contiguous()
    ws, wsb = resample_ws(n, dev)
    rc = lib.ax_r3_retain_thr(_ptr(res_sum), _ptr(thr), _ptr(x_soa), d, int(n), ldn, _ptr(x_out), ldn,
                              int(bool(place_at_end)), _ptr(n_keep), _ptr(ws), wsb, _stream())
    _check(rc, "ax_r3_retain_thr")
    return x_out, n_keep


def lbfgs_ws_floats(n_params, history):
    return int(load().ax_lbfgs_ws_floats(int(n_params), int(history)))


def lbfgs_iterate(x, g, loss, ws, state, history, lr, tol_grad, tol_change, begin_step):
    """One device-resident L-BFGS iteration on the flat parameter vector (csrc/lbfgs.cu)."""
    lib = load()
    _need_cuda(x, g, loss, ws, state)
    rc = lib.ax_lbfgs_iterate(_ptr(x), _ptr(g), _ptr(loss), x.numel(), _ptr(ws), _ptr(state), int(history),
                              float(lr), float(tol_grad), float(tol_change), int(begin_step), _stream())
    _check(rc, "ax_lbfgs_iterate")


def adam_step(params, grad, exp_avg, exp_avg_sq, state, lr, beta1, beta2, eps, weight_decay=0.0):
    """One Adam update of the flat float32 parameter vector, in place (csrc/adam.cu)."""
    lib = load()
    _need_cuda(params, grad, exp_avg, exp_avg_sq, state)
    rc = lib.ax_adam_step(_ptr(params), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), params.numel(), _ptr(state),
                          float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), _stream())
    _check(rc, "ax_adam_step")


def debug_tc_wgrad(zbar, a):
    """zbar, a: [rows, H] f32 CUDA -> zbar^T a through the tcgen05 weight-gradient kernel."""
    lib = load()
    _need_cuda(zbar, a)
    rows, H = zbar.shape
    out = torch.empty((H, H), dtype=torch.float32, device=zbar.device)
    ws = torch.empty((149 * H * H,), dtype=torch.float32, device=zbar.device)
    _check(lib.ax_debug_tc_wgrad(_ptr(zbar), _ptr(a), rows, H, _ptr(out), _ptr(ws), _stream()),
           "ax_debug_tc_wgrad")
    return out


def measure_fp32_fma_peak(seconds=0.3):
    """Measured FP32 FMA rate of this GPU in TFLOP/s (CUDA events around ax_debug_fma_peak, best of 3)."""
    lib = load()
    out = torch.zeros(1, device="cuda")
    props = torch.cuda.get_device_properties(out.device)
    blocks = props.multi_processor_count * 8
    iters = 1 << 14
    best = 0.0
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        flops = lib.ax_debug_fma_peak(_ptr(out), blocks, iters, _stream())
        e1.record()
        torch.cuda.synchronize()
        if flops <= 0:
            raise NativeLibraryError("ax_debug_fma_peak failed")
        ms = e0.elapsed_time(e1)
        if rep > 0:
            best = max(best, flops / (ms * 1e-3) / 1e12)
        if ms * 1e-3 < seconds / 4:
            iters = min(iters * 2, 1 << 20)
    return best


def uniform_box_points(x_soa, first, m, lo, hi, seed, offset):
    """Fill columns [first, first + m) of the SoA point buffer with Philox uniforms in the box [lo, hi)."""
    lib = load()
    _need_cuda(x_soa)
    d, ldn = x_soa.shape
    lo_a = (C.c_float * d)(*[float(v) for v in lo])
    hi_a = (C.c_float * d)(*[float(v) for v in hi])
    _check(lib.ax_uniform_box_points(_ptr(x_soa), d, ldn, int(first), int(m), lo_a, hi_a, int(seed) & (2**64 - 1),
                                     int(offset) & (2**64 - 1), _stream()), "ax_uniform_box_points")


def gather_points(x_soa, idx):
    lib = load()
    _need_cuda(x_soa, idx)
    d, ldn = x_soa.shape
    m = idx.numel()
    out = torch.zeros((d, pad32(m)), dtype=torch.float32, device=x_soa.device)
    _check(lib.ax_gather_points(_ptr(x_soa), d, ldn, _ptr(idx), m, _ptr(out), out.shape[1], _stream()),
           "ax_gather_points")
    return out
