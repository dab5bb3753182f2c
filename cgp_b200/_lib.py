"""ctypes binding of the C-ABI declared in include/cgp_b200.h.

There is NO CPU fallback: importing this module without the built extension, or calling it without a
CUDA device, raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_cgp_b200.so")

KIND = {"matern32": 0, "rbf": 1}

_vp, _i64, _i32, _f64, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_size_t
_pi64 = C.POINTER(C.c_int64)
_pi32 = C.POINTER(C.c_int32)

# name -> (restype, argtypes) — one entry per symbol of include/cgp_b200.h
SIGNATURES = {
    "cgp_version": (C.c_int, []),
    "cgp_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "cgp_destroy": (C.c_int, [_vp]),
    "cgp_padded_n": (_i64, [_i64]),
    "cgp_model_elems": (_i64, [_pi64, C.c_int]),
    "cgp_kernel_matrix": (C.c_int, [_vp, C.c_int, _vp, _i64, C.c_int, _vp, _f64, _vp, _vp]),
    "cgp_kernel_cross": (C.c_int, [_vp, C.c_int, _vp, _i64, _vp, _i64, C.c_int, _vp, _vp, _vp]),
    "cgp_lml_workspace_bytes": (_sz, [_pi64, C.c_int, _pi32, C.c_int, C.c_int]),
    "cgp_lml_grad_batched": (C.c_int, [_vp, C.c_int, _vp, _vp, _pi64, C.c_int, C.c_int, _pi32, C.c_int, _vp, _f64,
                                       _vp, _vp, _vp, _vp, _sz, _vp]),
    "cgp_factorize_workspace_bytes": (_sz, [_pi64, C.c_int, C.c_int]),
    "cgp_factorize": (C.c_int, [_vp, C.c_int, _vp, _vp, _pi64, C.c_int, C.c_int, _vp, _f64, _vp, _vp, _vp, _vp, _vp,
                                _vp, _sz, _vp]),
    "cgp_knn_labels": (C.c_int, [_vp, _vp, _vp, _i64, C.c_int, C.c_int, C.c_int, _vp, _i64, _vp, _vp]),
    "cgp_score_workspace_bytes": (_sz, [_pi64, C.c_int, C.c_int, _i64]),
    "cgp_predict": (C.c_int, [_vp, C.c_int, _vp, _pi64, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                              _vp, _vp, _vp, _sz, _vp]),
    "cgp_ei_argmax": (C.c_int, [_vp, C.c_int, _vp, _pi64, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cgp_ei_argmax_pruned": (C.c_int, [_vp, C.c_int, _vp, _pi64, C.c_int, C.c_int, _vp, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                       _i64, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cgp_append_workspace_bytes": (_sz, [_i64]),
    "cgp_append_sample": (C.c_int, [_vp, C.c_int, _vp, _i64, C.c_int, _vp, _f64, _vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "cgp_refresh_alpha": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cgp_quadform_workspace_bytes": (_sz, [_pi64, C.c_int, C.c_int, _i64]),
    "cgp_quadform_scan": (C.c_int, [_vp, C.c_int, _vp, _pi64, C.c_int, C.c_int, _vp, _vp, C.POINTER(C.c_double), _vp, _vp, _i64,
                                    _i64, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cgp_rowdot": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "cgp_argmax_sorted": (C.c_int, [_vp, _vp, _i64, _vp, _vp, C.c_int, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cgp_dgemm_nt": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "cgp_profile_classes": (C.c_int, []),
    "cgp_profile_class_name": (C.c_char_p, [C.c_int]),
    "cgp_profile_enable": (C.c_int, [_vp, C.c_int]),
    "cgp_profile_read": (C.c_int, [_vp, C.POINTER(C.c_double), _pi64, C.c_int, C.c_int]),
    "cgp_profile_flops": (C.c_int, [_vp, C.POINTER(C.c_double), _pi64, C.c_int]),
    "cgp_profile_mute": (C.c_int, [_vp, C.c_int]),
    "cgp_set_option": (C.c_int, [_vp, C.c_int, C.c_int]),
    "cgp_get_option": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int)]),
}

_lib = None
_handles = {}


class CgpError(RuntimeError):
    pass


def lib():
    """Load the shared library (no compute, usable without a GPU)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise CgpError(f"cgp_b200 CUDA extension not built: {SO_PATH} missing "
                           "(run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def handle(device=None, lane=0):
    """The library handle of (device, lane).  Lane 0 is the default; further lanes are independent handles (their own
    internal streams, events and metadata ring) for batches that run CONCURRENTLY on the same device (engine.fit)."""
    if not torch.cuda.is_available():
        raise CgpError("cgp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.cuda.current_device() if device is None else int(device)
    key = dev if lane == 0 else (dev, int(lane))
    if key not in _handles:
        h = _vp()
        with torch.cuda.device(dev):
            _chk(lib().cgp_create(C.byref(h), dev), "cgp_create")
        _handles[key] = h
    return _handles[key]


def lanes_of(device=None):
    """Lane ids of the handles that exist on `device`."""
    dev = torch.cuda.current_device() if device is None else int(device)
    return [0 if k == dev else k[1] for k in _handles if k == dev or (isinstance(k, tuple) and k[0] == dev)]


def _chk(rc, what):
    if rc != 0:
        kind = f"argument {-rc} invalid" if rc < 0 else f"CUDA error {rc}"
        raise CgpError(f"{what} failed: {kind}")


def _ptr(t, dtype=torch.float64):
    if t is None:
        return None
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous() and t.dtype == dtype):
        raise CgpError(f"expected a contiguous CUDA tensor of dtype {dtype}, got "
                       f"{type(t).__name__} {getattr(t, 'dtype', None)} {getattr(t, 'device', None)}")
    return t.data_ptr()


def _stream(device=None):
    """Torch's current stream ON THE DEVICE OF THE TENSORS (not the process' current device): the C entry points run on
    the handle's device whatever the caller's current device is."""
    return torch.cuda.current_stream(device).cuda_stream


def _host_i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_pi64)


def _host_i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_pi32)


def padded_n(n):
    return int(lib().cgp_padded_n(int(n)))


def model_elems(offs):
    o, po = _host_i64(offs)
    return int(lib().cgp_model_elems(po, len(o) - 1))


# ------------------------------------------------------------------------------------------------
def kernel_matrix(X, theta, kind="matern32", alpha=0.0):
    n, d = X.shape
    K = torch.empty((n, n), dtype=torch.float64, device=X.device)
    _chk(lib().cgp_kernel_matrix(handle(X.device.index), KIND[kind], _ptr(X), n, d, _ptr(theta), float(alpha), _ptr(K),
                                 _stream(X.device)), "cgp_kernel_matrix")
    return K


def kernel_cross(Q, X, theta, kind="matern32"):
    m, d = Q.shape
    n = X.shape[0]
    K = torch.empty((m, n), dtype=torch.float64, device=X.device)
    _chk(lib().cgp_kernel_cross(handle(X.device.index), KIND[kind], _ptr(Q), m, _ptr(X), n, d, _ptr(theta), _ptr(K),
                                _stream(X.device)), "cgp_kernel_cross")
    return K


def dgemm_nt(A, B, out=None):
    M, K = A.shape
    N = B.shape[0]
    Cm = out if out is not None else torch.empty((M, N), dtype=torch.float64, device=A.device)
    _chk(lib().cgp_dgemm_nt(handle(A.device.index), _ptr(A), _ptr(B), _ptr(Cm), M, N, K, _stream(A.device)), "cgp_dgemm_nt")
    return Cm


def lml_workspace_bytes(offs, pair_cluster, d):
    o, po = _host_i64(offs)
    pc, ppc = _host_i32(pair_cluster)
    return int(lib().cgp_lml_workspace_bytes(po, len(o) - 1, ppc, len(pc), d))


def lml_grad_batched(X, y, offs, pair_cluster, theta, alpha, kind, workspace, want_grad=True, out=None, lane=0):
    """theta [B,P] CUDA -> (lml [B], grad [B,P] | None, info [B]) CUDA tensors (asynchronous).
    `out` = (lml, grad, info) preallocated contiguous CUDA tensors of exactly these shapes (no allocation then)."""
    o, po = _host_i64(offs)
    pc, ppc = _host_i32(pair_cluster)
    B, P = theta.shape
    d = X.shape[1]
    assert P == d + 1 and len(pc) == B
    if out is not None:
        lml, grad, info = out
        if not want_grad:
            grad = None
        assert lml.shape == (B,) and info.shape == (B,) and (grad is None or grad.shape == (B, P))
    else:
        lml = torch.empty(B, dtype=torch.float64, device=X.device)
        grad = torch.empty((B, P), dtype=torch.float64, device=X.device) if want_grad else None
        info = torch.empty(B, dtype=torch.int32, device=X.device)
    _chk(lib().cgp_lml_grad_batched(handle(X.device.index, lane), KIND[kind], _ptr(X), _ptr(y), po, len(o) - 1, d, ppc, B,
                                    _ptr(theta), float(alpha), _ptr(lml), _ptr(grad), _ptr(info, torch.int32),
                                    _ptr(workspace, torch.uint8), workspace.numel(), _stream(X.device)),
         "cgp_lml_grad_batched")
    return lml, grad, info


def factorize_workspace_bytes(offs, d):
    o, po = _host_i64(offs)
    return int(lib().cgp_factorize_workspace_bytes(po, len(o) - 1, d))


def factorize(X, y, offs, theta, alpha, kind, workspace=None):
    """theta [C,P] -> dict(L, W (padded layout), alpha_vec [n], lml [C], info [C])."""
    o, po = _host_i64(offs)
    Cn = len(o) - 1
    d = X.shape[1]
    elems = model_elems(o)
    dev = X.device
    L = torch.empty(elems, dtype=torch.float64, device=dev)
    W = torch.empty(elems, dtype=torch.float64, device=dev)
    av = torch.empty(X.shape[0], dtype=torch.float64, device=dev)
    lml = torch.empty(Cn, dtype=torch.float64, device=dev)
    info = torch.empty(Cn, dtype=torch.int32, device=dev)
    if workspace is None:
        workspace = torch.empty(factorize_workspace_bytes(o, d), dtype=torch.uint8, device=dev)
    _chk(lib().cgp_factorize(handle(dev.index), KIND[kind], _ptr(X), _ptr(y), po, Cn, d, _ptr(theta), float(alpha),
                             _ptr(L), _ptr(W), _ptr(av), _ptr(lml), _ptr(info, torch.int32),
                             _ptr(workspace, torch.uint8), workspace.numel(), _stream(dev)), "cgp_factorize")
    return dict(L=L, W=W, alpha_vec=av, lml=lml, info=info)


def append_workspace_bytes(n_new):
    return int(lib().cgp_append_workspace_bytes(int(n_new)))


def append_sample(Xc, theta_c, alpha, kind, L_c, W_c, workspace):
    """Extend the factor blocks L_c / W_c ([npad, npad] views, old factor in the leading block) by the LAST row of Xc.
    -> info (1-element int32 CUDA tensor)."""
    n_new, d = Xc.shape
    npad = L_c.shape[0]
    info = torch.empty(1, dtype=torch.int32, device=Xc.device)
    _chk(lib().cgp_append_sample(handle(Xc.device.index), KIND[kind], _ptr(Xc), n_new, d, _ptr(theta_c), float(alpha),
                                 _ptr(L_c), _ptr(W_c), npad, _ptr(info, torch.int32), _ptr(workspace, torch.uint8),
                                 workspace.numel(), _stream(Xc.device)), "cgp_append_sample")
    return info


def refresh_alpha(L_c, W_c, n, y_c, alpha_out, workspace):
    """alpha_out[:n] = K^-1 y_c for one factorised cluster; -> lml (1-element CUDA tensor)."""
    lml = torch.empty(1, dtype=torch.float64, device=y_c.device)
    _chk(lib().cgp_refresh_alpha(handle(y_c.device.index), _ptr(L_c), _ptr(W_c), L_c.shape[0], int(n), _ptr(y_c),
                                 _ptr(alpha_out), _ptr(lml), _ptr(workspace, torch.uint8), workspace.numel(), _stream(y_c.device)),
         "cgp_refresh_alpha")
    return lml


def rowdot(A, B, n):
    """out[r] = sum_{k<n} A[r,k] B[r,k] for two contiguous [rows, ld] CUDA matrices."""
    rows, ld = A.shape
    assert B.shape == A.shape
    out = torch.empty(rows, dtype=torch.float64, device=A.device)
    _chk(lib().cgp_rowdot(handle(A.device.index), _ptr(A), _ptr(B), rows, int(n), ld, _ptr(out), _stream(A.device)), "cgp_rowdot")
    return out


def argmax_sorted(score_sorted, perm, coff, idx_base=0):
    """First maximum in bucket order -> (score, candidate index + idx_base, cluster) 1-element CUDA tensors."""
    dev = score_sorted.device
    bs = torch.empty(1, dtype=torch.float64, device=dev)
    bi = torch.empty(1, dtype=torch.int64, device=dev)
    bc = torch.empty(1, dtype=torch.int64, device=dev)
    scratch = torch.empty(16384, dtype=torch.uint8, device=dev)
    _chk(lib().cgp_argmax_sorted(handle(dev.index), _ptr(score_sorted), score_sorted.numel(), _ptr(perm, torch.int64),
                                 _ptr(coff, torch.int64), coff.numel() - 1, int(idx_base), _ptr(bs), _ptr(bi, torch.int64),
                                 _ptr(bc, torch.int64), _ptr(scratch, torch.uint8), scratch.numel(), _stream(dev)),
         "cgp_argmax_sorted")
    return bs, bi, bc


def quadform_workspace_bytes(offs, d, m):
    o, po = _host_i64(offs)
    return int(lib().cgp_quadform_workspace_bytes(po, len(o) - 1, d, int(m)))


def quadform_scan(X, offs, theta, Bmat, c0, Q, qlabel, kind, idx_base=0, penalty=None, want_values=False, workspace=None):
    """Dense MSPE scan (cgp_quadform_scan) -> ((min value, idx, cluster) 1-element CUDA tensors, values [m] | None)."""
    o, po = _host_i64(offs)
    d = X.shape[1]
    m = Q.shape[0]
    dev = X.device
    bs = torch.empty(1, dtype=torch.float64, device=dev)
    bi = torch.empty(1, dtype=torch.int64, device=dev)
    bc = torch.empty(1, dtype=torch.int64, device=dev)
    val = torch.empty(m, dtype=torch.float64, device=dev) if want_values else None
    if workspace is None:
        workspace = torch.empty(quadform_workspace_bytes(o, d, m), dtype=torch.uint8, device=dev)
    c0 = np.ascontiguousarray(c0, dtype=np.float64)
    _chk(lib().cgp_quadform_scan(handle(dev.index), KIND[kind], _ptr(X), po, len(o) - 1, d, _ptr(theta), _ptr(Bmat),
                                 c0.ctypes.data_as(C.POINTER(C.c_double)), _ptr(Q), _ptr(qlabel, torch.int64), m,
                                 int(idx_base), _ptr(penalty), _ptr(val), _ptr(bs), _ptr(bi, torch.int64),
                                 _ptr(bc, torch.int64), _ptr(workspace, torch.uint8), workspace.numel(), _stream(dev)),
         "cgp_quadform_scan")
    return (-bs, bi, bc), val


def knn_labels(X, labels, Q, k, n_classes):
    n, d = X.shape
    m = Q.shape[0]
    out = torch.empty(m, dtype=torch.int64, device=X.device)
    if m:
        _chk(lib().cgp_knn_labels(handle(X.device.index), _ptr(X), _ptr(labels, torch.int64), n, d, int(k),
                                  int(n_classes), _ptr(Q), m, _ptr(out, torch.int64), _stream(X.device)), "cgp_knn_labels")
    return out


def score_workspace_bytes(offs, d, m):
    o, po = _host_i64(offs)
    return int(lib().cgp_score_workspace_bytes(po, len(o) - 1, d, int(m)))


def predict(X, offs, theta, W, alpha_vec, y_mean, y_std, Q, qlabel, kind, want_std=True, workspace=None):
    o, po = _host_i64(offs)
    d = X.shape[1]
    m = Q.shape[0]
    dev = X.device
    mean = torch.empty(m, dtype=torch.float64, device=dev)
    std = torch.empty(m, dtype=torch.float64, device=dev) if want_std else None
    if m == 0:
        return mean, std
    if workspace is None:
        workspace = torch.empty(score_workspace_bytes(o, d, m), dtype=torch.uint8, device=dev)
    _chk(lib().cgp_predict(handle(dev.index), KIND[kind], _ptr(X), po, len(o) - 1, d, _ptr(theta), _ptr(W),
                           _ptr(alpha_vec), _ptr(y_mean), _ptr(y_std), _ptr(Q), _ptr(qlabel, torch.int64), m,
                           _ptr(mean), _ptr(std), _ptr(workspace, torch.uint8), workspace.numel(), _stream(dev)),
         "cgp_predict")
    return mean, std


def ei_argmax(X, offs, theta, W, alpha_vec, y_mean, y_std, y_max, Q, qlabel, kind, idx_base=0, want_arrays=False,
              workspace=None, prune=False, penalty=None, alpha=None):
    """-> (best [score f64, idx i64, cluster i64] as CUDA tensors, score/mean/std arrays or None).
    prune=True: exact bound-pruned scan (cgp_ei_argmax_pruned): same selection, no per-candidate arrays.
    penalty: [m] CUDA f64 additive soft-constraint term (score = (EI + penalty) / n_c) or None."""
    if penalty is not None and penalty.shape != (Q.shape[0],):
        raise CgpError("penalty must have one entry per candidate")
    o, po = _host_i64(offs)
    d = X.shape[1]
    m = Q.shape[0]
    dev = X.device
    bs = torch.empty(1, dtype=torch.float64, device=dev)
    bi = torch.empty(1, dtype=torch.int64, device=dev)
    bc = torch.empty(1, dtype=torch.int64, device=dev)
    score = mean = std = None
    if want_arrays:
        score = torch.empty(m, dtype=torch.float64, device=dev)
        mean = torch.empty(m, dtype=torch.float64, device=dev)
        std = torch.empty(m, dtype=torch.float64, device=dev)
    if workspace is None:
        workspace = torch.empty(score_workspace_bytes(o, d, m), dtype=torch.uint8, device=dev)
    if prune:
        if want_arrays:
            raise CgpError("the pruned scan has no per-candidate outputs (pruned candidates are never scored exactly)")
        if alpha is None:
            raise CgpError("the pruned scan needs the jitter `alpha` the model was factorised with (single-neighbour bound)")
        _chk(lib().cgp_ei_argmax_pruned(handle(dev.index), KIND[kind], _ptr(X), po, len(o) - 1, d, _ptr(theta), float(alpha), _ptr(W),
                                        _ptr(alpha_vec), _ptr(y_mean), _ptr(y_std), _ptr(y_max), _ptr(Q),
                                        _ptr(qlabel, torch.int64), m, int(idx_base), _ptr(penalty), _ptr(bs),
                                        _ptr(bi, torch.int64), _ptr(bc, torch.int64), _ptr(workspace, torch.uint8),
                                        workspace.numel(),
                                        _stream(dev)), "cgp_ei_argmax_pruned")
        return (bs, bi, bc), (None, None, None)
    _chk(lib().cgp_ei_argmax(handle(dev.index), KIND[kind], _ptr(X), po, len(o) - 1, d, _ptr(theta), _ptr(W),
                             _ptr(alpha_vec), _ptr(y_mean), _ptr(y_std), _ptr(y_max), _ptr(Q),
                             _ptr(qlabel, torch.int64), m, int(idx_base), _ptr(penalty), _ptr(score), _ptr(mean), _ptr(std),
                             _ptr(bs), _ptr(bi, torch.int64), _ptr(bc, torch.int64), _ptr(workspace, torch.uint8),
                             workspace.numel(), _stream(dev)), "cgp_ei_argmax")
    return (bs, bi, bc), (score, mean, std)


OPTIONS = {"knn_mode": 1, "fork_max": 2, "tile_mode": 3}
KNN_MODES = {"auto": 0, "exact": 1, "tc": 2}


def set_option(name, value, device=None):
    """Runtime switch on every lane handle of `device` (include/cgp_b200.h: cgp_set_option).  Returns the old value."""
    if name == "knn_mode" and isinstance(value, str):
        value = KNN_MODES[value]
    old = C.c_int()
    _chk(lib().cgp_get_option(handle(device), OPTIONS[name], C.byref(old)), "cgp_get_option")
    for lane in lanes_of(device):
        _chk(lib().cgp_set_option(handle(device, lane), OPTIONS[name], int(value)), "cgp_set_option")
    return old.value


# ------------------------------------------------------------------------------------------------ measurement
def profile_enable(on=True, device=None):
    """on: False/0 off, True/1 every launch timed, 2 only the large kernels (see include/cgp_b200.h).  All lanes."""
    handle(device)
    for lane in lanes_of(device):
        _chk(lib().cgp_profile_enable(handle(device, lane), int(on)), "cgp_profile_enable")


def profile_mute(on, device=None, lane=0):
    """Count but do not time the following launches of this lane (set while another lane runs concurrently: an event
    bracket would then include the other lane's work)."""
    _chk(lib().cgp_profile_mute(handle(device, lane), 1 if on else 0), "cgp_profile_mute")


def profile_read(reset=True, device=None, detail=False):
    """-> {class name: (milliseconds, launches)} accumulated since the last reset (synchronises).
    detail=True: {class: dict(ms, launches, timed, flops)} — ms / flops / timed cover the launches that were bracketed by
    events (exclusive kernel time: batches on the three-stream schedule are counted in `launches` but not timed)."""
    L = lib()
    n = L.cgp_profile_classes()
    names = [L.cgp_profile_class_name(i).decode() for i in range(n)]
    tot = {nm: dict(ms=0.0, launches=0, timed=0, flops=0.0) for nm in names}
    handle(device)
    for lane in lanes_of(device):  # totals over all lanes of the device
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        fl = (C.c_double * n)()
        timed = (C.c_int64 * n)()
        _chk(L.cgp_profile_flops(handle(device, lane), fl, timed, n), "cgp_profile_flops")
        _chk(L.cgp_profile_read(handle(device, lane), ms, cnt, n, 1 if reset else 0), "cgp_profile_read")
        for i, nm in enumerate(names):
            t = tot[nm]
            t["ms"] += float(ms[i]); t["launches"] += int(cnt[i]); t["timed"] += int(timed[i]); t["flops"] += float(fl[i])
    if detail:
        return tot
    return {nm: (t["ms"], t["launches"]) for nm, t in tot.items()}
