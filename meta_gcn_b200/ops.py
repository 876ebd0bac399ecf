"""torch.library custom ops (namespace ``mgcn``) over the C-ABI of libmgcn.so.

Every op takes/returns torch CUDA tensors and forwards raw device pointers, sizes and the current
CUDA stream to the library; PyTorch only provides memory and streams.  CUDA-only: a CPU tensor
raises, there is no fallback (the CPU restatement lives in oracle/ and is test infrastructure).

Raw ops (no autograd) are registered as ``torch.ops.mgcn.*`` with fake (meta) implementations;
the differentiable wrappers used by the model classes are in ``meta_gcn_b200.functional``.
"""
import ctypes
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import MgcnCsr

DEFAULT_HUB_THRESHOLD = 256


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mgcn ops run on CUDA tensors only (no CPU fallback)")


def _f32c(t, name):
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _csr_struct(rowptr, nbr, perm, hub_rows, hub_count, hub_threshold):
    s = MgcnCsr()
    s.n_rows = rowptr.numel() - 1
    s.nnz_cap = nbr.numel()
    s.rowptr = rowptr.data_ptr()
    s.nbr = nbr.data_ptr() if nbr.numel() else None
    s.perm = perm.data_ptr() if perm is not None and perm.numel() else None
    if hub_rows is not None and hub_rows.numel() > 0:
        s.hub_rows = hub_rows.data_ptr()
        s.hub_count = hub_count.data_ptr()
        s.hub_cap = hub_rows.numel()
    else:
        s.hub_rows = None
        s.hub_count = None
        s.hub_cap = 0
    s.hub_threshold = int(hub_threshold)
    return s


# ------------------------------------------------------------------------------------------------
# implementations (plain functions; also what meta_gcn_b200.functional calls directly)
# ------------------------------------------------------------------------------------------------
def csr_build_impl(edge_index, N, by, loop_mode, hub_threshold):
    _need_cuda(edge_index)
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise TypeError("edge_index must be int64 [2,E]")
    ei = edge_index.contiguous()
    E = ei.size(1)
    N = int(N)
    dev = ei.device
    nnz_cap = E + (N if loop_mode == 2 else 0)
    hub_cap = nnz_cap // max(int(hub_threshold), 1) + 1
    rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    nbr = torch.empty(nnz_cap, dtype=torch.int32, device=dev)
    perm = torch.empty(nnz_cap, dtype=torch.int32, device=dev)
    hub_rows = torch.empty(hub_cap, dtype=torch.int32, device=dev)
    hub_count = torch.empty(1, dtype=torch.int32, device=dev)
    bad = torch.empty(1, dtype=torch.int32, device=dev)
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)
    args = (_ptr(ei), E, N, int(by), int(loop_mode), int(hub_threshold), _ptr(rowptr), _ptr(nbr),
            _ptr(perm), _ptr(hub_rows), hub_cap, _ptr(hub_count), _ptr(bad))
    _lib.check(lib.mgcn_csr_build(*args, None, ctypes.byref(nbytes), None))
    ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)
    _lib.check(lib.mgcn_csr_build(*args, _ptr(ws), ctypes.byref(nbytes), _stream()))
    return rowptr, nbr, perm, hub_rows, hub_count, bad


def degree_impl(rowptr):
    _need_cuda(rowptr)
    N = rowptr.numel() - 1
    deg = torch.empty(N, dtype=torch.float32, device=rowptr.device)
    _lib.check(_lib.load().mgcn_degree_from_rowptr(_ptr(rowptr), N, _ptr(deg), _stream()))
    return deg


def weighted_degree_impl(rowptr, nbr, perm, edge_weight, loop_weight):
    _need_cuda(rowptr, edge_weight)
    ew = _f32c(edge_weight, "edge_weight")
    s = _csr_struct(rowptr, nbr, perm, None, None, 0)
    deg = torch.empty(rowptr.numel() - 1, dtype=torch.float32, device=rowptr.device)
    _lib.check(_lib.load().mgcn_weighted_degree(ctypes.byref(s), _ptr(ew), ew.numel(),
                                                float(loop_weight), _ptr(deg), _stream()))
    return deg


def gcn_norm_impl(deg, mode):
    _need_cuda(deg)
    d = _f32c(deg, "deg")
    dis = torch.empty_like(d)
    _lib.check(_lib.load().mgcn_gcn_norm(_ptr(d), d.numel(), int(mode), _ptr(dis), _stream()))
    return dis


def permute_edge_values_impl(rowptr, nbr, perm, vals, loop_value):
    _need_cuda(rowptr, vals)
    v = _f32c(vals, "vals")
    s = _csr_struct(rowptr, nbr, perm, None, None, 0)
    out = torch.empty(nbr.numel(), dtype=torch.float32, device=rowptr.device)
    _lib.check(_lib.load().mgcn_permute_edge_values(ctypes.byref(s), _ptr(v), v.numel(),
                                                    float(loop_value), _ptr(out), _stream()))
    return out


def spmm_impl(rowptr, nbr, perm, hub_rows, hub_count, hub_threshold, x, gather_perm=False,
              edge_val=None, nbr_scale=None, row_scale=None, reduce=0, bias=None, residual=None,
              act=0):
    _need_cuda(rowptr, x, edge_val, nbr_scale, row_scale, bias, residual)
    if x.dim() != 2:
        raise ValueError("x must be [n_in, H]")
    x = _f32c(x, "x")
    edge_val = _f32c(edge_val, "edge_val")
    nbr_scale = _f32c(nbr_scale, "nbr_scale")
    row_scale = _f32c(row_scale, "row_scale")
    bias = _f32c(bias, "bias")
    residual = _f32c(residual, "residual")
    n_rows = rowptr.numel() - 1
    H = x.size(1)
    if edge_val is not None and edge_val.numel() != nbr.numel():
        raise ValueError("edge_val must be in row order with nnz_cap entries")
    if row_scale is not None and row_scale.numel() != n_rows:
        raise ValueError("row_scale must have one entry per row")
    if residual is not None and tuple(residual.shape) != (n_rows, H):
        raise ValueError("residual must be [n_rows, H]")
    if bias is not None and bias.numel() != H:
        raise ValueError("bias must have H entries")
    out = torch.empty(n_rows, H, dtype=torch.float32, device=x.device)
    s = _csr_struct(rowptr, nbr, perm, hub_rows, hub_count, hub_threshold)
    _lib.check(_lib.load().mgcn_spmm(ctypes.byref(s), _ptr(x), x.size(0), H, int(bool(gather_perm)),
                                     _ptr(edge_val), _ptr(nbr_scale), _ptr(row_scale), int(reduce),
                                     _ptr(bias), _ptr(residual), int(act), _ptr(out), _stream()))
    return out


def linear_impl(x, w, w_out_in, bias=None, add=None, act=0):
    """y = act(x @ W + bias + add); w is [Hi,Ho] (weight_node) or, if w_out_in, [Ho,Hi] (nn.Linear)."""
    _need_cuda(x, w, bias, add)
    x = _f32c(x, "x")
    w = _f32c(w, "w")
    bias = _f32c(bias, "bias")
    add = _f32c(add, "add")
    N, Hi = x.shape
    if w_out_in:
        Ho, Hi_w = w.shape
        sk, sc = 1, Hi
    else:
        Hi_w, Ho = w.shape
        sk, sc = Ho, 1
    if Hi_w != Hi:
        raise ValueError(f"width mismatch: x has {Hi}, weight expects {Hi_w}")
    y = torch.empty(N, Ho, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().mgcn_linear(_ptr(x), N, Hi, _ptr(w), sk, sc, Ho, _ptr(bias), _ptr(add),
                                       int(act), _ptr(y), _stream()))
    return y


def linear_wgrad_impl(x, g, w_out_in, want_bias):
    _need_cuda(x, g)
    x = _f32c(x, "x")
    g = _f32c(g, "g")
    N, Hi = x.shape
    Ho = g.size(1)
    dev = x.device
    if w_out_in:
        dw = torch.empty(Ho, Hi, dtype=torch.float32, device=dev)
        sk, sc = 1, Hi
    else:
        dw = torch.empty(Hi, Ho, dtype=torch.float32, device=dev)
        sk, sc = Ho, 1
    db = torch.empty(Ho if want_bias else 0, dtype=torch.float32, device=dev)
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)
    args = (_ptr(x), N, Hi, _ptr(g), Ho, _ptr(dw), sk, sc, _ptr(db) if want_bias else None)
    _lib.check(lib.mgcn_linear_wgrad(*args, None, ctypes.byref(nbytes), None))
    ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)
    _lib.check(lib.mgcn_linear_wgrad(*args, _ptr(ws), ctypes.byref(nbytes), _stream()))
    return dw, db


def relu_backward_impl(g, y):
    _need_cuda(g, y)
    g = _f32c(g, "g")
    y = _f32c(y, "y")
    out = torch.empty_like(g)
    _lib.check(_lib.load().mgcn_relu_backward(_ptr(g), _ptr(y), g.numel(), _ptr(out), _stream()))
    return out


def batch_to_offsets_impl(batch, G):
    _need_cuda(batch)
    if batch.dtype != torch.int64:
        raise TypeError("batch must be int64")
    b = batch.contiguous()
    off = torch.empty(int(G) + 1, dtype=torch.int32, device=b.device)
    _lib.check(_lib.load().mgcn_batch_to_offsets(_ptr(b), b.numel(), int(G), _ptr(off), _stream()))
    return off


def segment_reduce_impl(x, offsets, mode):
    _need_cuda(x, offsets)
    x = _f32c(x, "x")
    G = offsets.numel() - 1
    out = torch.empty(G, x.size(1), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().mgcn_segment_reduce(_ptr(x), x.size(1), _ptr(offsets), G, int(mode),
                                               _ptr(out), _stream()))
    return out


def segment_broadcast_impl(gout, offsets, N, mode):
    _need_cuda(gout, offsets)
    gout = _f32c(gout, "gout")
    G = offsets.numel() - 1
    dx = torch.empty(int(N), gout.size(1), dtype=torch.float32, device=gout.device)
    _lib.check(_lib.load().mgcn_segment_broadcast(_ptr(gout), gout.size(1), _ptr(offsets), G, int(N),
                                                  int(mode), _ptr(dx), _stream()))
    return dx


# ------------------------------------------------------------------------------------------------
# torch.library registration
# ------------------------------------------------------------------------------------------------
_LIBDEF = torch.library.Library("mgcn", "DEF")
_LIBDEF.define("csr_build(Tensor edge_index, int N, int by, int loop_mode, int hub_threshold) -> "
               "(Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)")
_LIBDEF.define("degree(Tensor rowptr) -> Tensor")
_LIBDEF.define("weighted_degree(Tensor rowptr, Tensor nbr, Tensor perm, Tensor edge_weight, "
               "float loop_weight) -> Tensor")
_LIBDEF.define("gcn_norm(Tensor deg, int mode) -> Tensor")
_LIBDEF.define("permute_edge_values(Tensor rowptr, Tensor nbr, Tensor perm, Tensor vals, "
               "float loop_value) -> Tensor")
_LIBDEF.define("spmm(Tensor rowptr, Tensor nbr, Tensor perm, Tensor hub_rows, Tensor hub_count, "
               "int hub_threshold, Tensor x, bool gather_perm, Tensor? edge_val, Tensor? nbr_scale, "
               "Tensor? row_scale, int reduce, Tensor? bias, Tensor? residual, int act) -> Tensor")
_LIBDEF.define("linear(Tensor x, Tensor w, bool w_out_in, Tensor? bias, Tensor? add, int act) -> Tensor")
_LIBDEF.define("linear_wgrad(Tensor x, Tensor g, bool w_out_in, bool want_bias) -> (Tensor, Tensor)")
_LIBDEF.define("relu_backward(Tensor g, Tensor y) -> Tensor")
_LIBDEF.define("batch_to_offsets(Tensor batch, int G) -> Tensor")
_LIBDEF.define("segment_reduce(Tensor x, Tensor offsets, int mode) -> Tensor")
_LIBDEF.define("segment_broadcast(Tensor gout, Tensor offsets, int N, int mode) -> Tensor")

_IMPLS = {
    "csr_build": csr_build_impl,
    "degree": degree_impl,
    "weighted_degree": weighted_degree_impl,
    "gcn_norm": gcn_norm_impl,
    "permute_edge_values": permute_edge_values_impl,
    "spmm": spmm_impl,
    "linear": linear_impl,
    "linear_wgrad": linear_wgrad_impl,
    "relu_backward": relu_backward_impl,
    "batch_to_offsets": batch_to_offsets_impl,
    "segment_reduce": segment_reduce_impl,
    "segment_broadcast": segment_broadcast_impl,
}
for _name, _fn in _IMPLS.items():
    _LIBDEF.impl(_name, _fn, "CUDA")


def _cpu_refusal(name):
    def _raise(*args, **kwargs):
        raise RuntimeError(f"mgcn::{name} has no CPU implementation (CUDA sm_100a only)")
    return _raise


for _name in _IMPLS:
    _LIBDEF.impl(_name, _cpu_refusal(_name), "CPU")


# fake (meta) implementations so the ops trace under FakeTensor / torch.export
@torch.library.register_fake("mgcn::csr_build")
def _(edge_index, N, by, loop_mode, hub_threshold):
    E = edge_index.size(1)
    cap = E + (N if loop_mode == 2 else 0)
    i32 = dict(dtype=torch.int32, device=edge_index.device)
    return (torch.empty(N + 1, **i32), torch.empty(cap, **i32), torch.empty(cap, **i32),
            torch.empty(cap // max(hub_threshold, 1) + 1, **i32), torch.empty(1, **i32),
            torch.empty(1, **i32))


@torch.library.register_fake("mgcn::degree")
def _(rowptr):
    return torch.empty(rowptr.numel() - 1, dtype=torch.float32, device=rowptr.device)


@torch.library.register_fake("mgcn::weighted_degree")
def _(rowptr, nbr, perm, edge_weight, loop_weight):
    return torch.empty(rowptr.numel() - 1, dtype=torch.float32, device=rowptr.device)


@torch.library.register_fake("mgcn::gcn_norm")
def _(deg, mode):
    return torch.empty_like(deg)


@torch.library.register_fake("mgcn::permute_edge_values")
def _(rowptr, nbr, perm, vals, loop_value):
    return torch.empty(nbr.numel(), dtype=torch.float32, device=rowptr.device)


@torch.library.register_fake("mgcn::spmm")
def _(rowptr, nbr, perm, hub_rows, hub_count, hub_threshold, x, gather_perm, edge_val, nbr_scale,
      row_scale, reduce, bias, residual, act):
    return torch.empty(rowptr.numel() - 1, x.size(1), dtype=torch.float32, device=x.device)


@torch.library.register_fake("mgcn::linear")
def _(x, w, w_out_in, bias, add, act):
    Ho = w.size(0) if w_out_in else w.size(1)
    return torch.empty(x.size(0), Ho, dtype=torch.float32, device=x.device)


@torch.library.register_fake("mgcn::linear_wgrad")
def _(x, g, w_out_in, want_bias):
    Hi, Ho = x.size(1), g.size(1)
    shape = (Ho, Hi) if w_out_in else (Hi, Ho)
    return (torch.empty(shape, dtype=torch.float32, device=x.device),
            torch.empty(Ho if want_bias else 0, dtype=torch.float32, device=x.device))


@torch.library.register_fake("mgcn::relu_backward")
def _(g, y):
    return torch.empty_like(g)


@torch.library.register_fake("mgcn::batch_to_offsets")
def _(batch, G):
    return torch.empty(G + 1, dtype=torch.int32, device=batch.device)


@torch.library.register_fake("mgcn::segment_reduce")
def _(x, offsets, mode):
    return torch.empty(offsets.numel() - 1, x.size(1), dtype=torch.float32, device=x.device)


@torch.library.register_fake("mgcn::segment_broadcast")
def _(gout, offsets, N, mode):
    return torch.empty(N, gout.size(1), dtype=torch.float32, device=gout.device)


OP_NAMES = tuple(_IMPLS)
