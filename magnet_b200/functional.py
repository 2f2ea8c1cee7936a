"""Autograd bindings of the C-ABI kernels.  Forward and backward both run hand-written CUDA;
nothing here computes on the CPU or through PyTorch ops (torch is used for allocation only).
Backward runs on PyTorch's autograd thread: every call passes the current stream explicitly.
"""
from typing import Optional

import torch

from . import _lib
from .graph import AggregationPlan, GraphSegments

ACT = {"none": 0, "relu": 1, "swish": 2}

# Arithmetic of the per-edge contractions (include/magnet_b200.h, `precision`):
#   "fp32"      0  fp32 FFMA
#   "fp32_tc"   1  tcgen05 tensor cores, bf16 hi/lo split + fp32 accumulation (1e-5 contract)
#   "bf16"      2  tcgen05 tensor cores, plain bf16 operands (1e-2 contract)
PRECISIONS = {"fp32": 0, "fp32_tc": 1, "bf16": 2}
_precision = "fp32_tc"


def set_precision(name: str) -> str:
    """Select the edge-kernel arithmetic; returns the previous setting."""
    global _precision
    if name not in PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
    old, _precision = _precision, name
    return old


def get_precision() -> str:
    return _precision


def precision_key() -> tuple:
    """Everything that selects kernels globally (part of the key of captured CUDA graphs)."""
    return (_precision, _linear_tc)


def _empty(shape, ref):
    return torch.empty(shape, dtype=torch.float32, device=ref.device)


# --------------------------------------------------------------------------------------------
# GNN_Layer (models/mpnn_2d.py:27-90)
# --------------------------------------------------------------------------------------------
@_lib.guard
def pack_gnn_layer(W1, b1, W2, W3, W4, tw: int, dp: int, nv: int) -> torch.Tensor:
    """Packed / transposed / bf16-imaged copies of a layer's weights for the kernels (the nn.Parameters stay the
    canonical [out, in] fp32 tensors).  GNN_Layer caches the result per module, keyed on the parameters' version counters."""
    L = _lib.lib()
    packed = _empty((L.mgb_gnn_layer_packed_floats(tw, dp, nv),), W1)
    _lib.check(L.mgb_gnn_layer_pack(_lib.ptr(W1), _lib.ptr(b1), _lib.ptr(W2), _lib.ptr(W3), _lib.ptr(W4), tw, dp, nv,
                                    _lib.ptr(packed), _lib.stream()), "gnn_layer_pack")
    return packed


class GNNLayerFn(torch.autograd.Function):
    """y = InstanceNorm(x + update(x, mean_j message(x_i, x_j, u, pos, var)))"""

    @staticmethod
    def forward(ctx, x, u, pos, var, W1, b1, W2, b2, W3, b3, W4, b4, plan: AggregationPlan, seg: GraphSegments, packed=None):
        _lib.require_cuda(x, u, pos, var, W1)
        L = _lib.lib()
        x, u, pos, var = (_lib.f32c(t) for t in (x, u, pos, var))
        W1, b1, W2, b2, W3, b3, W4, b4 = (_lib.f32c(t.detach()) for t in (W1, b1, W2, b2, W3, b3, W4, b4))
        N, H = x.shape
        tw, dp, nv = u.shape[1], pos.shape[1], var.shape[1]
        if H != 128 or W2.shape != (128, 128) or W1.shape != (128, 256 + tw + dp + nv) or W3.shape != (128, 256 + nv):
            raise RuntimeError("GNN_Layer kernels are built for in=out=hidden=128 and message_net_1 of width "
                               f"256+tw+dp+nv; got x {tuple(x.shape)}, W1 {tuple(W1.shape)}, W3 {tuple(W3.shape)}")
        if plan.n_nodes != N:
            raise RuntimeError("aggregation plan was built for a different node count")
        if packed is None:
            packed = pack_gnn_layer(W1, b1, W2, W3, W4, tw, dp, nv)
        if W4.shape != (128, 128) or b2.shape != (128,) or b3.shape != (128,) or b4.shape != (128,):
            raise RuntimeError(f"GNN_Layer kernels need out_features = 128; got W4 {tuple(W4.shape)}")
        prec = PRECISIONS[_precision]
        y = _empty((N, H), x)
        pq = _empty((N, 2 * H), x)
        agg = _empty((N, H), x)
        y1_pre = _empty((N, H), x)
        y2_pre = _empty((N, H), x)
        rstd = _empty((max(seg.n_graphs, 1), H), x)
        _lib.same_device(x, u, pos, var, W1, plan.rowptr, seg.gptr)
        with _lib.on(x):
            ws = _lib.workspace(L.mgb_gnn_layer_fwd_workspace(N, plan.n_edges, seg.n_graphs, seg.max_nodes), x.device)
            _lib.check(L.mgb_gnn_layer_fwd(N, plan.n_edges, tw, dp, nv, seg.n_graphs, seg.max_nodes, _lib.ptr(x), _lib.ptr(u),
                                           _lib.ptr(pos), _lib.ptr(var), _lib.ptr(plan.rowptr), _lib.ptr(plan.dst),
                                           _lib.ptr(plan.src), _lib.ptr(seg.gptr), _lib.ptr(packed), _lib.ptr(b2),
                                           _lib.ptr(b3), _lib.ptr(b4), _lib.ptr(y), _lib.ptr(pq), _lib.ptr(agg),
                                           _lib.ptr(y1_pre), _lib.ptr(y2_pre), _lib.ptr(rstd), prec, _lib.ptr(ws), ws.numel(),
                                           _lib.stream()), "gnn_layer_fwd")
        ctx.save_for_backward(x, u, pos, var, y, pq, agg, y1_pre, y2_pre, rstd, packed, W1, W2, b2, W3, W4)
        ctx.plan, ctx.seg = plan, seg
        ctx.dims = (N, tw, dp, nv)
        ctx.prec = prec
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        x, u, pos, var, y, pq, agg, y1_pre, y2_pre, rstd, packed, W1, W2, b2, W3, W4 = ctx.saved_tensors
        plan, seg = ctx.plan, ctx.seg
        N, tw, dp, nv = ctx.dims
        dy = _lib.f32c(dy)
        need = ctx.needs_input_grad
        dx = _empty(x.shape, x)
        du = _empty(u.shape, x) if need[1] else None
        dpos = _empty(pos.shape, x) if need[2] else None
        dvar = _empty(var.shape, x) if need[3] else None
        dW1, db1 = _empty(W1.shape, x), _empty((128,), x)
        dW2, db2 = _empty((128, 128), x), _empty((128,), x)
        dW3, db3 = _empty(W3.shape, x), _empty((128,), x)
        dW4, db4 = _empty((128, 128), x), _empty((128,), x)
        with torch.cuda.device(x.device):
            ws = _lib.workspace(L.mgb_gnn_layer_bwd_workspace(N, plan.n_edges, tw, dp, nv, seg.n_graphs, seg.max_nodes),
                                x.device)
            _lib.check(L.mgb_gnn_layer_bwd(
                N, plan.n_edges, tw, dp, nv, seg.n_graphs, seg.max_nodes, _lib.ptr(dy), _lib.ptr(x), _lib.ptr(u),
                _lib.ptr(pos), _lib.ptr(var), _lib.ptr(y), _lib.ptr(pq), _lib.ptr(agg), _lib.ptr(y1_pre),
                _lib.ptr(y2_pre), _lib.ptr(rstd), _lib.ptr(plan.rowptr), _lib.ptr(plan.dst), _lib.ptr(plan.src),
                _lib.ptr(plan.rowptr_t), _lib.ptr(plan.pos_t), _lib.ptr(seg.gptr), _lib.ptr(packed), _lib.ptr(W2),
                _lib.ptr(b2), _lib.ptr(W3), _lib.ptr(W4), _lib.ptr(dx), _lib.ptr(du), _lib.ptr(dpos), _lib.ptr(dvar),
                _lib.ptr(dW1), _lib.ptr(db1), _lib.ptr(dW2), _lib.ptr(db2), _lib.ptr(dW3), _lib.ptr(db3),
                _lib.ptr(dW4), _lib.ptr(db4), 0, ctx.prec, _lib.ptr(ws), ws.numel(), _lib.stream()), "gnn_layer_bwd")
        return dx, du, dpos, dvar, dW1, db1, dW2, db2, dW3, db3, dW4, db4, None, None, None


# --------------------------------------------------------------------------------------------
# temporal-bundling decoder + Euler update (models/mpnn_2d.py:138-162,196-200) — csrc/decoder.cu
# --------------------------------------------------------------------------------------------
class BundlingDecoderFn(torch.autograd.Function):
    """out[n, j] = u[n, -1] + (j + 1) dt * Conv1d(8->1, k2)(act(Conv1d(1->8, k1, stride)(h[n, None, :])))[j] in one launch per
    direction; dt is a 0-dim device tensor (no host sync)."""

    @staticmethod
    def _args(h, u, w1, b1, w2, b2, dt, stride, act):
        k1, k2 = w1.shape[-1], w2.shape[-1]
        tw = (128 - k1) // stride + 1 - k2 + 1
        return (_lib.ptr(h), h.shape[0], h.shape[1], _lib.ptr(u), u.stride(0), u.shape[1] - 1, _lib.ptr(w1), _lib.ptr(b1), k1, stride,
                _lib.ptr(w2), _lib.ptr(b2), k2, tw, act, _lib.ptr(dt)), tw

    @staticmethod
    def forward(ctx, h, u, w1, b1, w2, b2, dt, stride: int, act: int):
        _lib.require_cuda(h, u, w1, w2, dt)
        L = _lib.lib()
        h, u = _lib.f32c(h), _lib.f32c(u)
        w1c, b1c, w2c, b2c = (_lib.f32c(v.detach()) for v in (w1, b1, w2, b2))
        dtc = _lib.f32c(dt.detach()).reshape(1)
        with torch.cuda.device(h.device):
            args, tw = BundlingDecoderFn._args(h, u, w1c, b1c, w2c, b2c, dtc, stride, act)
            out = _empty((h.shape[0], tw), h)
            _lib.check(L.mgb_bundling_decoder_fwd(*args, _lib.ptr(out), _lib.stream()), "bundling_decoder_fwd")
        ctx.save_for_backward(h, u, w1c, b1c, w2c, b2c, dtc)
        ctx.stride, ctx.act, ctx.shapes = stride, act, (w1.shape, w2.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib.lib()
        h, u, w1c, b1c, w2c, b2c, dtc = ctx.saved_tensors
        dout = _lib.f32c(dout)
        with torch.cuda.device(h.device):
            args, tw = BundlingDecoderFn._args(h, u, w1c, b1c, w2c, b2c, dtc, ctx.stride, ctx.act)
            dh = _empty(h.shape, h)
            du = torch.zeros_like(u) if ctx.needs_input_grad[1] else None
            dw1, db1, dw2, db2 = _empty(w1c.shape, h), _empty(b1c.shape, h), _empty(w2c.shape, h), _empty(b2c.shape, h)
            ws = _lib.workspace(L.mgb_bundling_decoder_bwd_workspace(h.shape[0]), h.device)
            _lib.check(L.mgb_bundling_decoder_bwd(*args, _lib.ptr(dout), _lib.ptr(dh), _lib.ptr(du), u.shape[1] if du is not None else 0,
                                                  _lib.ptr(dw1), _lib.ptr(db1), _lib.ptr(dw2), _lib.ptr(db2), _lib.ptr(ws), ws.numel(),
                                                  _lib.stream()), "bundling_decoder_bwd")
        return dh, du, dw1.reshape(ctx.shapes[0]), db1, dw2.reshape(ctx.shapes[1]), db2, None, None, None


def bundling_decoder(h, u, conv1, conv2, dt, swish: bool):
    """``u[:, -1:] + cumsum(dt) * output_mlp(h[:, None]).squeeze(1)`` (models/mpnn_2d.py:196-200) for the reference's decoder shapes
    (Conv1d(1, 8, k1 <= 16, stride) [Swish] Conv1d(8, 1, k2 <= 16) on 128 hidden features)."""
    ok = (h.is_cuda and h.shape[1] == 128 and conv1.in_channels == 1 and conv1.out_channels == 8 and conv2.in_channels == 8
          and conv2.out_channels == 1 and conv1.kernel_size[0] <= 16 and conv2.kernel_size[0] <= 16 and conv2.stride[0] == 1
          and conv1.padding[0] == 0 and conv2.padding[0] == 0 and conv1.dilation[0] == 1 and conv2.dilation[0] == 1
          and conv1.bias is not None and conv2.bias is not None)
    if not ok:
        raise RuntimeError("bundling_decoder kernels cover the reference's decoders: Conv1d(1,8,k1<=16,stride) [Swish] Conv1d(8,1,k2<=16) "
                           "on 128 hidden features")
    dt = torch.as_tensor(dt, dtype=torch.float32, device=h.device)
    return BundlingDecoderFn.apply(h, u, conv1.weight, conv1.bias, conv2.weight, conv2.bias, dt, conv1.stride[0], ACT["swish"] if swish else 0)


# --------------------------------------------------------------------------------------------
# row-wise Linear (+activation, +residual) and LayerNorm
# --------------------------------------------------------------------------------------------
@_lib.guard
def _linear_forward(x, W, b, act: int, residual, packed, want_pre: bool):
    """The forward kernels of LinearActFn; also called directly (no autograd node) when nothing requires a gradient."""
    _lib.require_cuda(x, W)
    L = _lib.lib()
    shape = x.shape
    x2 = _lib.f32c(x).reshape(-1, shape[-1])
    bc = _lib.f32c(b.detach())
    rows, fin = x2.shape
    fout = W.shape[0]
    y = _empty((rows, fout), x2)
    y_pre = _empty((rows, fout), x2) if (act != 0 and want_pre) else None    # inference skips the copy
    res = _lib.f32c(residual).reshape(rows, fout) if residual is not None else None
    if packed is not None:      # tensor-core path (fp16 hi/lo split, or bf16 in the bf16 mode): weight images from linear_act
        _lib.check(L.mgb_linear_tc_fwd(_lib.ptr(x2), rows, fin, fout, _lib.ptr(packed), _lib.ptr(bc), act, _lib.ptr(res),
                                       _lib.ptr(y), _lib.ptr(y_pre), 2 if _precision == "bf16" else _LINEAR_TC_PRECISION,
                                       _lib.stream()), "linear_tc_fwd")
    else:
        Wc = _lib.f32c(W.detach())
        wt = _empty((fin, fout), x2)
        _lib.check(L.mgb_transpose(_lib.ptr(Wc), fout, fin, _lib.ptr(wt), _lib.stream()), "transpose")
        _lib.check(L.mgb_linear_fwd(_lib.ptr(x2), rows, fin, fout, _lib.ptr(wt), _lib.ptr(bc), act, _lib.ptr(res),
                                    _lib.ptr(y), _lib.ptr(y_pre), _lib.stream()), "linear_fwd")
    return y.reshape(*shape[:-1], fout), x2, y_pre


class LinearActFn(torch.autograd.Function):
    """y = act(x W^T + b) (+ residual) — nn.Linear followed by Swish/ReLU (models/backbones/mlp.py:24-27)."""

    @staticmethod
    def forward(ctx, x, W, b, act: int, residual: Optional[torch.Tensor], packed=None, owner=None):
        y, x2, y_pre = _linear_forward(x, W, b, act, residual, packed, any(ctx.needs_input_grad))
        ctx.save_for_backward(x2, _lib.f32c(W.detach()), y_pre)
        ctx.act, ctx.shape, ctx.has_res = act, x.shape, residual is not None
        ctx.w_ref, ctx.owner = W, owner          # for the backward's tensor-core weight images (cached on the parameter)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        x2, Wc, y_pre = ctx.saved_tensors
        rows, fin = x2.shape
        fout = Wc.shape[0]
        dy2 = _lib.f32c(dy).reshape(rows, fout)
        dx = _empty((rows, fin), x2) if ctx.needs_input_grad[0] else None
        dW, db = _empty(Wc.shape, x2), _empty((fout,), x2)
        with torch.cuda.device(x2.device):
            tc_ws = L.mgb_linear_tc_bwd_workspace(rows, fin, fout) if (_linear_tc and _precision != "fp32" and rows >= 512) else 0
            if tc_ws:
                # tensor cores: bf16 hi/lo split of both operands (gradients span too many binades for the fp16 split)
                prec = 2 if _precision == "bf16" else 1
                packed = _tc_weight_images(ctx.w_ref, ctx.owner, prec)
                ws = _lib.workspace(tc_ws, x2.device)
                _lib.check(L.mgb_linear_tc_bwd(_lib.ptr(dy2), _lib.ptr(y_pre), ctx.act, _lib.ptr(x2), rows, fin, fout, _lib.ptr(packed),
                                               _lib.ptr(dx), _lib.ptr(dW), _lib.ptr(db), 0, prec, _lib.ptr(ws), ws.numel(), _lib.stream()),
                           "linear_tc_bwd")
            else:
                ws = _lib.workspace(L.mgb_linear_bwd_workspace(rows, fin, fout), x2.device)
                _lib.check(L.mgb_linear_bwd(_lib.ptr(dy2), _lib.ptr(y_pre), ctx.act, _lib.ptr(x2), rows, fin, fout,
                                            _lib.ptr(Wc), _lib.ptr(dx), _lib.ptr(dW), _lib.ptr(db), 0, _lib.ptr(ws),
                                            ws.numel(), _lib.stream()), "linear_bwd")
        dres = dy if ctx.has_res else None
        return (dx.reshape(ctx.shape) if dx is not None else None), dW, db, None, dres, None, None


# The 128-wide Linears of MLP / Encoder / Decoder / projector (models/backbones/mlp.py) run on the tensor cores
# (mgb_linear_tc_fwd) with both operands split into two fp16 values (22 significant bits, small terms accumulated first):
# measured 2.6e-7 per Linear against fp64 (the fp32 FFMA GEMM: 2.9e-7) and 3.9e-6 on MAgNet's ill-conditioned final
# `hr_points` (FFMA: 3.0e-6) — inside the 1e-5 contract, 2.3x faster decode and rollout.  fp16 limits the inputs to
# |x| < 32768 (raised as an error by the kernel) and is tuned for O(1) activations, which is what these MLPs see behind
# their LayerNorms; set_linear_tc(False) restores the exact fp32 GEMM.  The backward pass of these Linears is fp32 FFMA.
_linear_tc = True
_LINEAR_TC_PRECISION = 3      # fp16 hi/lo split (include/magnet_b200.h); the per-edge kernels keep the bf16 split


def check_fp16_range(sync: bool = False) -> None:
    """Raise if a fp16-split kernel (the default arithmetic of the MAgNet MLPs) has met |x| >= 32768 since the last check.
    ``sync=True`` synchronises the current stream first, so the answer covers every call issued so far; without it the check
    costs one host read and reports what has already executed.  The models call it at the end of ``forward``."""
    if sync:
        torch.cuda.current_stream().synchronize()
    if _lib.lib().mgb_f16_range_check():
        raise RuntimeError("magnet_b200: a fp16-split tensor-core kernel met |x| >= 32768 (outside the fp16 range); results of this "
                           "forward are not valid — use functional.set_linear_tc(False) (fp32 FFMA Linears) for this data")


def set_linear_tc(on: bool) -> bool:
    """Route eligible nn.Linear forwards through the tcgen05 kernel; returns the previous setting."""
    global _linear_tc
    old, _linear_tc = _linear_tc, bool(on)
    return old


def _tc_weight_images(W: torch.Tensor, owner: Optional[torch.Tensor] = None, prec: Optional[int] = None):
    """Swizzled 16-bit (hi | lo) images of W for the tensor-core Linear, or None when the shape is not covered.  Cached on
    the owning parameter object (``owner``; W may be a column slice of it — a slice taken under torch.inference_mode()
    has neither ``_base`` nor a version counter, so callers that slice pass the parameter explicitly), keyed on the slice
    and the parameter's version counter: an optimizer step or load_state_dict rebuilds them, a rollout reuses them."""
    if not _linear_tc or _precision == "fp32" or W.dim() != 2 or W.stride(1) != 1 or W.dtype != torch.float32 or not W.is_cuda:
        return None
    L = _lib.lib()
    fout, fin = W.shape
    n = L.mgb_linear_tc_packed_floats(fin, fout)
    if n == 0:
        return None
    if owner is None:
        owner = W._base if W._base is not None else W
    cache = owner.__dict__.setdefault("_mgb_tc_images", {})
    if prec is None:
        prec = 2 if _precision == "bf16" else _LINEAR_TC_PRECISION
    key = (W.storage_offset(), fout, fin, W.stride(0), owner.data_ptr(), prec, torch.cuda.current_stream().cuda_stream)
    hit = cache.get(key)
    version = _lib.ver(owner)
    if hit is not None and hit[0] == version:
        return hit[1]
    packed = torch.empty(n, dtype=torch.float32, device=W.device)
    Wd = W.detach()
    with _lib.on(W):
        _lib.check(L.mgb_linear_tc_pack(_lib.ptr(Wd), Wd.stride(0), fin, fout, prec, _lib.ptr(packed), _lib.stream()), "linear_tc_pack")
    cache[key] = (version, packed)
    return packed


@_lib.guard
def _chain_packed(linears, cache_owner, device):
    """fp16 hi | lo weight images + biases of a 128-wide Linear chain (mgb_mlp_chain_pack_layer), cached on ``cache_owner``
    keyed on the version counters of the parameters."""
    L = _lib.lib()
    nl = len(linears)
    key = tuple((l.weight.data_ptr(), _lib.ver(l.weight), l.bias.data_ptr(), _lib.ver(l.bias)) for l in linears) + \
        (torch.cuda.current_stream().cuda_stream,)
    store = cache_owner.__dict__ if cache_owner is not None else None
    hit = store.get("_mgb_chain") if store is not None else None
    if hit is not None and hit[0] == key:
        return hit[1]
    packed = torch.empty(L.mgb_mlp_chain_packed_floats(nl), dtype=torch.float32, device=device)
    for i, l in enumerate(linears):
        W, b = l.weight.detach(), l.bias.detach()
        if W.stride(1) != 1:
            W = W.contiguous()
        _lib.check(L.mgb_mlp_chain_pack_layer(_lib.ptr(W), W.stride(0), l.out_features, _lib.ptr(b.contiguous()), i, nl,
                                              _lib.ptr(packed), _lib.stream()), "mlp_chain_pack_layer")
    if store is not None:
        store["_mgb_chain"] = (key, packed)
    return packed


def _chain_ok(linears) -> bool:
    return (1 <= len(linears) <= 8 and all(l.in_features == 128 for l in linears) and all(l.out_features == 128 for l in linears[:-1])
            and linears[-1].out_features <= 128)


def inr_decode_fusable(x_lr, lr_encoded, wp, linears) -> bool:
    """The fused decoder covers the reference configuration (latent_dim = n_chan = mlp_hidden = 128, ReLU projector) on the
    fp16-split tensor-core arithmetic, forward only."""
    return (_linear_tc and _precision == "fp32_tc" and x_lr.is_cuda and lr_encoded.shape[-1] == 128 and wp.shape[0] == 128
            and _chain_ok(linears) and _no_grad_needed(x_lr, lr_encoded, wp, *[p for l in linears for p in (l.weight, l.bias)]))


_GRID_CACHE = {}


@_lib.guard
def inr_decode_fused(xlr, lr_encoded, lr_coords, hr_coords, t, wp, bp, linears, B, L_, nq, k, interpolation, cache_owner=None, idx=None,
                     grid_owner=None):
    """projector(continuous_decoder(...)) (models/magnet_gnn.py:338-339) in one launch plus the per-node Linear for the latent
    part of proj_head: xlr [B,T,L], lr_encoded [B*L,128], lr_coords [B*L,d], hr_coords [B*nq,d], t [B,>=T] -> [B*nq, T, n_out].
    ``idx`` None: the nearest low-res nodes are searched inside the kernel.  ``grid_owner``: the caller's low-res coordinate
    tensor (``lr_coords`` is usually a fresh view of it): its identity and version key the cached search grid."""
    from . import graph as MG
    Lb = _lib.lib()
    T, d = xlr.shape[1], lr_coords.shape[1]
    owner = grid_owner if grid_owner is not None else lr_coords
    xlr, lr_coords, hr_coords, t = (_lib.f32c(v) for v in (xlr, lr_coords, hr_coords, t))
    a = linear_act(lr_encoded, wp[:, :128], bp, "none", owner=wp)
    wpc = _lib.f32c(wp.detach())
    packed = _chain_packed(linears, cache_owner, xlr.device)
    Q = hr_coords.shape[0]
    n_out = linears[-1].out_features
    y = _empty((Q * T, n_out), a)
    ptr_x = MG.uniform_ptr(B, L_, xlr.device) if idx is None else None
    # the grid hash of the low-res mesh lives in the workspace and is reused while the mesh tensor is unchanged (a rollout
    # decodes every step over the same mesh: SURVEY F10 applied to the decoder's search structure)
    ready = 0
    if idx is None:
        key = (lr_coords.data_ptr(), owner.data_ptr(), _lib.ver(owner), tuple(lr_coords.shape), B, L_, torch.cuda.current_stream().cuda_stream)
        hit = _GRID_CACHE.get(key)
        if hit is not None and hit[0]() is owner:
            ws, ready = hit[1], 1
        else:
            import weakref
            ws = _lib.workspace(Lb.mgb_inr_decode_fused_workspace(B * L_, B), xlr.device)
            if len(_GRID_CACHE) >= 4:
                _GRID_CACHE.clear()
            _GRID_CACHE[key] = (weakref.ref(owner), ws)
    else:
        ws = _lib.workspace(0, xlr.device)
    _lib.check(Lb.mgb_inr_decode_fused(_lib.ptr(a), _lib.ptr(xlr), _lib.ptr(lr_coords), _lib.ptr(hr_coords), _lib.ptr(t), t.shape[1],
                                       ctypes_offset(wpc, 128), wpc.shape[1], _lib.ptr(idx), k, _lib.ptr(ptr_x), B, Q, nq, L_, T, d,
                                       INTERP[interpolation], len(linears), _lib.ptr(packed), n_out, _lib.ptr(y), ready, _lib.ptr(ws),
                                       ws.numel(), _lib.stream()), "inr_decode_fused")
    return y.reshape(Q, T, n_out)


def mlp_chain(x, linears, act: str = "relu", in_act: str = "none", cache_owner=None):
    """Inference forward of ``linears`` (nn.Linear modules: Linear(128,128) + act ... Linear(128, out <= 128)) in one launch
    (mgb_mlp_chain_fwd): the activations never leave the SM between the layers.  Returns None when the shapes or the
    current settings do not allow it (the caller then runs the layers one by one).  The packed weights are cached on
    ``cache_owner`` (the MLP module), keyed on the version counters of its parameters."""
    if not _linear_tc or _precision != "fp32_tc" or torch.is_grad_enabled() and any(p.requires_grad for l in linears for p in (l.weight, l.bias)):
        return None
    if x.shape[-1] != 128 or not x.is_cuda or x.dtype != torch.float32 or not (1 <= len(linears) <= 8):
        return None
    if any(l.in_features != 128 for l in linears) or any(l.out_features != 128 for l in linears[:-1]) or linears[-1].out_features > 128:
        return None
    L = _lib.lib()
    nl = len(linears)
    packed = _chain_packed(linears, cache_owner, x.device)
    shape = x.shape
    x2 = _lib.f32c(x).reshape(-1, 128)
    n_out = linears[-1].out_features
    y = _empty((x2.shape[0], n_out), x2)
    _lib.check(L.mgb_mlp_chain_fwd(_lib.ptr(x2), 128, x2.shape[0], nl, _lib.ptr(packed), ACT[act], ACT[in_act], n_out, _lib.ptr(y),
                                   n_out, _lib.stream()), "mlp_chain_fwd")
    return y.reshape(*shape[:-1], n_out)


def linear_act(x, W, b, act: str = "none", residual=None, owner=None):
    """``owner``: the nn.Parameter ``W`` is a column slice of (keeps the weight-image cache valid under inference_mode)."""
    packed = _tc_weight_images(W, owner)
    if not torch.is_grad_enabled() or not (x.requires_grad or W.requires_grad or b.requires_grad or
                                           (residual is not None and residual.requires_grad)):
        return _linear_forward(x, W, b, ACT[act], residual, packed, False)[0]      # rollout / decode: no autograd node
    return LinearActFn.apply(x, W, b, ACT[act], residual, packed, owner)


@_lib.guard
def magnet_features(u_, x_, t_last, edge_index):
    """(node_features, edge_features) of MAgNetGNN._build_graph (models/magnet_gnn.py:298-308) in one launch (inference):
    u_ [N,C], x_ [N,d], t_last [B] (tiled over the rows: quirk F7), edge_index int64 [2,E]."""
    L = _lib.lib()
    u_, x_, t_last = _lib.f32c(u_), _lib.f32c(x_), _lib.f32c(t_last)
    ei = edge_index.contiguous()
    N, C, d, E = u_.shape[0], u_.shape[1], x_.shape[1], ei.shape[1]
    nf, ef = _empty((N, C + d + 1), u_), _empty((E, C + d), u_)
    _lib.check(L.mgb_magnet_features(_lib.ptr(u_), C, _lib.ptr(x_), d, _lib.ptr(t_last), t_last.numel(), N, _lib.ptr(ei), E,
                                     _lib.ptr(nf), _lib.ptr(ef), _lib.stream()), "magnet_features")
    return nf, ef


@_lib.guard
def linear_act2(x0, x1, W, b, act: str = "none"):
    """act(cat([x0, x1], -1) W^T + b) for two [rows, 128] tensors without the concatenation (inference; returns None when the
    tensor-core path does not cover the shape and the caller should concatenate)."""
    packed = _tc_weight_images(W)
    if (packed is None or x0.shape != x1.shape or x0.shape[-1] != 128 or W.shape[1] != 256 or W.shape[0] > 256
            or not _no_grad_needed(x0, x1, W, b)):
        return None
    L = _lib.lib()
    a0, a1 = _lib.f32c(x0).reshape(-1, 128), _lib.f32c(x1).reshape(-1, 128)
    rows, fout = a0.shape[0], W.shape[0]
    y = _empty((rows, fout), a0)
    _lib.check(L.mgb_linear_tc_fwd2(_lib.ptr(a0), 128, _lib.ptr(a1), 128, rows, fout, _lib.ptr(packed), _lib.ptr(_lib.f32c(b.detach())),
                                    ACT[act], _lib.ptr(y), 2 if _precision == "bf16" else _LINEAR_TC_PRECISION, _lib.stream()),
               "linear_tc_fwd2")
    return y.reshape(*x0.shape[:-1], fout)


@_lib.guard
def layer_norm_residual(x, gamma, beta, residual):
    """LayerNorm(x) + residual in one pass (inference)."""
    if not _no_grad_needed(x, gamma, beta, residual) or x.shape[-1] != 128:
        return layer_norm(x, gamma, beta) + residual
    L = _lib.lib()
    x2, r2 = _lib.f32c(x).reshape(-1, 128), _lib.f32c(residual).reshape(-1, 128)
    y = _empty(x2.shape, x2)
    _lib.check(L.mgb_layernorm_residual_fwd(_lib.ptr(x2), _lib.ptr(_lib.f32c(gamma.detach())), _lib.ptr(_lib.f32c(beta.detach())), _lib.ptr(r2),
                                            x2.shape[0], 128, _lib.ptr(y), _lib.stream()), "layernorm_residual_fwd")
    return y.reshape(x.shape)


@_lib.guard
def _layernorm_forward(x, gamma, beta):
    _lib.require_cuda(x, gamma)
    L = _lib.lib()
    shape = x.shape
    x2 = _lib.f32c(x).reshape(-1, shape[-1])
    g, b = _lib.f32c(gamma.detach()), _lib.f32c(beta.detach())
    rows, cols = x2.shape
    y = _empty((rows, cols), x2)
    stats = _empty((max(rows, 1), 2), x2)
    _lib.check(L.mgb_layernorm_fwd(_lib.ptr(x2), _lib.ptr(g), _lib.ptr(b), rows, cols, _lib.ptr(y), _lib.ptr(stats),
                                   _lib.stream()), "layernorm_fwd")
    return y.reshape(shape), x2, g, stats


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm(128): eps 1e-5, biased variance, affine (models/magnet_gnn.py:27,35,60,67)."""

    @staticmethod
    def forward(ctx, x, gamma, beta):
        y, x2, g, stats = _layernorm_forward(x, gamma, beta)
        ctx.save_for_backward(x2, g, stats)
        ctx.shape = x.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        x2, g, stats = ctx.saved_tensors
        rows, cols = x2.shape
        dy2 = _lib.f32c(dy).reshape(rows, cols)
        dx = _empty((rows, cols), x2)
        dg, db = _empty((cols,), x2), _empty((cols,), x2)
        with torch.cuda.device(x2.device):
            ws = _lib.workspace(L.mgb_layernorm_bwd_workspace(rows, cols), x2.device)
            _lib.check(L.mgb_layernorm_bwd(_lib.ptr(dy2), _lib.ptr(x2), _lib.ptr(g), _lib.ptr(stats), rows, cols,
                                           _lib.ptr(dx), _lib.ptr(dg), _lib.ptr(db), 0, _lib.ptr(ws), ws.numel(),
                                           _lib.stream()), "layernorm_bwd")
        return dx.reshape(ctx.shape), dg, db


def _no_grad_needed(*tensors) -> bool:
    return not torch.is_grad_enabled() or not any(t is not None and t.requires_grad for t in tensors)


def layer_norm(x, gamma, beta):
    if _no_grad_needed(x, gamma, beta):
        return _layernorm_forward(x, gamma, beta)[0]         # rollout / decode: no autograd node
    return LayerNormFn.apply(x, gamma, beta)


@_lib.guard
def instance_norm(x: torch.Tensor, seg: GraphSegments) -> torch.Tensor:
    """Forward-only PyG InstanceNorm (tests / inference helper)."""
    L = _lib.lib()
    x = _lib.f32c(x)
    y = torch.empty_like(x)
    rstd = _empty((max(seg.n_graphs, 1), x.shape[1]), x)
    ws = _lib.workspace(L.mgb_instance_norm_workspace(seg.n_graphs, seg.max_nodes), x.device)
    _lib.check(L.mgb_instance_norm_fwd(_lib.ptr(x), _lib.ptr(seg.gptr), seg.n_graphs, seg.max_nodes, _lib.ptr(y),
                                       _lib.ptr(rstd), _lib.ptr(ws), ws.numel(), _lib.stream()), "instance_norm_fwd")
    return y


# --------------------------------------------------------------------------------------------
# InteractionNetwork glue (models/magnet_gnn.py:70-90) and the INR decoder (:224-283)
# --------------------------------------------------------------------------------------------
@_lib.guard
def _segment_sum(rows, cols, rowptr, idx, n_nodes, mean, out=None, ld_out=None):
    L = _lib.lib()
    if out is None:
        out = _empty((n_nodes, cols), rows)
        ld_out = cols
    _lib.check(L.mgb_segment_sum_rows(_lib.ptr(rows), cols, _lib.ptr(rowptr), _lib.ptr(idx), n_nodes, int(mean),
                                      _lib.ptr(out), ld_out, _lib.stream()), "segment_sum_rows")
    return out


class EdgeCombineFn(torch.autograd.Function):
    """out[e] = act(p[edge_index[1][e]] + q[edge_index[0][e]] + r[e]) — the factorised first Linear of edge_fn
    applied to cat([x_i, x_j, e_features]) (models/magnet_gnn.py:79-82)."""

    @staticmethod
    @_lib.guard
    def run(p, q, r, edge_index, act: int):
        _lib.require_cuda(p, q, r, edge_index)
        L = _lib.lib()
        p, q, r = _lib.f32c(p), _lib.f32c(q), _lib.f32c(r)
        ei = edge_index.contiguous()
        E = ei.shape[1]
        if p.shape[1] != 128 or q.shape != p.shape or r.shape != (E, 128):
            raise RuntimeError("edge_combine kernels are built for hidden width 128: got p "
                               f"{tuple(p.shape)}, q {tuple(q.shape)}, r {tuple(r.shape)} for {E} edges")
        out = _empty((E, 128), p)
        _lib.check(L.mgb_edge_combine_fwd(_lib.ptr(p), _lib.ptr(q), _lib.ptr(r), _lib.ptr(ei), E, act, _lib.ptr(out),
                                          _lib.stream()), "edge_combine_fwd")
        return out

    @staticmethod
    def forward(ctx, p, q, r, edge_index, plan: AggregationPlan, act: int):
        out = EdgeCombineFn.run(p, q, r, edge_index, act)
        ctx.save_for_backward(out)
        ctx.plan, ctx.act, ctx.n = plan, act, p.shape[0]
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib.lib()
        (out,) = ctx.saved_tensors
        plan = ctx.plan
        dout = _lib.f32c(dout)
        with torch.cuda.device(out.device):
            if ctx.act:
                dz = torch.empty_like(dout)
                _lib.check(L.mgb_relu_mask(_lib.ptr(dout), _lib.ptr(out), dout.numel(), _lib.ptr(dz), _lib.stream()), "relu_mask")
            else:
                dz = dout
            dp = _segment_sum(dz, 128, plan.rowptr, plan.perm, ctx.n, False)
            dq = _segment_sum(dz, 128, plan.rowptr_t, plan.perm_src(), ctx.n, False)
        return dp, dq, dz, None, None, None


def edge_combine(p, q, r, edge_index, plan, act: str = "none"):
    if _no_grad_needed(p, q, r):
        return EdgeCombineFn.run(p, q, r, edge_index, ACT[act])
    return EdgeCombineFn.apply(p, q, r, edge_index, plan, ACT[act])


class ScatterMeanFn(torch.autograd.Function):
    """scatter(m, edge_index[1], reduce='mean') through the destination-sorted plan (aggr='mean', models/magnet_gnn.py:54)."""

    @staticmethod
    def forward(ctx, m, edge_index, plan: AggregationPlan):
        m = _lib.f32c(m)
        ctx.plan = plan
        ctx.save_for_backward(edge_index)
        return _segment_sum(m, 128, plan.rowptr, plan.perm, plan.n_nodes, True)

    @staticmethod
    def backward(ctx, dagg):
        L = _lib.lib()
        (edge_index,) = ctx.saved_tensors
        plan = ctx.plan
        dagg = _lib.f32c(dagg)
        E = plan.n_edges
        dm = _empty((E, 128), dagg)
        with torch.cuda.device(dagg.device):
            ei1 = edge_index[1].contiguous()
            _lib.check(L.mgb_gather_rows(_lib.ptr(dagg), _lib.ptr(ei1), _lib.ptr(plan.rowptr), E, _lib.ptr(dm), _lib.stream()),
                       "gather_rows")
        return dm, None, None


def scatter_mean(m, edge_index, plan):
    if m.shape[-1] != 128 or m.shape[0] != plan.n_edges:
        raise RuntimeError(f"scatter_mean kernels are built for [E, 128] messages; got {tuple(m.shape)} for {plan.n_edges} edges")
    if _no_grad_needed(m):
        return _segment_sum(_lib.f32c(m), 128, plan.rowptr, plan.perm, plan.n_nodes, True)
    return ScatterMeanFn.apply(m, edge_index, plan)


# --------------------------------------------------------------------------------------------
# fused InteractionNetwork edge function (csrc/in_edge_tc.cu): gather + edge_fn MLP + LayerNorm + mean in one launch
# --------------------------------------------------------------------------------------------
def params_key(params):
    """Cache key for kernel-side copies of parameters: storage identity + version counters + arithmetic + stream."""
    return tuple((p.data_ptr(), _lib.ver(p)) for p in params) + (_precision, _linear_tc, torch.cuda.current_stream().cuda_stream)


def in_edge_fusable(x, e_features, linears) -> bool:
    """The fused kernel covers the reference configuration (hidden 128, mlp_layers 4) on the tensor-core arithmetic,
    forward only (rollout / decode: nothing requires a gradient)."""
    return (_linear_tc and _precision != "fp32" and x.is_cuda and x.dtype == torch.float32 and e_features.dtype == torch.float32
            and x.dim() == 2 and x.shape[1] == 128 and e_features.dim() == 2 and e_features.shape[1] == 128 and len(linears) == 5
            and linears[0].in_features == 384 and all(l.out_features == 128 for l in linears)
            and all(l.in_features == 128 for l in linears[1:])
            and _no_grad_needed(x, e_features, *[p for l in linears for p in (l.weight, l.bias)]))


@_lib.guard
def in_edge_pack(x_like, linears, norm):
    """Kernel-side copies for in_edge_fused: (pq weight images, pq bias, packed edge_fn images + biases + LayerNorm affine)."""
    L = _lib.lib()
    dev = x_like.device
    lin_prec = 2 if _precision == "bf16" else _LINEAR_TC_PRECISION
    W0 = linears[0].weight.detach()
    if W0.stride(1) != 1:
        W0 = W0.contiguous()
    b0 = linears[0].bias.detach()
    w_pq = torch.cat([W0[:, :128], W0[:, 128:256]], dim=0).contiguous()                  # [256, 128]: rows of P, then of Q
    b_pq = torch.cat([b0, torch.zeros_like(b0)]).contiguous()
    pq_img = torch.empty(L.mgb_linear_tc_packed_floats(128, 256), dtype=torch.float32, device=dev)
    _lib.check(L.mgb_linear_tc_pack(_lib.ptr(w_pq), 128, 128, 256, lin_prec, _lib.ptr(pq_img), _lib.stream()), "linear_tc_pack")
    packed = torch.empty(L.mgb_in_edge_packed_floats(), dtype=torch.float32, device=dev)
    eprec = 2 if _precision == "bf16" else 3
    _lib.check(L.mgb_in_edge_pack_layer(ctypes_offset(W0, 256), W0.stride(0), None, 0, eprec, _lib.ptr(packed), _lib.stream()),
               "in_edge_pack_layer")
    for i, l in enumerate(linears[1:], start=1):
        W = l.weight.detach().contiguous()
        b = l.bias.detach().contiguous()
        _lib.check(L.mgb_in_edge_pack_layer(_lib.ptr(W), 128, _lib.ptr(b), i, eprec, _lib.ptr(packed), _lib.stream()),
                   "in_edge_pack_layer")
    g, bt = norm.weight.detach().contiguous(), norm.bias.detach().contiguous()
    _lib.check(L.mgb_in_edge_pack_norm(_lib.ptr(g), _lib.ptr(bt), _lib.ptr(packed), _lib.stream()), "in_edge_pack_norm")
    return pq_img, b_pq, packed


@_lib.guard
def in_edge_fused(x, e_features, e_scale: float, plan: AggregationPlan, packs):
    """agg [N,128] = mean over edge_index[1] of LayerNorm(edge_fn(cat[x_i, x_j, e_scale * e_features])) — one node-level
    Linear for P | Q plus ONE fused edge launch (mgb_in_edge_fwd); no [E,128] intermediate exists."""
    _lib.require_cuda(x, e_features)
    L = _lib.lib()
    pq_img, b_pq, packed = packs
    x = _lib.f32c(x)
    e = _lib.f32c(e_features)
    N, E = x.shape[0], plan.n_edges
    if plan.n_nodes != N or e.shape[0] != E:
        raise RuntimeError(f"in_edge_fused: plan is for {plan.n_nodes} nodes / {plan.n_edges} edges, got x {tuple(x.shape)}, "
                           f"e_features {tuple(e.shape)}")
    lin_prec = 2 if _precision == "bf16" else _LINEAR_TC_PRECISION
    pq = _empty((N, 256), x)
    _lib.check(L.mgb_linear_tc_fwd(_lib.ptr(x), N, 128, 256, _lib.ptr(pq_img), _lib.ptr(b_pq), 0, None, _lib.ptr(pq), None,
                                   lin_prec, _lib.stream()), "linear_tc_fwd")
    agg = _empty((N, 128), x)
    ws = _lib.workspace(L.mgb_in_edge_fwd_workspace(E), x.device)
    _lib.check(L.mgb_in_edge_fwd(_lib.ptr(e), float(e_scale), _lib.ptr(plan.perm), _lib.ptr(pq), _lib.ptr(plan.rowptr),
                                 _lib.ptr(plan.dst), _lib.ptr(plan.src), N, E, _lib.ptr(packed), 2 if _precision == "bf16" else 3,
                                 _lib.ptr(agg), _lib.ptr(ws), ws.numel(), _lib.stream()), "in_edge_fwd")
    return agg


_fused_training = True


def set_fused_training(on: bool) -> bool:
    """Route MAgNet's InteractionNetwork training step through the fused edge kernels (forward + recompute backward); off: the
    row-wise kernels with saved [E,128] activations.  Returns the previous setting."""
    global _fused_training
    old, _fused_training = _fused_training, bool(on)
    return old


def fused_training() -> bool:
    return _fused_training


def in_edge_trainable(x, e_features, linears) -> bool:
    """Training: the same fused forward launch plus the recompute-based fused backward (csrc/in_edge_bwd_tc.cu)."""
    return (_linear_tc and _precision != "fp32" and x.is_cuda and x.dtype == torch.float32 and e_features.dtype == torch.float32
            and x.dim() == 2 and x.shape[1] == 128 and e_features.dim() == 2 and e_features.shape[1] == 128 and len(linears) == 5
            and linears[0].in_features == 384 and all(l.out_features == 128 for l in linears)
            and all(l.in_features == 128 for l in linears[1:]))


class InEdgeFn(torch.autograd.Function):
    """agg = mean_{e -> i} LayerNorm(edge_fn(cat[x_i, x_j, e_scale * e])) given pq = [P | Q] (the node part of the first Linear,
    computed by the caller through autograd).  Forward: mgb_in_edge_fwd (one launch, nothing of size [E,128] saved);
    backward: mgb_in_edge_bwd (two recompute passes on the tensor cores) + the by-source sum for dQ + one tensor-core Linear
    backward for d e_features and dWe."""

    @staticmethod
    def forward(ctx, pq, e, W0, W1, b1, W2, b2, W3, b3, W4, b4, gamma, beta, e_scale, plan, packed):
        L = _lib.lib()
        pq, e = _lib.f32c(pq), _lib.f32c(e)
        N, E = pq.shape[0], plan.n_edges
        if plan.n_nodes != N or e.shape[0] != E:
            raise RuntimeError(f"InEdgeFn: plan is for {plan.n_nodes} nodes / {plan.n_edges} edges, got pq {tuple(pq.shape)}, "
                               f"e_features {tuple(e.shape)}")
        prec = 2 if _precision == "bf16" else 3
        agg = _empty((N, 128), pq)
        with torch.cuda.device(pq.device):
            ws = _lib.workspace(L.mgb_in_edge_fwd_workspace(E), pq.device)
            _lib.check(L.mgb_in_edge_fwd(_lib.ptr(e), float(e_scale), _lib.ptr(plan.perm), _lib.ptr(pq), _lib.ptr(plan.rowptr),
                                         _lib.ptr(plan.dst), _lib.ptr(plan.src), N, E, _lib.ptr(packed), prec, _lib.ptr(agg),
                                         _lib.ptr(ws), ws.numel(), _lib.stream()), "in_edge_fwd")
        ctx.save_for_backward(pq, e, packed, W0)
        ctx.plan, ctx.e_scale, ctx.prec = plan, float(e_scale), prec
        return agg

    @staticmethod
    def backward(ctx, dagg):
        L = _lib.lib()
        pq, e, packed, W0 = ctx.saved_tensors
        plan, e_scale = ctx.plan, ctx.e_scale
        N, E = pq.shape[0], plan.n_edges
        dagg = _lib.f32c(dagg)
        with torch.cuda.device(pq.device):
            dpq = _empty((N, 256), pq)
            dz0 = _empty((max(E, 1), 128), pq)
            dW, db = _empty((4, 128, 128), pq), _empty((4, 128), pq)
            dgamma, dbeta = _empty((128,), pq), _empty((128,), pq)
            ws = _lib.workspace(L.mgb_in_edge_bwd_workspace(E), pq.device)
            _lib.check(L.mgb_in_edge_bwd(_lib.ptr(dagg), _lib.ptr(e), e_scale, _lib.ptr(plan.perm), _lib.ptr(pq), _lib.ptr(plan.rowptr),
                                         _lib.ptr(plan.dst), _lib.ptr(plan.src), N, E, _lib.ptr(packed), ctx.prec, _lib.ptr(dpq),
                                         _lib.ptr(dz0), _lib.ptr(dW), _lib.ptr(db), _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.ptr(ws),
                                         ws.numel(), _lib.stream()), "in_edge_bwd")
            del ws
            de = dW0 = None
            if E > 0:
                # dQ[j] = sum of the dz0 rows of the edges leaving j (COO rows grouped by source: the transposed plan)
                dq = _segment_sum(dz0, 128, plan.rowptr_t, plan.perm_src(), N, False)
                dpq[:, 128:] = dq * (1.0 / e_scale) if e_scale != 1.0 else dq
                # d e_features = dz0 We and dWe = dz0^T e_features: one tensor-core Linear backward (x = e, W = We, dy = dz0)
                prec = 2 if ctx.prec == 2 else 1
                We = W0.detach()[:, 256:]
                img = _tc_weight_images(We, W0, prec)
                need_de = ctx.needs_input_grad[1]
                de = _empty((E, 128), pq) if need_de else None
                dWe, dbe = _empty((128, 128), pq), _empty((128,), pq)
                lws = _lib.workspace(L.mgb_linear_tc_bwd_workspace(E, 128, 128), pq.device)
                _lib.check(L.mgb_linear_tc_bwd(_lib.ptr(dz0), None, 0, _lib.ptr(e), E, 128, 128, _lib.ptr(img), _lib.ptr(de),
                                               _lib.ptr(dWe), _lib.ptr(dbe), 0, prec, _lib.ptr(lws), lws.numel(), _lib.stream()),
                           "linear_tc_bwd")
                dW0 = torch.zeros_like(W0)
                dW0[:, 256:] = dWe
        return (dpq, de, dW0, dW[0], db[0], dW[1], db[1], dW[2], db[2], dW[3], db[3], dgamma, dbeta, None, None, None)


INTERP = {"area": 0, "knn": 1, "sph": 2}


class InrBlendFn(torch.autograd.Function):
    """z[q, i, :] = blend of the two nearest low-res nodes' proj_head outputs (models/magnet_gnn.py:254-279) given
    a = lr_encoded Wp[:, :C]^T + b (per node).  Differentiable in a, xlr and proj_head.weight[:, C:]."""

    @staticmethod
    def forward(ctx, a, xlr, wp, lr_coords, hr_coords, t, idx, geom):
        nq, L_, T, d, mode = geom
        _lib.require_cuda(a, xlr, wp, lr_coords, hr_coords, t, idx)
        L = _lib.lib()
        a, xlr, lr_coords, hr_coords, t = (_lib.f32c(v) for v in (a, xlr, lr_coords, hr_coords, t))
        wpc = _lib.f32c(wp.detach())
        Q, k = idx.shape
        if a.shape[1] != 128 or wpc.shape != (128, 128 + d + 2):
            raise RuntimeError("inr_decode kernels are built for latent_dim = n_chan = 128: got a "
                               f"{tuple(a.shape)}, proj_head.weight {tuple(wpc.shape)} (d = {d})")
        z = _empty((Q, T, 128), a)
        ldw = wpc.shape[1]
        wsmall = ctypes_offset(wpc, 128)
        _lib.check(L.mgb_inr_decode_fwd(_lib.ptr(a), _lib.ptr(xlr), _lib.ptr(lr_coords), _lib.ptr(hr_coords), _lib.ptr(t),
                                        t.shape[1], wsmall, ldw, _lib.ptr(idx), k, Q, nq, L_, T, d, mode, _lib.ptr(z),
                                        _lib.stream()), "inr_decode_fwd")
        ctx.save_for_backward(a, xlr, wpc, lr_coords, hr_coords, t, idx)
        ctx.geom = geom
        return z

    @staticmethod
    def backward(ctx, dz):
        from .graph import build_plan
        L = _lib.lib()
        a, xlr, wpc, lr_coords, hr_coords, t, idx = ctx.saved_tensors
        nq, L_, T, d, mode = ctx.geom
        dz = _lib.f32c(dz)
        Q, k = idx.shape
        n_lr = a.shape[0]
        with torch.cuda.device(a.device):
            g = _empty((2 * Q, 128), a)
            sx = _empty((2 * Q, T), a)
            dwp = torch.zeros_like(wpc)
            ws = _lib.workspace(L.mgb_inr_decode_bwd_workspace(Q), a.device)
            _lib.check(L.mgb_inr_decode_bwd(_lib.ptr(a), _lib.ptr(xlr), _lib.ptr(lr_coords), _lib.ptr(hr_coords), _lib.ptr(t),
                                            t.shape[1], ctypes_offset(wpc, 128), wpc.shape[1], _lib.ptr(idx), k, Q, nq, L_, T, d,
                                            mode, _lib.ptr(dz), _lib.ptr(g), _lib.ptr(sx), ctypes_offset(dwp, 128), 0,
                                            _lib.ptr(ws), ws.numel(), _lib.stream()), "inr_decode_bwd")
            # per-low-res-node sums of the two per-query contribution rows (deterministic, via a sorted plan)
            sel = idx[:, :2].reshape(-1).contiguous()
            plan = build_plan(torch.stack([sel, sel]), n_lr)
            da = _segment_sum(g, 128, plan.rowptr, plan.perm, n_lr, False)
            dx_flat = _segment_sum(sx, T, plan.rowptr, plan.perm, n_lr, False)      # [B*L, T]
            B = xlr.shape[0]
            dxlr = dx_flat.reshape(B, L_, T).permute(0, 2, 1).contiguous()
        return da, dxlr, dwp, None, None, None, None, None


def ctypes_offset(t: torch.Tensor, n_elems: int):
    import ctypes
    return ctypes.c_void_p(t.data_ptr() + n_elems * t.element_size())


@_lib.guard
def inr_decode(xlr, lr_encoded, lr_coords, hr_coords, t, wp, bp, B, L_, nq, k, interpolation, *, idx=None):
    """continuous_decoder (models/magnet_gnn.py:224-283): xlr [B,T,L], lr_encoded [B*L,C], lr_coords [B*L,d],
    hr_coords [B*nq,d], t [B,>=T] -> z [B*nq, T, n_chan]."""
    from . import graph as MG
    T = xlr.shape[1]
    d = lr_coords.shape[1]
    if idx is None:
        ptr_x = MG.uniform_ptr(B, L_, xlr.device)
        ptr_y = MG.uniform_ptr(B, nq, xlr.device)
        idx = MG.knn_indices(lr_coords, hr_coords, k, ptr_x, ptr_y)
    if lr_encoded.shape[-1] != 128 or wp.shape != (128, 128 + d + 2):
        raise RuntimeError("inr_decode kernels are built for latent_dim = n_chan = 128: got lr_encoded "
                           f"{tuple(lr_encoded.shape)}, proj_head.weight {tuple(wp.shape)} (d = {d})")
    a = linear_act(lr_encoded, wp[:, :128], bp, "none", owner=wp)
    return InrBlendFn.apply(a, xlr, wp, lr_coords, hr_coords, t, idx, (nq, L_, T, d, INTERP[interpolation]))
