"""``torch.ops.mde.*``: the hot-path operators registered with torch.library on top of the C ABI (include/mde_b200.h).

north_star: "host code stays Python/PyTorch and calls hand-written sm_100a CUDA kernels through a thin C-ABI torch custom-op
layer".  ``ops.py`` is the ctypes binding of libmde_b200.so; this module registers the same entry points as PyTorch custom
operators -- schema, CUDA implementation, fake (meta) implementation for shape propagation / FakeTensor tracing, and
``register_autograd`` formulas for the differentiable ones -- so that they appear as ``torch.ops.mde.<name>`` to dispatcher-
level tooling (torch.export, FakeTensorMode, profilers, opcheck).  There is still no CPU implementation: a CPU tensor raises.

    import mde_biological_vision_systems_b200.torch_ops      # registers the library
    loss = torch.ops.mde.silog(pred, depth, mask, True)

Split-bf16 operands (ops.SplitBF16) cross this boundary as their ``planes`` tensor: bfloat16 [2, B, H, W, C].
"""
import torch
from torch.library import custom_op, register_autograd

from . import _lib, ops

_p, _s = ops._p, ops._s


def _cuda_only(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MdeError("torch.ops.mde.* operators take CUDA tensors only (no CPU fallback exists)")


# ------------------------------------------------------------------------------------------------------------
# K3 gather
# ------------------------------------------------------------------------------------------------------------
@custom_op("mde::gather_embed", mutates_args=())
def gather_embed(labels: torch.Tensor, table: torch.Tensor, background: int) -> torch.Tensor:
    """labels [B,1,H,W] int64/int32/uint8, table [rows,D] f32/f64 -> [B,D,H,W]; background < 0: no clamp (out-of-range
    labels raise IndexError like index_select)."""
    _cuda_only(labels, table)
    return ops.gather_embed(labels, table, background=None if background < 0 else background)


@gather_embed.register_fake
def _(labels, table, background):
    b, _, h, w = labels.shape
    return table.new_empty((b, table.shape[1], h, w))


# ------------------------------------------------------------------------------------------------------------
# split-bf16 pairs and the tensor-core kernels that consume them
# ------------------------------------------------------------------------------------------------------------
@custom_op("mde::split_bf16", mutates_args=())
def split_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 [B,C,H,W] (either memory format) -> bfloat16 planes [2,B,H,W,C] (hi, mid)."""
    _cuda_only(x)
    return ops.split_bf16(x).planes


@split_bf16.register_fake
def _(x):
    b, c, h, w = x.shape
    return x.new_empty((2, b, h, w, c), dtype=torch.bfloat16)


@custom_op("mde::conv3x3_x3", mutates_args=())
def conv3x3_x3(x_planes: torch.Tensor, w_planes: torch.Tensor, scale: torch.Tensor | None, shift: torch.Tensor | None,
               slope: float, pair_out: bool) -> torch.Tensor:
    """lrelu(conv3x3(x, pad 1) * scale + shift) on the tcgen05 implicit GEMM; w_planes from ops.prepare_conv3x3_weight.
    Returns fp32 channels_last [B,Cout,H,W], or bfloat16 planes [2,B,H,W,Cout] when pair_out."""
    _cuda_only(x_planes, w_planes)
    out = ops.conv3x3_nhwc(ops.SplitBF16(x_planes), w_planes, scale, shift, slope=slope, pair_out=pair_out)
    return out.planes if pair_out else out


@conv3x3_x3.register_fake
def _(x_planes, w_planes, scale, shift, slope, pair_out):
    _, b, h, w, _ = x_planes.shape
    cout = w_planes.shape[3]
    if pair_out:
        return x_planes.new_empty((2, b, h, w, cout))
    return x_planes.new_empty((b, cout, h, w), dtype=torch.float32).to(memory_format=torch.channels_last)


@custom_op("mde::patch_embed", mutates_args=())
def patch_embed(x_planes: torch.Tensor, w_planes: torch.Tensor, bias: torch.Tensor, pos: torch.Tensor, patch: int) -> torch.Tensor:
    """tokens [S,B,E] of PatchTransformerEncoder's embedding conv + positional rows (models/layers.py:16-19)."""
    _cuda_only(x_planes, w_planes, bias, pos)
    return ops.patch_embed(ops.SplitBF16(x_planes), w_planes, bias, pos, patch)


@patch_embed.register_fake
def _(x_planes, w_planes, bias, pos, patch):
    _, b, h, w, _ = x_planes.shape
    return bias.new_empty(((h // patch) * (w // patch), b, w_planes.shape[1]), dtype=torch.float32)


@custom_op("mde::fold_queries", mutates_args=())
def fold_queries(w_out: torch.Tensor, bias: torch.Tensor, queries: torch.Tensor,
                 feat_bias: torch.Tensor | None) -> tuple[torch.Tensor, torch.Tensor]:
    """(split-bf16 planes of log2e * w_out @ queries[b]  [2,B,n_bins,K],  biasf [B,n_bins])."""
    _cuda_only(w_out, bias, queries)
    return ops.fold_queries(w_out, bias, queries, feat_bias)


@fold_queries.register_fake
def _(w_out, bias, queries, feat_bias):
    b, _, k = queries.shape
    n_bins = w_out.shape[0]
    return queries.new_empty((2, b, n_bins, k), dtype=torch.bfloat16), queries.new_empty((b, n_bins), dtype=torch.float32)


@custom_op("mde::head_chain", mutates_args=())
def head_chain(x_planes: torch.Tensor, wf_planes: torch.Tensor, biasf: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    """Fused range attention -> conv_out -> softmax -> centre-weighted sum: pred [B,1,h,w]."""
    _cuda_only(x_planes, wf_planes, biasf, centers)
    return ops.head_chain(ops.SplitBF16(x_planes), wf_planes, biasf, centers)


@head_chain.register_fake
def _(x_planes, wf_planes, biasf, centers):
    _, b, h, w, _ = x_planes.shape
    return biasf.new_empty((b, 1, h, w))


# ------------------------------------------------------------------------------------------------------------
# losses (differentiable)
# ------------------------------------------------------------------------------------------------------------
@custom_op("mde::silog_fwd", mutates_args=())
def silog_fwd(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor | None, interpolate: bool) -> tuple[torch.Tensor, torch.Tensor]:
    """-> (loss scalar, workspace holding the sums the backward needs)."""
    _cuda_only(pred, target, mask)
    lib = _lib.load()
    pred, target = pred.contiguous().float(), target.contiguous().float()
    b, _, h, w = pred.shape
    hh, ww = target.shape[-2:]
    if mask is not None:
        mask = mask.expand_as(target).contiguous()
        mask = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
    ws = torch.empty(int(lib.mde_silog_ws_bytes()), dtype=torch.uint8, device=pred.device)
    loss = torch.empty((), dtype=torch.float32, device=pred.device)
    _lib.check(lib.mde_silog_fwd(_p(pred), _p(target), _p(mask), b, h, w, hh, ww, 1 if interpolate else 0, _p(ws), _p(loss),
                                 _s()), "mde_silog_fwd")
    return loss, ws


@silog_fwd.register_fake
def _(pred, target, mask, interpolate):
    return pred.new_empty((), dtype=torch.float32), pred.new_empty((_silog_ws_bytes(),), dtype=torch.uint8)


def _silog_ws_bytes():
    return int(_lib.load(check_device=False).mde_silog_ws_bytes())


@custom_op("mde::silog_bwd", mutates_args=())
def silog_bwd(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor | None, interpolate: bool, ws: torch.Tensor,
              grad: torch.Tensor) -> torch.Tensor:
    _cuda_only(pred, target, ws, grad)
    lib = _lib.load()
    pred, target = pred.contiguous().float(), target.contiguous().float()
    b, _, h, w = pred.shape
    hh, ww = target.shape[-2:]
    if mask is not None:
        mask = mask.expand_as(target).contiguous()
        mask = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
    gp = torch.empty_like(pred)
    _lib.check(lib.mde_silog_bwd(_p(pred), _p(target), _p(mask), b, h, w, hh, ww, 1 if interpolate else 0, _p(ws),
                                 _p(grad.contiguous().float()), _p(gp), _s()), "mde_silog_bwd")
    return gp


@silog_bwd.register_fake
def _(pred, target, mask, interpolate, ws, grad):
    return torch.empty_like(pred, dtype=torch.float32)


def _silog_setup(ctx, inputs, output):
    pred, target, mask, interpolate = inputs
    ctx.save_for_backward(pred, target, mask, output[1])
    ctx.interpolate = interpolate


def _silog_backward(ctx, g_loss, g_ws):
    pred, target, mask, ws = ctx.saved_tensors
    return torch.ops.mde.silog_bwd(pred, target, mask, ctx.interpolate, ws, g_loss), None, None, None


register_autograd("mde::silog_fwd", _silog_backward, setup_context=_silog_setup)


def silog(pred, target, mask=None, interpolate=True):
    """SILogLoss.forward (loss.py:12-25) through the dispatcher: torch.ops.mde.silog_fwd (+ registered autograd)."""
    return torch.ops.mde.silog_fwd(pred, target, mask, bool(interpolate))[0]


@custom_op("mde::depth_losses_fwd", mutates_args=())
def depth_losses_fwd(pred: torch.Tensor, edges: torch.Tensor, target: torch.Tensor, min_depth: float, min_target: float,
                     interpolate: bool, want_edge_grad: bool) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """SILog (mask = target > min_depth) and bins-chamfer in one pass over the target: (silog, chamfer, ws_silog, ws_chamfer)."""
    _cuda_only(pred, edges, target)
    lib = _lib.load()
    pred, edges, target = pred.contiguous().float(), edges.contiguous().float(), target.contiguous().float()
    b, _, h, w = pred.shape
    hh, ww = target.shape[-2:]
    n1 = edges.shape[1]
    ws_s = torch.empty(int(lib.mde_silog_ws_bytes()), dtype=torch.uint8, device=pred.device)
    ws_c = torch.empty(int(lib.mde_chamfer_ws_bytes(b, n1 - 1)), dtype=torch.uint8, device=pred.device)
    out = torch.empty(2, dtype=torch.float32, device=pred.device)
    rc = lib.mde_depth_losses_fwd(_p(pred), _p(edges), _p(target), b, h, w, hh, ww, n1 - 1, 1 if interpolate else 0,
                                  float(min_depth), float(min_target), 1 if want_edge_grad else 0, _p(ws_s), _p(ws_c),
                                  _p(out[0:1]), _p(out[1:2]), _s())
    _lib.check(rc, "mde_depth_losses_fwd")
    return out[0].clone(), out[1].clone(), ws_s, ws_c


@depth_losses_fwd.register_fake
def _(pred, edges, target, min_depth, min_target, interpolate, want_edge_grad):
    b, n1 = edges.shape
    lib = _lib.load(check_device=False)
    return (pred.new_empty((), dtype=torch.float32), pred.new_empty((), dtype=torch.float32),
            pred.new_empty((_silog_ws_bytes(),), dtype=torch.uint8),
            pred.new_empty((int(lib.mde_chamfer_ws_bytes(b, n1 - 1)),), dtype=torch.uint8))


def _dl_setup(ctx, inputs, output):
    pred, edges, target, min_depth, _, interpolate, _ = inputs
    ctx.save_for_backward(pred, edges, target, output[2], output[3])
    ctx.cfg = (float(min_depth), bool(interpolate))


def _dl_backward(ctx, g_s, g_c, _g2, _g3):
    lib = _lib.load()
    pred, edges, target, ws_s, ws_c = ctx.saved_tensors
    min_depth, interpolate = ctx.cfg
    pred, edges, target = pred.contiguous().float(), edges.contiguous().float(), target.contiguous().float()
    b, _, h, w = pred.shape
    hh, ww = target.shape[-2:]
    n1 = edges.shape[1]
    gp = ge = None
    if ctx.needs_input_grad[0]:
        gp = torch.empty_like(pred)
        _lib.check(lib.mde_silog_bwd_thr(_p(pred), _p(target), min_depth, b, h, w, hh, ww, 1 if interpolate else 0, _p(ws_s),
                                         _p(g_s.contiguous().float()), _p(gp), _s()), "mde_silog_bwd_thr")
    if ctx.needs_input_grad[1]:
        ge = torch.empty_like(edges)
        _lib.check(lib.mde_chamfer_bwd(_p(edges), b, n1 - 1, _p(ws_c), _p(g_c.contiguous().float()), _p(ge), _s()),
                   "mde_chamfer_bwd")
    return gp, ge, None, None, None, None, None


register_autograd("mde::depth_losses_fwd", _dl_backward, setup_context=_dl_setup)


def depth_losses(pred, edges, target, min_depth=1e-3, min_target=1e-3, interpolate=True):
    """train.py:414-419 through the dispatcher: (silog, chamfer) from torch.ops.mde.depth_losses_fwd (+ registered autograd)."""
    want = bool(torch.is_grad_enabled() and edges.requires_grad)
    s, c, _, _ = torch.ops.mde.depth_losses_fwd(pred, edges, target, float(min_depth), float(min_target), bool(interpolate), want)
    return s, c


OPS = ("gather_embed", "split_bf16", "conv3x3_x3", "patch_embed", "fold_queries", "head_chain", "silog_fwd", "silog_bwd",
       "depth_losses_fwd")
