"""Thin Python layer over the C ABI (include/mde_b200.h): tensor checks, stream plumbing, autograd glue.

Every function here launches hand-written sm_100a kernels from libmde_b200.so on torch's current CUDA stream.
Inputs must be CUDA tensors; nothing here computes on the CPU or through a PyTorch substitute.
"""
import contextlib
import ctypes

import torch

from . import _lib

LOG2E = 1.4426950408889634
NUM_SMS = 148  # B200


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _f32(t):
    """The kernels are fp32: tensors produced under autocast (bf16 / fp16 activations of the stock torch modules) are
    widened here -- handing their storage to a kernel as-is would be read as garbage."""
    return t if t.dtype == torch.float32 else t.float()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MdeError("mde_b200 operators take CUDA tensors only (no CPU fallback exists)")


@contextlib.contextmanager
def exact_fp32_library():
    """Run the enclosed cuDNN / cuBLAS calls (the passthrough bodies: EfficientNet encoder, decoder conv2, shapes our
    kernels do not cover) in true fp32 rather than the libraries' TF32 default, so that the whole inference path is
    fp32-grade -- north_star's 1e-3 on depth maps is against the fp32 reference."""
    c, m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32 = c
        torch.backends.cuda.matmul.allow_tf32 = m


_PRODUCTS = [3]


@contextlib.contextmanager
def bf16_products(enabled=True):
    """The bf16 mode of the tensor-core kernels (north star: depth and bin edges within 2e-2 of the fp32 reference): inside this
    context conv3x3_nhwc and head_chain multiply the hi planes of their split-bf16 operands only -- ONE bf16 product per K
    step, fp32 accumulation -- instead of the three products of the fp32-grade default.  Operand formats do not change."""
    old = _PRODUCTS[0]
    if enabled:
        _PRODUCTS[0] = 1
    try:
        yield
    finally:
        _PRODUCTS[0] = old


def products():
    """3 (fp32-grade, default) or 1 (inside ops.bf16_products())."""
    return _PRODUCTS[0]


def launch_count():
    return int(_lib.load(False).mde_launch_count())


# Optional per-kernel CUDA-event timing (bench.py's roofline numbers are measured live with this, on the launching
# stream): timing("head_chain") returns a context manager that records an event pair when enabled, else a no-op.
_TIMING = {"on": False, "records": [], "work": {}}


class _Timed:
    def __init__(self, name, work=None):
        self.name = name
        if work is not None and _TIMING["on"]:
            _TIMING["work"][name] = work

    def __enter__(self):
        if _TIMING["on"]:
            self.start = torch.cuda.Event(enable_timing=True)
            self.end = torch.cuda.Event(enable_timing=True)
            self.start.record()
        return self

    def __exit__(self, *exc):
        if _TIMING["on"]:
            self.end.record()
            _TIMING["records"].append((self.name, self.start, self.end))
        return False


def timing(name, work=None):
    """``work``: algorithmic work of the launch (FLOPs or bytes), recorded next to the name for the bench's roofline table."""
    return _Timed(name, work)


def enable_kernel_timing(on=True):
    _TIMING["on"] = bool(on)
    _TIMING["records"] = []
    if on:
        _TIMING["work"] = {}


def kernel_work():
    """name -> algorithmic work per launch recorded by timing(name, work=...) while kernel timing was enabled."""
    return dict(_TIMING["work"])


def kernel_times_ms():
    """name -> list of per-launch durations (ms); call after torch.cuda.synchronize()."""
    out = {}
    for name, s, e in _TIMING["records"]:
        out.setdefault(name, []).append(s.elapsed_time(e))
    return out


# ------------------------------------------------------------------------------------------------------------
# split-bf16 pairs ("bf16x3"): the operand format of the tensor-core kernels (conv3x3, patch embedding, fused chain)
# ------------------------------------------------------------------------------------------------------------
class SplitBF16:
    """An fp32 feature map [B,C,H,W] carried as two bf16 planes in NHWC element order: ``planes`` is a contiguous
    torch.bfloat16 tensor [2,B,H,W,C] with planes[0] = bf16(v) and planes[1] = bf16(v - planes[0]).  The tensor-core
    kernels multiply such operands as hi*hi + mid*hi + hi*mid with fp32 accumulation (~2^-17 relative error per product)."""
    __slots__ = ("planes",)

    def __init__(self, planes):
        if planes.dtype != torch.bfloat16 or planes.dim() != 5 or planes.shape[0] != 2 or not planes.is_contiguous():
            raise ValueError("SplitBF16 wraps a contiguous bfloat16 [2,B,H,W,C] tensor")
        self.planes = planes

    @property
    def shape(self):  # logical NCHW shape
        _, b, h, w, c = self.planes.shape
        return torch.Size((b, c, h, w))

    @property
    def device(self):
        return self.planes.device

    @property
    def is_cuda(self):
        return self.planes.is_cuda

    def float(self):
        """hi + mid as a channels_last fp32 [B,C,H,W] tensor."""
        return merge_bf16(self)


def split_bf16_flat(x):
    """fp32 tensor (any shape, numel % 8 == 0) -> bfloat16 [2, *x.shape] (plane 0 = hi, plane 1 = mid), same element order."""
    lib = _lib.load()
    _need_cuda(x)
    x = _f32(x).contiguous()
    out = torch.empty((2,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.mde_split_bf16(_p(x), _p(out), x.numel(), _s()), "mde_split_bf16")
    return out


def split_bf16(x):
    """fp32 [B,C,H,W] (NCHW-contiguous or channels_last) -> SplitBF16 (NHWC planes); one streaming pass either way."""
    lib = _lib.load()
    _need_cuda(x)
    if isinstance(x, SplitBF16):
        return x
    x = _f32(x)
    b, c, h, w = x.shape
    planes = torch.empty((2, b, h, w, c), dtype=torch.bfloat16, device=x.device)
    with timing("split_bf16"):
        if x.is_contiguous(memory_format=torch.channels_last) and (c * h * w * b) % 8 == 0:
            rc = lib.mde_split_bf16(_p(x), _p(planes), x.numel(), _s())
        else:
            rc = lib.mde_split_bf16_nchw(_p(x.contiguous()), _p(planes), b, c, h * w, _s())
    _lib.check(rc, "mde_split_bf16")
    return SplitBF16(planes)


def merge_bf16(pair):
    lib = _lib.load()
    _, b, h, w, c = pair.planes.shape
    out = torch.empty((b, c, h, w), dtype=torch.float32, device=pair.planes.device, memory_format=torch.channels_last)
    _lib.check(lib.mde_merge_bf16(_p(pair.planes), _p(out), out.numel(), _s()), "mde_merge_bf16")
    return out


# ------------------------------------------------------------------------------------------------------------
# K3: gathers
# ------------------------------------------------------------------------------------------------------------
_LABEL_DTYPES = {torch.int64: 0, torch.int32: 1, torch.uint8: 2}


def gather_embed(labels, table, background=None, write_back=False, out=None, table_image_stride=0, labels_out=None):
    """labels [B,1,H,W] int64 (the reference's batch tensors) or int32 / uint8 (the on-disk label formats, "next" row
    (f)3); table [rows,D] (f32/f64, CUDA) -> [B,D,H,W] of table.dtype.
    background: clamp target for labels outside [0,rows-1] (None = no clamp; out-of-range raises IndexError like
    the reference's index_select).  write_back=True stores the clamped labels into ``labels`` (int64 only; the
    reference clamps the batch tensor in place: SemanticsLoader.py:115-118); ``labels_out`` (int64 [B,1,H,W]) receives
    them for any label dtype."""
    lib = _lib.load()
    _need_cuda(labels, table)
    if labels.dtype not in _LABEL_DTYPES or labels.dim() != 4 or labels.shape[1] != 1 or not labels.is_contiguous():
        raise ValueError("labels must be a contiguous int64 / int32 / uint8 [B,1,H,W] tensor")
    if table.dtype not in (torch.float32, torch.float64) or not table.is_contiguous():
        raise ValueError("table must be contiguous float32/float64")
    if write_back:
        if labels.dtype != torch.int64:
            raise ValueError("write_back needs int64 labels (pass labels_out for compact label types)")
        labels_out = labels
    if labels_out is not None and (labels_out.dtype != torch.int64 or labels_out.shape != labels.shape
                                   or not labels_out.is_contiguous()):
        raise ValueError("labels_out must be a contiguous int64 tensor of the labels' shape")
    b, _, h, w = labels.shape
    if table_image_stride:
        rows, d = table.shape[1], table.shape[2]
    else:
        rows, d = table.shape
    if out is None:
        out = torch.empty((b, d, h, w), dtype=table.dtype, device=labels.device)
    if labels.numel() == 0:
        return out
    flag = None
    if background is None:
        flag = torch.zeros(1, dtype=torch.int32, device=labels.device)
    with timing("gather_embed"):
        rc = lib.mde_gather_embed_labels(_p(labels), _LABEL_DTYPES[labels.dtype], _p(labels_out), _p(table), _p(out), b,
                                         h * w, rows, d, -1 if background is None else int(background),
                                         0 if table.dtype == torch.float32 else 1, int(table_image_stride), _p(flag), _s())
    _lib.check(rc, "mde_gather_embed_labels")
    if flag is not None:
        if _DEFERRED is not None and torch.cuda.is_current_stream_capturing():
            _DEFERRED.append((flag, rows))  # a captured launch cannot raise: the owner of the graph reads the flag after a replay
        elif int(flag.item()) != 0:
            raise IndexError("index out of range in label gather (table has %d rows)" % rows)
    return out


_DEFERRED = None


def begin_deferred_checks():
    """CUDA-graph capture of a step that contains un-clamped gathers (150-class table): collect their out-of-range flags instead
    of reading them (a device -> host read is illegal during capture).  Pair with end_deferred_checks()."""
    global _DEFERRED
    _DEFERRED = []


def end_deferred_checks():
    """-> list of (flag tensor, table rows) recorded since begin_deferred_checks(); the flags are rewritten by every replay."""
    global _DEFERRED
    out, _DEFERRED = (_DEFERRED or []), None
    return out


def raise_deferred_checks(deferred):
    """IndexError (like the reference's index_select) if a replayed gather met an out-of-range label.  Synchronises."""
    for flag, rows in deferred:
        if int(flag.item()) != 0:
            raise IndexError("index out of range in label gather (table has %d rows)" % rows)


def gather_embed_nhwc(labels, table, background, c_before, c_after=0, pads=(0, 0, 0, 0), labels_out=None, image=None):
    """Clamp + gather straight into a channels_last buffer laid out like the encoder input: returns (buffer, view) where
    ``buffer`` is a zero-bordered channels_last [B, c_before + D + c_after, H + top + bottom, W + left + right] fp32 tensor
    whose channels [c_before, c_before + D) hold the embeddings, and ``view`` = the [B, D, H, W] embedding tensor the
    reference's loader returns -- a strided view into the buffer (no planar copy exists).  With ``image`` (NCHW fp32
    [B, c_before, H, W]) the leading channels are written by the same kernel; otherwise the caller fills them
    (ops.fill_channels_nhwc) before feeding ``buffer`` to the encoder."""
    lib = _lib.load()
    _need_cuda(labels, table)
    if labels.dtype not in _LABEL_DTYPES or labels.dim() != 4 or labels.shape[1] != 1 or not labels.is_contiguous():
        raise ValueError("labels must be a contiguous int64 / int32 / uint8 [B,1,H,W] tensor")
    if table.dtype != torch.float32 or not table.is_contiguous():
        raise ValueError("gather_embed_nhwc takes a contiguous float32 table")
    b, _, h, w = labels.shape
    rows, d = table.shape
    pt, pb, pl, pr = (int(v) for v in pads)
    ho, wo, pitch = h + pt + pb, w + pl + pr, c_before + d + c_after
    buf = torch.empty((b, pitch, ho, wo), dtype=torch.float32, device=labels.device, memory_format=torch.channels_last)
    if pt:
        buf[:, :, :pt].zero_()
    if pb:
        buf[:, :, h + pt:].zero_()
    if pl:
        buf[:, :, :, :pl].zero_()
    if pr:
        buf[:, :, :, w + pl:].zero_()
    if image is not None:
        if image.dtype != torch.float32 or tuple(image.shape) != (b, c_before, h, w) or not image.is_cuda:
            raise ValueError("image must be a CUDA float32 [B, c_before, H, W] tensor")
        image = image.contiguous()
    with timing("gather_embed", work=float(b * h * w * (labels.element_size() + 4 * d))):
        rc = lib.mde_gather_embed_nhwc(_p(labels), _LABEL_DTYPES[labels.dtype], _p(labels_out), _p(table), _p(image), _p(buf), b,
                                       h, w, rows, d, int(background), pitch, c_before, ho, wo, pt, pl, _s())
    _lib.check(rc, "mde_gather_embed_nhwc")
    view = buf[:, c_before:c_before + d, pt:pt + h, pl:pl + w]
    return buf, view


def fill_channels_nhwc(buf, src, c0, pads=(0, 0, 0, 0)):
    """Write the NCHW-contiguous fp32 ``src`` [B,C,H,W] into channels [c0, c0 + C) of the (padded) channels_last ``buf``."""
    lib = _lib.load()
    src = _f32(src).contiguous()
    b, c, h, w = src.shape
    pt, pb, pl, pr = (int(v) for v in pads)
    ctot = buf.shape[1]
    dst = ctypes.c_void_p(buf.data_ptr() + 4 * c0)
    with timing("nchw_to_nhwc"):
        if pt or pb or pl or pr:
            rc = lib.mde_nchw_to_nhwc_slice_padded(_p(src), dst, b, c, h, w, ctot, pt, pb, pl, pr, _s())
        else:
            rc = lib.mde_nchw_to_nhwc_slice(_p(src), dst, b, c, h * w, ctot, _s())
    _lib.check(rc, "mde_nchw_to_nhwc_slice")
    return buf


def class_area_fraction(labels, rows):
    """SemanticsLoader.get_semantics_inst_areas: float64 [B,1,H,W] of per-image class pixel fractions."""
    lib = _lib.load()
    _need_cuda(labels)
    b, _, h, w = labels.shape
    counts = torch.empty((b, rows), dtype=torch.int32, device=labels.device)
    frac = torch.empty((b, rows, 1), dtype=torch.float64, device=labels.device)
    _lib.check(lib.mde_class_area_table(_p(labels), b, h * w, rows, _p(counts), _p(frac), _s()), "mde_class_area_table")
    return gather_embed(labels, frac, background=None, table_image_stride=rows)


def cast_i64_f32(x):
    lib = _lib.load()
    _need_cuda(x)
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    _lib.check(lib.mde_cast_i64_f32(_p(x), _p(out), x.numel(), _s()), "mde_cast_i64_f32")
    return out


# ------------------------------------------------------------------------------------------------------------
# A3: per-pixel MLP
# ------------------------------------------------------------------------------------------------------------
def _plane_view_ok(t):
    b, c, h, w = t.shape
    return t.stride(3) == 1 and t.stride(2) == w and t.stride(1) == h * w


def _aux_mlp_launch(x, w0, b0, w1, b1, in_div, out):
    lib = _lib.load()
    b, c, h, w = x.shape
    rc = lib.mde_aux_mlp_fwd(_p(x), x.stride(0), _p(w0), _p(b0), _p(w1), _p(b1), _p(out), out.stride(0), b, c,
                             w0.shape[0], w1.shape[0], h * w, float(in_div), _s())
    _lib.check(rc, "mde_aux_mlp_fwd")
    return out


class _AuxMlp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w0, b0, w1, b1, in_div):
        b, c, h, w = x.shape
        out = torch.empty((b, w1.shape[0], h, w), dtype=torch.float32, device=x.device)
        _aux_mlp_launch(x, w0, b0, w1, b1, in_div, out)
        ctx.save_for_backward(x, w0, b0, w1, b1)
        ctx.in_div = float(in_div)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        x, w0, b0, w1, b1 = ctx.saved_tensors
        b, c, h, w = x.shape
        if not _plane_view_ok(gout):
            gout = gout.contiguous()
        gw0, gb0, gw1, gb1 = (torch.zeros_like(t) for t in (w0, b0, w1, b1))
        rc = lib.mde_aux_mlp_bwd(_p(x), x.stride(0), _p(w0), _p(b0), _p(w1), _p(b1), _p(gout), gout.stride(0), None,
                                 _p(gw0), _p(gb0), _p(gw1), _p(gb1), b, c, w0.shape[0], w1.shape[0], h * w, ctx.in_div,
                                 _s())
        _lib.check(rc, "mde_aux_mlp_bwd")
        return None, gw0, gb0, gw1, gb1, None


def aux_mlp(x, w0, b0, w1, b1, in_div=1.0, out=None):
    """relu(conv1x1(relu(conv1x1(x / in_div)))) with conv weights [10,C,1,1] / [10,10,1,1] (or 2-D).
    ``out`` (a channel slice of a contiguous NCHW tensor) is only honoured when no gradient is being recorded."""
    _need_cuda(x, w0, b0, w1, b1)
    if x.dtype != torch.float32 or not _plane_view_ok(x):
        x = x.float().contiguous()
    w0f, w1f = w0.reshape(w0.shape[0], -1).contiguous(), w1.reshape(w1.shape[0], -1).contiguous()
    b0, b1 = b0.contiguous(), b1.contiguous()
    needs_grad = torch.is_grad_enabled() and any(t.requires_grad for t in (w0, b0, w1, b1))
    if not needs_grad:
        if out is None:
            out = torch.empty((x.shape[0], w1f.shape[0], x.shape[2], x.shape[3]), dtype=torch.float32, device=x.device)
        elif not _plane_view_ok(out):
            raise ValueError("out must be a channel slice of a contiguous NCHW tensor")
        return _aux_mlp_launch(x, w0f.detach(), b0.detach(), w1f.detach(), b1.detach(), in_div, out)
    res = _AuxMlp.apply(x, w0f, b0, w1f, b1, in_div)
    if out is not None:
        out.copy_(res)
        return out
    return res


def bias_act_nhwc_(x_cl, bias, act=0, residual=None):
    """In place on a channels_last tensor: x = act(x + bias[c]) (+ residual); act 0 none / 1 SiLU."""
    lib = _lib.load()
    b, c, h, w = x_cl.shape
    rc = lib.mde_bias_act_nhwc(_p(x_cl), _p(bias), _p(residual), _p(x_cl), b * h * w, c, int(act), _s())
    _lib.check(rc, "mde_bias_act_nhwc")
    return x_cl


def bias_act_pad_nhwc(x_cl, bias, act, pads):
    """pad(act(x + bias[c])) into a new channels_last tensor; pads = (top, bottom, left, right) zeros."""
    lib = _lib.load()
    b, c, h, w = x_cl.shape
    pt, pb, pl, pr = (int(v) for v in pads)
    out = torch.empty((b, c, h + pt + pb, w + pl + pr), dtype=torch.float32, device=x_cl.device,
                      memory_format=torch.channels_last)
    rc = lib.mde_bias_act_pad_nhwc(_p(x_cl), _p(bias), _p(out), b, h, w, c, pt, pb, pl, pr, int(act), _s())
    _lib.check(rc, "mde_bias_act_pad_nhwc")
    return out


def bias_act_supported(x, residual=None):
    ok = (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] % 4 == 0 and _is_nhwc_or_dense_cl(x))
    if residual is not None:
        ok = ok and residual.shape == x.shape and residual.dtype == torch.float32 and _is_nhwc_or_dense_cl(residual)
    return ok


def _is_nhwc_or_dense_cl(x):
    return x.is_contiguous(memory_format=torch.channels_last)


# ------------------------------------------------------------------------------------------------------------
# regressor + bins
# ------------------------------------------------------------------------------------------------------------
_NORM = {"linear": 0, "softmax": 1, "sigmoid": 2}


def linear(x, weight, bias, act=0):
    """act(x @ weight.T + bias) for a 2-D x whose rows may be strided (fp32 SIMT kernel); act 0/1/2 = none/ReLU/LeakyReLU."""
    lib = _lib.load()
    x, weight = _f32(x), _f32(weight)
    if x.stride(1) != 1:
        x = x.contiguous()
    m, k = x.shape
    n = weight.shape[0]
    out = torch.empty((m, n), dtype=torch.float32, device=x.device)
    rc = lib.mde_linear_fwd(_p(x), x.stride(0), _p(weight.contiguous()), k, _p(bias.contiguous()) if bias is not None else None,
                            _p(out), n, m, n, k, int(act), _s())
    _lib.check(rc, "mde_linear_fwd")
    return out


def regressor_bins(t0, w1, b1, w2, b2, w3, b3, norm, min_val, max_val, split=False):
    """t0 [B,E] (rows may be strided) -> (widths_normed [B,n], edges [B,n+1], centers [B,n], y_raw [B,n]).
    split=False (default): ONE launch, a 1024-thread CTA per image walks the three dense layers (four output rows per warp
    pass, 16-byte weight loads), the normalisation and the edge scan.  split=True: the three layers as separate launches of
    the generic SIMT linear kernel plus a finalise kernel (4 launches, 64 us at B = 16: M = 16 rows cannot fill a 64 x 64
    tile grid; it used to be the faster form while the fused kernel computed one row per warp with 4-byte loads)."""
    lib = _lib.load()
    _need_cuda(t0, w1, w2, w3)
    t0 = _f32(t0)
    if t0.stride(1) != 1:
        t0 = t0.contiguous()
    b, e = t0.shape
    hdim, n = w1.shape[0], w3.shape[0]
    dev = t0.device
    wn = torch.empty((b, n), dtype=torch.float32, device=dev)
    edges = torch.empty((b, n + 1), dtype=torch.float32, device=dev)
    centers = torch.empty((b, n), dtype=torch.float32, device=dev)
    if split:
        with timing("regressor_bins"):
            y_raw = linear(linear(linear(t0, w1, b1, 2), w2, b2, 2), w3, b3, 0)
            rc = lib.mde_bins_finalize_fwd(_p(y_raw), b, n, _NORM.get(norm, 2), float(min_val), float(max_val), _p(wn),
                                           _p(edges), _p(centers), _s())
        _lib.check(rc, "mde_bins_finalize_fwd")
        return wn, edges, centers, y_raw
    y_raw = torch.empty((b, n), dtype=torch.float32, device=dev)
    with timing("regressor_bins"):
        rc = lib.mde_regressor_bins_fwd(_p(t0), t0.stride(0), _p(w1.contiguous()), _p(b1.contiguous()), _p(w2.contiguous()),
                                        _p(b2.contiguous()), _p(w3.contiguous()), _p(b3.contiguous()), b, e, hdim, n,
                                        _NORM.get(norm, 2), float(min_val), float(max_val), _p(y_raw), _p(wn), _p(edges),
                                        _p(centers), _s())
    _lib.check(rc, "mde_regressor_bins_fwd")
    return wn, edges, centers, y_raw


# ------------------------------------------------------------------------------------------------------------
# patch embedding
# ------------------------------------------------------------------------------------------------------------
def prepare_patch_weight(weight):
    """Conv filter [E,C,p,p] -> channels_last order [E,p,p,C] as a split-bf16 pair, bfloat16 [2,E,p,p,C] (the B operand of
    the patch-embedding GEMM as it lies in memory)."""
    return split_bf16_flat(weight.detach().float().permute(0, 2, 3, 1).contiguous())


def patch_embed_supported(x, conv):
    """x: SplitBF16 or a CUDA fp32 [B,C,H,W] tensor."""
    k = conv.kernel_size
    b, c, h, w = x.shape
    return (x.is_cuda and conv.out_channels == 128 and k[0] == k[1] and conv.stride == k and conv.padding == (0, 0)
            and (k[0] * c) % 64 == 0 and c % 8 == 0 and h >= k[0] and 1 <= w // k[0] <= 128
            and -(-(h // k[0]) // (128 // (w // k[0]))) <= 4)


def patch_embed(x, w_prepared, bias, pos, patch):
    """tokens [S,B,E] = conv(k = s = patch)(x).flatten(2).permute(2,0,1) + pos[:S] on the tcgen05 split-K GEMM (three bf16
    products per K step).  x: SplitBF16 (or an fp32 tensor, split here)."""
    lib = _lib.load()
    x = split_bf16(x)
    b, c, h, w = x.shape
    e = w_prepared.shape[1]
    s = (h // patch) * (w // patch)
    tokens = torch.empty((s, b, e), dtype=torch.float32, device=x.device)
    ws = torch.empty(int(lib.mde_patch_embed_ws_floats(b, h, w, patch, c)), dtype=torch.float32, device=x.device)
    with timing("patch_embed"):
        rc = lib.mde_patch_embed_fwd(_p(x.planes), _p(w_prepared), _p(bias.contiguous()), _p(pos.contiguous()), _p(tokens),
                                     _p(ws), b, h, w, c, patch, e, _s())
    _lib.check(rc, "mde_patch_embed_fwd")
    return tokens


# ------------------------------------------------------------------------------------------------------------
# batched NT GEMM (tcgen05, TF32)
# ------------------------------------------------------------------------------------------------------------
def gemm_nt(a, b, out=None, splits=1, alpha=1.0):
    """out[..., m, n] (+)= alpha * sum_k a[..., m, k] * b[..., n, k] on the tcgen05 GEMM (TF32 inputs).  a [batch,M,K] or
    [M,K], b [batch,N,K] or [N,K]; rows must be dense along k (stride 1), row / batch pitches multiples of 4 floats.
    splits > 1: split-K with atomic accumulation into ``out`` (zeroed here unless the caller passes it)."""
    lib = _lib.load()
    _need_cuda(a, b)
    squeeze = a.dim() == 2
    if squeeze:
        a, b = a.unsqueeze(0), b.unsqueeze(0)
    if a.stride(2) != 1:
        a = a.contiguous()
    if b.stride(2) != 1:
        b = b.contiguous()
    batch, m, k = a.shape
    n = b.shape[1]
    if out is None:
        out = (torch.zeros if splits > 1 else torch.empty)((batch, m, n), dtype=torch.float32, device=a.device)
    o3 = out if out.dim() == 3 else out.unsqueeze(0)
    if o3.stride(2) != 1:
        raise ValueError("gemm_nt: out must be dense along its last dimension")
    with timing("gemm_nt"):
        rc = lib.mde_gemm_nt_tf32(_p(a), a.stride(1), a.stride(0), _p(b), b.stride(1), b.stride(0), _p(o3), o3.stride(1),
                                  o3.stride(0), batch, m, n, k, int(splits), float(alpha), _s())
    _lib.check(rc, "mde_gemm_nt_tf32")
    return out[0] if (squeeze and out.dim() == 3) else out


# ------------------------------------------------------------------------------------------------------------
# 3x3 convolution (tcgen05 implicit GEMM, NHWC)
# ------------------------------------------------------------------------------------------------------------
def prepare_conv3x3_weight(weight, cin_pad_to=1):
    """Conv filter [Cout,C,3,3] -> [dx][dy][Cout][C] as a split-bf16 pair, bfloat16 [2,3,3,Cout,C] (the B operand tiles as
    TMA reads them).  ``cin_pad_to``: C is rounded up to a multiple of it with zero input channels (for an input whose
    channel pitch was padded the same way, ops.upsample_concat_nhwc_pair(pad_to=...))."""
    lib = _lib.load()
    _need_cuda(weight)
    w = weight.detach().float()
    if w.shape[1] % cin_pad_to:
        w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, cin_pad_to - w.shape[1] % cin_pad_to))
    w = w.contiguous()
    cout, c = w.shape[0], w.shape[1]
    out = torch.empty((2, 3, 3, cout, c), dtype=torch.bfloat16, device=w.device)
    _lib.check(lib.mde_conv3x3_prep_weight_x3(_p(w), _p(out), cout, c, _s()), "mde_conv3x3_prep_weight_x3")
    return out


def prepare_conv3x3_weight_tf32(weight, operand_scale=1.0):
    """Conv filter [Cout,C,3,3] -> fp32 [dx][dy][Cout][C], TF32-rounded (single-pass TF32 form of the kernel)."""
    lib = _lib.load()
    _need_cuda(weight)
    w = weight.detach().contiguous().float()
    cout, c = w.shape[0], w.shape[1]
    out = torch.empty((3, 3, cout, c), dtype=torch.float32, device=w.device)
    _lib.check(lib.mde_conv3x3_prep_weight(_p(w), _p(out), cout, c, float(operand_scale), _s()), "mde_conv3x3_prep_weight")
    return out


def conv3x3_cout_ok(cout, pair_out=False):
    return bool(cout >= 8 and cout % (8 if pair_out else 4) == 0)


def conv3x3_supported(x, cout, pair_out=False):
    """x: SplitBF16 or CUDA fp32 [B,C,H,W]."""
    return bool(x.is_cuda and x.shape[1] % 8 == 0 and conv3x3_cout_ok(cout, pair_out))


def conv3x3_nhwc(x, w_prep, scale=None, shift=None, slope=1.0, pair_out=False, name="conv3x3"):
    """y = lrelu(conv3x3(x, pad 1) * scale + shift) on the tcgen05 implicit GEMM, three bf16 products per K step (fp32-grade).
    x: SplitBF16 (an fp32 [B,C,H,W] tensor is split first); w_prep from prepare_conv3x3_weight.  Returns a channels_last fp32
    [B,Cout,H,W] tensor, or a SplitBF16 for the next tensor-core consumer when ``pair_out``."""
    lib = _lib.load()
    x = split_bf16(x)
    _need_cuda(x.planes, w_prep)
    b, c, h, w = x.shape
    cout = w_prep.shape[3]
    if w_prep.dtype != torch.bfloat16 or w_prep.shape[4] != c:
        raise ValueError("conv3x3_nhwc: w_prep must come from prepare_conv3x3_weight for this input width")
    if pair_out:
        out = torch.empty((2, b, h, w, cout), dtype=torch.bfloat16, device=x.device)
    else:
        out = torch.empty((b, cout, h, w), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    with timing(name, work=2.0 * b * h * w * cout * 9 * c):
        rc = lib.mde_conv3x3_nhwc_x3_fwd(_p(x.planes), _p(w_prep), _p(scale), _p(shift), _p(out), 1 if pair_out else 0, b, h, w,
                                         c, cout, float(slope), products(), _s())
    _lib.check(rc, "mde_conv3x3_nhwc_x3_fwd")
    return SplitBF16(out) if pair_out else out


def conv3x3_nhwc_tf32(x_cl, w_prep, scale=None, shift=None, slope=1.0, round_tf32=False, out=None):
    """Single-pass TF32 form (x_cl's values must already be TF32-representable; w_prep from prepare_conv3x3_weight_tf32)."""
    lib = _lib.load()
    _need_cuda(x_cl, w_prep)
    b, c, h, w = x_cl.shape
    if x_cl.dtype != torch.float32 or not x_cl.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("conv3x3_nhwc_tf32 expects a float32 channels_last tensor")
    cout = w_prep.shape[2]
    if out is None:
        out = torch.empty((b, cout, h, w), dtype=torch.float32, device=x_cl.device, memory_format=torch.channels_last)
    with timing("conv3x3_tf32"):
        rc = lib.mde_conv3x3_nhwc_fwd(_p(x_cl), _p(w_prep), _p(scale), _p(shift), _p(out), b, h, w, c, cout, float(slope),
                                      1 if round_tf32 else 0, _s())
    _lib.check(rc, "mde_conv3x3_nhwc_fwd")
    return out


class _Conv3x3Fn(torch.autograd.Function):
    """Training form of the 3x3 / stride 1 / pad 1 convolution on our kernels (autograd of models/miniViT.py:16 and the
    DecoderBN convs, models/unet_adaptive_bins.py:39-49,73):
      forward : tcgen05 implicit GEMM on split-bf16 pairs (three bf16 products per K step, as in inference);
      dgrad   : the SAME kernel on the spatially flipped, channel-transposed filter applied to the output gradient;
      wgrad   : ONE launch of the TF32 NT GEMM with the nine taps on the grid (mde_conv3x3_wgrad_tf32), both operands laid out
                channel-major over the zero-padded pixel axis and TF32-rounded by a transpose kernel (split-K over pixels);
      dbias   : a column sum (torch)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        xp = split_bf16(x)
        y = conv3x3_nhwc(xp, prepare_conv3x3_weight(weight), None, bias, name="conv3x3_train_fwd")
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x, weight = ctx.saved_tensors
        gy = _f32(gy)
        if not gy.is_contiguous(memory_format=torch.channels_last):
            gy = gy.contiguous(memory_format=torch.channels_last)
        b, c, h, w = x.shape
        cout = weight.shape[0]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            wt = weight.detach().flip(2, 3).transpose(0, 1).contiguous()   # [C, Cout, 3, 3]
            gx = conv3x3_nhwc(split_bf16(gy), prepare_conv3x3_weight(wt), name="conv3x3_dgrad")
        if ctx.needs_input_grad[1]:
            wp = (w + 2 + 3) // 4 * 4           # padded row pitch: vertical taps become 16-byte aligned shifts
            kp = b * (h + 2) * wp
            x_cl = x if x.is_contiguous(memory_format=torch.channels_last) else x.contiguous(memory_format=torch.channels_last)
            xt = torch.empty((c, kp), dtype=torch.float32, device=x.device)
            gt3 = torch.empty((3, cout, kp), dtype=torch.float32, device=x.device)   # horizontal taps baked in
            gw9 = torch.empty((9, cout, c), dtype=torch.float32, device=x.device)
            with timing("conv3x3_wgrad", work=2.0 * b * h * w * cout * 9 * c):
                _lib.check(lib.mde_nhwc_to_cpad_tf32(_p(x_cl), _p(xt), b, h, w, c, wp, 0, _s()), "mde_nhwc_to_cpad_tf32")
                _lib.check(lib.mde_nhwc_to_cpad_tf32(_p(gy), _p(gt3), b, h, w, cout, wp, 1, _s()), "mde_nhwc_to_cpad_tf32")
                tiles = ((cout + 127) // 128) * max(1, (c + 255) // 256) * 9
                splits = max(1, min(64, (2 * NUM_SMS) // tiles, kp // 4096))
                rc = lib.mde_conv3x3_wgrad_tf32(_p(gt3), _p(xt), _p(gw9), cout, c, kp, kp, wp, splits, _s())
            _lib.check(rc, "mde_conv3x3_wgrad_tf32")
            gw = gw9.view(3, 3, cout, c).permute(2, 3, 0, 1).contiguous(memory_format=torch.channels_last) \
                if weight.is_contiguous(memory_format=torch.channels_last) and not weight.is_contiguous() \
                else gw9.view(3, 3, cout, c).permute(2, 3, 0, 1).contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(dim=(0, 2, 3))
        return gx, gw, gb


def conv3x3_train_supported(x, conv):
    return bool(x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and conv.kernel_size == (3, 3) and conv.stride == (1, 1)
                and conv.padding == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1 and x.shape[1] % 8 == 0
                and conv.out_channels % 8 == 0 and conv3x3_cout_ok(conv.out_channels) and conv3x3_cout_ok(x.shape[1]))


def conv3x3_autograd(x, weight, bias=None):
    """conv2d(x, weight, bias, padding=1) with forward, input gradient and weight gradient on the tcgen05 kernels; returns a
    channels_last fp32 tensor."""
    _need_cuda(x, weight)
    return _Conv3x3Fn.apply(_f32(x), weight, bias)


def conv3x3_small(x_cl, weight, bias=None):
    """Exact-fp32 direct 3x3 conv for Cout <= 4 (the noAdaBins decoder's conv3): x_cl channels_last fp32, torch-layout
    weight [Cout,C,3,3] -> [B,Cout,H,W]."""
    lib = _lib.load()
    _need_cuda(x_cl, weight)
    x_cl = _f32(x_cl)
    if not x_cl.is_contiguous(memory_format=torch.channels_last):
        x_cl = x_cl.contiguous(memory_format=torch.channels_last)
    b, c, h, w = x_cl.shape
    cout = weight.shape[0]
    out = torch.empty((b, cout, h, w), dtype=torch.float32, device=x_cl.device, memory_format=torch.channels_last)
    with timing("conv3x3_small"):
        rc = lib.mde_conv3x3_small_nhwc_fwd(_p(x_cl), _p(weight.detach().contiguous().float()),
                                            _p(bias.detach().contiguous().float()) if bias is not None else None, _p(out), b, h, w,
                                            c, cout, _s())
    _lib.check(rc, "mde_conv3x3_small_nhwc_fwd")
    return out


# ------------------------------------------------------------------------------------------------------------
# 1x1 convolution (tcgen05 GEMM, activations split to bf16 pairs inside the kernel)
# ------------------------------------------------------------------------------------------------------------
def prepare_pointwise_weight(weight):
    """Conv filter [Cout,Cin,1,1] (or [Cout,Cin]) -> split-bf16 pair, bfloat16 [2,Cout,Cin]."""
    w = weight.detach().float().reshape(weight.shape[0], -1).contiguous()
    return split_bf16_flat(w)


def pointwise_supported(x, cin, cout):
    return bool(x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and cin % 8 == 0 and cout % 4 == 0
                and x.is_contiguous(memory_format=torch.channels_last) and x.shape[0] * x.shape[2] * x.shape[3] < 2 ** 31)


def pointwise_conv(x_cl, w_pair, bias=None, act=0, residual=None, name="pointwise", gate=None, out_pads=None):
    """act(conv1x1(x * gate) + bias) (+ residual) on the tcgen05 GEMM, three bf16 products per K step (fp32-grade): x_cl
    channels_last fp32 [B,Cin,H,W], w_pair from prepare_pointwise_weight; act 0 none / 1 SiLU; residual channels_last
    [B,Cout,H,W]; gate fp32 [B,Cin] (the squeeze-excite scale, multiplied into x in fp32 inside the kernel); out_pads
    (top, bottom, left, right): the result is written inside a zero border of that size.
    Returns a channels_last fp32 [B,Cout,H(+pads),W(+pads)] tensor."""
    import ctypes
    lib = _lib.load()
    _need_cuda(x_cl, w_pair)
    b, c, h, w = x_cl.shape
    cout = w_pair.shape[1]
    if w_pair.dtype != torch.bfloat16 or w_pair.shape[2] != c:
        raise ValueError("pointwise_conv: w_pair must come from prepare_pointwise_weight for this input width")
    if not x_cl.is_contiguous(memory_format=torch.channels_last) or x_cl.dtype != torch.float32:
        raise ValueError("pointwise_conv expects a float32 channels_last tensor")
    if residual is not None and not (residual.is_contiguous(memory_format=torch.channels_last) and residual.dtype == torch.float32
                                     and residual.shape == (b, cout, h, w)):
        raise ValueError("pointwise_conv: residual must be a float32 channels_last [B,Cout,H,W] tensor")
    if gate is not None and not (gate.dtype == torch.float32 and gate.shape == (b, c) and gate.is_contiguous()):
        raise ValueError("pointwise_conv: gate must be a contiguous float32 [B,Cin] tensor")
    pad_arg = None
    if out_pads is not None and any(out_pads):
        if residual is not None:
            raise ValueError("pointwise_conv: out_pads and residual are exclusive")
        pt, pb, pl, pr = (int(v) for v in out_pads)
        out = torch.empty((b, cout, h + pt + pb, w + pl + pr), dtype=torch.float32, device=x_cl.device,
                          memory_format=torch.channels_last)
        for sl in ((slice(0, pt), slice(None)), (slice(h + pt, None), slice(None)), (slice(None), slice(0, pl)),
                   (slice(None), slice(w + pl, None))):
            border = out[:, :, sl[0], sl[1]]
            if border.numel():
                border.zero_()
        pad_arg = (ctypes.c_int * 6)(h, w, pt, pb, pl, pr)
    else:
        out = torch.empty((b, cout, h, w), dtype=torch.float32, device=x_cl.device, memory_format=torch.channels_last)
    m = b * h * w
    with timing(name, work=2.0 * m * cout * c):
        rc = lib.mde_pointwise_x3_fwd(_p(x_cl), _p(gate), h * w, _p(w_pair), _p(bias), int(act), _p(residual), _p(out), m, cout,
                                      c, cout, cout, pad_arg, _s())
    _lib.check(rc, "mde_pointwise_x3_fwd")
    return out


def bias_act_pool_nhwc_(x_cl, bias, act):
    """In place x = act(x + bias[c]) on a channels_last tensor and, from the same pass, the per-slab channel sums
    [B, slabs, C] of the result (ops.se_gate turns them into the squeeze-excite gate)."""
    lib = _lib.load()
    b, c, h, w = x_cl.shape
    slabs = int(lib.mde_pool_slabs(b, h * w))
    partial = torch.empty((b, slabs, c), dtype=torch.float32, device=x_cl.device)
    rc = lib.mde_bias_act_pool_nhwc(_p(x_cl), _p(bias), _p(x_cl), _p(partial), b, h * w, c, int(act), _s())
    _lib.check(rc, "mde_bias_act_pool_nhwc")
    return x_cl, partial


def stem_conv_supported(x, cin, cout):
    return bool(x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == cin and cin % 4 == 0 and cout in (32, 48)
                and x.is_contiguous(memory_format=torch.channels_last))


def prepare_stem_weight(weight):
    """Conv filter [Cout,Cin,3,3] -> fp32 [9,Cin,Cout] (tap-major, C_out innermost: what stem_conv3x3s2 stages in shared memory)."""
    return weight.detach().float().permute(2, 3, 1, 0).reshape(9, weight.shape[1], weight.shape[0]).contiguous()


def stem_conv3x3s2(x_cl, w_tcc, bias, act, pad_top, pad_left, out_hw):
    """act(conv3x3 stride 2 (x) + bias) in exact fp32 on channels_last x [B,Cin,Hi,Wi]; zero padding pad_top / pad_left (and what
    out_hw implies at the bottom / right); w_tcc from prepare_stem_weight.  Returns channels_last [B,Cout,Ho,Wo]."""
    lib = _lib.load()
    _need_cuda(x_cl, w_tcc)
    b, cin, hi, wi = x_cl.shape
    cout = w_tcc.shape[2]
    ho, wo = (int(v) for v in out_hw)
    y = torch.empty((b, cout, ho, wo), dtype=torch.float32, device=x_cl.device, memory_format=torch.channels_last)
    with timing("stem_conv", work=2.0 * b * ho * wo * cout * 9 * cin):
        rc = lib.mde_stem_conv3x3s2_nhwc(_p(x_cl), _p(w_tcc), _p(bias), _p(y), b, hi, wi, cin, cout, int(pad_top), int(pad_left), ho,
                                         wo, int(act), _s())
    _lib.check(rc, "mde_stem_conv3x3s2_nhwc")
    return y


def depthwise_supported(x, channels, kernel_size, stride, dilation=(1, 1)):
    return bool(x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == channels and channels % 4 == 0
                and x.is_contiguous(memory_format=torch.channels_last) and tuple(kernel_size) in ((3, 3), (5, 5))
                and tuple(stride) in ((1, 1), (2, 2)) and tuple(dilation) == (1, 1))


def depthwise_bias_act_pool(x_cl, w_kkc, bias, act, stride, pad_top, pad_left, out_hw):
    """act(depthwise_conv(x) + bias) and, from the same pass, the per-slab channel sums of the result (for ops.se_gate).
    x_cl channels_last fp32 [B,C,Hi,Wi]; w_kkc fp32 [k,k,C] (the depthwise filter, channel innermost); zero padding pad_top /
    pad_left (and whatever ``out_hw`` implies at the bottom / right).  Returns (y channels_last [B,C,Ho,Wo], partial [B,slabs,C])."""
    lib = _lib.load()
    _need_cuda(x_cl, w_kkc)
    b, c, hi, wi = x_cl.shape
    ho, wo = (int(v) for v in out_hw)
    k = w_kkc.shape[0]
    if w_kkc.shape != (k, k, c) or w_kkc.dtype != torch.float32 or not w_kkc.is_contiguous():
        raise ValueError("depthwise_bias_act_pool: w_kkc must be a contiguous float32 [k,k,C] tensor")
    y = torch.empty((b, c, ho, wo), dtype=torch.float32, device=x_cl.device, memory_format=torch.channels_last)
    slabs = int(lib.mde_pool_slabs(b, ho * wo))
    partial = torch.empty((b, slabs, c), dtype=torch.float32, device=x_cl.device)
    with timing("depthwise", work=float(b * ho * wo * c * 8)):
        rc = lib.mde_depthwise_bias_act_pool_nhwc(_p(x_cl), _p(w_kkc), _p(bias), _p(y), _p(partial), b, hi, wi, c, k, int(stride),
                                                  int(pad_top), int(pad_left), ho, wo, int(act), _s())
    _lib.check(rc, "mde_depthwise_bias_act_pool_nhwc")
    return y, partial


def se_gate(partial, hw, w_reduce, b_reduce, w_expand, b_expand):
    """sigmoid(expand(silu(reduce(mean)))) of a squeeze-excite block from the slab sums of bias_act_pool_nhwc_:
    partial [B, slabs, C], w_reduce [R, C], w_expand [C, R] -> gate [B, C]."""
    lib = _lib.load()
    b, slabs, c = partial.shape
    r = w_reduce.shape[0]
    gate = torch.empty((b, c), dtype=torch.float32, device=partial.device)
    rc = lib.mde_se_gate(_p(partial), slabs, 1.0 / float(hw), _p(w_reduce), _p(b_reduce), _p(w_expand), _p(b_expand), _p(gate),
                         b, c, r, _s())
    _lib.check(rc, "mde_se_gate")
    return gate


# ------------------------------------------------------------------------------------------------------------
# transformer encoder layer
# ------------------------------------------------------------------------------------------------------------
def encoder_layer(x, layer, ws=None):
    """One nn.TransformerEncoderLayer (post-LN, ReLU, eval semantics) on tokens x [S, N, E]; ``layer`` is the torch
    module holding the parameters.  Returns a new [S, N, E] tensor."""
    lib = _lib.load()
    _need_cuda(x)
    x = _f32(x).contiguous()
    s, n, e = x.shape
    a = layer.self_attn
    ff = layer.linear1.out_features
    need = int(lib.mde_encoder_layer_ws_floats(s, n, e, ff))
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.float32, device=x.device)
    y = torch.empty_like(x)
    with timing("encoder_layer"):
        rc = lib.mde_encoder_layer_fwd(
            _p(x), _p(y), _p(a.in_proj_weight), _p(a.in_proj_bias), _p(a.out_proj.weight), _p(a.out_proj.bias),
            _p(layer.norm1.weight), _p(layer.norm1.bias), _p(layer.linear1.weight), _p(layer.linear1.bias),
            _p(layer.linear2.weight), _p(layer.linear2.bias), _p(layer.norm2.weight), _p(layer.norm2.bias), _p(ws),
            s, n, e, a.num_heads, ff, float(layer.norm1.eps), _s())
    _lib.check(rc, "mde_encoder_layer_fwd")
    return y, ws


def prepare_linear3(weight):
    """nn.Linear weight [N,K] -> [N,3K] = [w_hi | w_hi | w_lo] (w_hi = round_tf32(w), w_lo = round_tf32(w - w_hi)): the B
    operand of a 3xTF32 product against split-form activations [a | a_lo | a]."""
    w = _f32(weight.detach()).contiguous()
    hi = round_tf32(w)
    lo = round_tf32(w - hi)
    return torch.cat((hi, hi, lo), dim=1).contiguous()


def encoder_layers_tc(tokens, layers, prepared):
    """Run nn.TransformerEncoderLayer modules (post-LN, ReLU, eval semantics) on tokens [S,N,E] with the four linear
    products of each layer on the tcgen05 GEMM in 3xTF32 (fp32-grade accuracy).  ``prepared[i]`` = the four
    prepare_linear3() weights of layer i (in_proj, out_proj, linear1, linear2)."""
    lib = _lib.load()
    _need_cuda(tokens)
    x = _f32(tokens).contiguous()
    s, n, e = x.shape
    ff = layers[0].linear1.out_features
    ws = torch.empty(int(lib.mde_encoder_layer_tc_ws_floats(s, n, e, ff)), dtype=torch.float32, device=x.device)
    cur = torch.empty((s * n, 3 * e), dtype=torch.float32, device=x.device)
    nxt = torch.empty_like(cur)
    out = torch.empty_like(x)
    with timing("encoder_layers_tc"):
        _lib.check(lib.mde_split3_tf32(_p(x), _p(cur), s * n, e, _s()), "mde_split3_tf32")
        for i, (layer, (w_in, w_out, w_1, w_2)) in enumerate(zip(layers, prepared)):
            last = i == len(layers) - 1
            a = layer.self_attn
            rc = lib.mde_encoder_layer_tc_fwd(
                _p(cur), _p(out if last else nxt), 0 if last else 1, _p(w_in), _p(a.in_proj_bias), _p(w_out),
                _p(a.out_proj.bias), _p(layer.norm1.weight), _p(layer.norm1.bias), _p(w_1), _p(layer.linear1.bias), _p(w_2),
                _p(layer.linear2.bias), _p(layer.norm2.weight), _p(layer.norm2.bias), _p(ws), s, n, e, a.num_heads, ff,
                float(layer.norm1.eps), _s())
            _lib.check(rc, "mde_encoder_layer_tc_fwd")
            cur, nxt = nxt, cur
    return out


# ------------------------------------------------------------------------------------------------------------
# range attention / conv_out / bins
# ------------------------------------------------------------------------------------------------------------
def round_tf32(x, scale=1.0):
    lib = _lib.load()
    x = x.contiguous()
    out = torch.empty_like(x)
    _lib.check(lib.mde_round_tf32(_p(x), _p(out), x.numel(), float(scale), _s()), "mde_round_tf32")
    return out


def round_tf32_(x):
    """In place on a contiguous tensor."""
    lib = _lib.load()
    _lib.check(lib.mde_round_tf32(_p(x), _p(x), x.numel(), 1.0, _s()), "mde_round_tf32")
    return x


def range_attention(x, queries, impl="auto"):
    """y[b,n,h,w] = sum_k x[b,k,h,w] * queries[b,n,k]  (PixelWiseDotProduct).  impl: 'simt' (fp32 FMA),
    'tc' (TMA + tcgen05, three bf16 products per K step) or 'auto' (tc when the shape allows).  x may be a SplitBF16."""
    lib = _lib.load()
    _need_cuda(x, queries)
    queries = queries.contiguous().float()
    b, k, h, w = x.shape
    n = queries.shape[1]
    p = h * w
    tc_ok = (p % 128 == 0) and k == 128 and n == 128
    if impl == "auto":
        impl = "tc" if tc_ok else "simt"
    if impl == "tc" and not tc_ok:
        raise _lib.MdeError("tcgen05 range attention needs K = N = 128 and h*w % 128 == 0")
    y = torch.empty((b, n, h, w), dtype=torch.float32, device=queries.device)
    if impl == "tc":
        xp = split_bf16(x)
        qp = split_bf16_flat(queries)
        with timing("range_attention_tc"):
            rc = lib.mde_range_attention_tc(_p(xp.planes), _p(qp), _p(y), b, k, n, p, _s())
        _lib.check(rc, "mde_range_attention_tc")
        return y
    if isinstance(x, SplitBF16):
        x = x.float()
    x = x.contiguous().float()
    rc = lib.mde_range_attention(_p(x), _p(queries), _p(y), b, k, n, p, 0, _s())
    _lib.check(rc, "mde_range_attention")
    return y


def conv1x1(ram, weight, bias):
    """logits[b,j,h,w] = bias[j] + sum_n weight[j,n] ram[b,n,h,w]   (conv_out's Conv2d(128,n_bins,1), SIMT fp32)."""
    lib = _lib.load()
    ram = _f32(ram).contiguous()
    b, k, h, w = ram.shape
    wt = weight.reshape(weight.shape[0], -1).contiguous()
    out = torch.empty((b, wt.shape[0], h, w), dtype=torch.float32, device=ram.device)
    rc = lib.mde_conv1x1_fwd(_p(ram), _p(wt), _p(bias.contiguous()) if bias is not None else None, _p(out), b, k,
                             wt.shape[0], h * w, _s())
    _lib.check(rc, "mde_conv1x1_fwd")
    return out


def bins_pred(logits, centers):
    """pred[b,0,h,w] = sum_j softmax_j(logits[b,:,h,w]) * centers[b,j]  (streaming K2 kernel)."""
    lib = _lib.load()
    logits = _f32(logits).contiguous()
    b, n, h, w = logits.shape
    pred = torch.empty((b, 1, h, w), dtype=torch.float32, device=logits.device)
    with timing("bins_pred"):
        rc = lib.mde_bins_pred_fwd(_p(logits), _p(centers.contiguous()), _p(pred), b, n, h * w, _s())
    _lib.check(rc, "mde_bins_pred_fwd")
    return pred


def fold_queries_f32(w_out, bias, queries, feat_bias=None, scale=1.0, round_tf32=False):
    """wf[b] = scale * log2e * w_out @ queries[b]  [B,n_bins,K] fp32 (optionally TF32-rounded);  biasf [B,n_bins] = log2e *
    (bias + (w_out @ q[b]) @ feat_bias).  ``feat_bias`` is the bias of the conv that produced the chain's activations (folded
    in so that the producer can run bias-free)."""
    lib = _lib.load()
    wt = _f32(w_out).reshape(w_out.shape[0], -1).contiguous()
    n_bins, n = wt.shape
    queries = _f32(queries).contiguous()
    bias = _f32(bias)
    b, _, k = queries.shape
    wf = torch.empty((b, n_bins, k), dtype=torch.float32, device=queries.device)
    biasf = torch.empty((b, n_bins), dtype=torch.float32, device=queries.device)
    with timing("fold_queries"):
        rc = lib.mde_fold_queries(_p(wt), _p(bias.contiguous()), _p(queries), n * k,
                                  _p(feat_bias.contiguous()) if feat_bias is not None else None, _p(wf), _p(biasf), b,
                                  n_bins, n, k, float(scale), 1 if round_tf32 else 0, _s())
    _lib.check(rc, "mde_fold_queries")
    return wf, biasf


def fold_queries(w_out, bias, queries, feat_bias=None):
    """The per-image operand of the fused chain: (split-bf16 pair of wf = log2e * w_out @ queries[b], bfloat16
    [2,B,n_bins,K];  biasf [B,n_bins]) -- see fold_queries_f32."""
    wf, biasf = fold_queries_f32(w_out, bias, queries, feat_bias)
    return split_bf16_flat(wf), biasf


def head_chain(x, wf_pair, biasf, centers):
    """Fused range-attention -> conv_out -> softmax -> centre-weighted sum on tcgen05.  x: SplitBF16 [B,128,h,w] (what the
    conv3x3 epilogue writes), or an fp32 tensor in either memory format (split here, one extra pass); wf_pair / biasf from
    fold_queries -> pred [B,1,h,w]."""
    lib = _lib.load()
    x = split_bf16(x)
    biasf, centers = _f32(biasf), _f32(centers)
    if wf_pair.dtype != torch.bfloat16:
        wf_pair = split_bf16_flat(wf_pair)
    b, k, h, w = x.shape
    pred = torch.empty((b, 1, h, w), dtype=torch.float32, device=x.device)
    fn = lib.mde_head_chain_fwd if products() == 3 else lib.mde_head_chain_bf16_fwd
    with timing("head_chain"):
        rc = fn(_p(x.planes), _p(wf_pair), _p(biasf.contiguous()), _p(centers.contiguous()), _p(pred), b, wf_pair.shape[2], h * w,
                _s())
    _lib.check(rc, "mde_head_chain_fwd")
    return pred


def _is_nhwc(x):
    return x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()


def nhwc_to_nchw(x_cl):
    """channels_last [B,C,H,W] -> contiguous NCHW copy (the same tiled-transpose kernel, roles of C and P swapped)."""
    lib = _lib.load()
    b, c, h, w = x_cl.shape
    out = torch.empty((b, c, h, w), dtype=torch.float32, device=x_cl.device)
    _lib.check(lib.mde_nchw_to_nhwc(_p(x_cl), _p(out), b, h * w, c, _s()), "mde_nchw_to_nhwc")
    return out


class _HeadChainFn(torch.autograd.Function):
    """Training wrapper of the fused chain, forward and backward on the hand-written kernels:
    forward  = fold_queries + head_chain (training form: also stores the per-pixel softmax state, 8 B/px);
    backward = (1) the chain re-run with the backward epilogue: logits recomputed on the tensor cores, d loss / d logit
               written once in both layouts, d centres and the bias sums reduced in the epilogue;
               (2) d feat = gl W'  and (3) d W' = gl^T feat (split-K) on the tcgen05 NT GEMM;
               (4) the 256x128x128-sized products that unfold W' = W_out Q_b (a few MFLOP) in torch."""

    @staticmethod
    def forward(ctx, feat, queries, w_out, b_out, centers):
        lib = _lib.load()
        nhwc = _is_nhwc(feat)
        if not nhwc:
            feat = feat.contiguous()
        b, k, h, w = feat.shape
        fpair = split_bf16(feat)
        wf, biasf = fold_queries(w_out, b_out, queries)
        pred = torch.empty((b, 1, h, w), dtype=torch.float32, device=feat.device)
        stats = torch.empty((b, h * w, 2), dtype=torch.float32, device=feat.device)
        with timing("head_chain"):
            rc = lib.mde_head_chain_fwd_train(_p(fpair.planes), _p(wf), _p(biasf), _p(centers), _p(pred), _p(stats), b,
                                              wf.shape[2], h * w, _s())
        _lib.check(rc, "mde_head_chain_fwd_train")
        ctx.save_for_backward(feat, queries, w_out, b_out, centers, pred, stats, wf, biasf, fpair.planes)
        ctx.nhwc = nhwc
        return pred

    @staticmethod
    def backward(ctx, gpred):
        lib = _lib.load()
        feat, queries, w_out, b_out, centers, pred, stats, wf, biasf, fplanes = ctx.saved_tensors
        nhwc = ctx.nhwc
        b, k, h, w = feat.shape
        p = h * w
        nb = wf.shape[2]
        dev = feat.device
        gpred = gpred.contiguous().float()
        gl = torch.empty((b, p, nb), dtype=torch.float32, device=dev)
        glT = torch.empty((b, nb, p), dtype=torch.float32, device=dev)
        gc = torch.empty((b, nb), dtype=torch.float32, device=dev)
        gb_img = torch.empty((b, nb), dtype=torch.float32, device=dev)
        with timing("head_chain_bwd"):
            rc = lib.mde_head_chain_bwd_logits(_p(fplanes), _p(wf), _p(biasf), _p(centers), _p(pred), _p(stats), _p(gpred),
                                               _p(gl), _p(glT), _p(gc), _p(gb_img), b, nb, p, _s())
        _lib.check(rc, "mde_head_chain_bwd_logits")
        # W' = W_out Q_b without the log2(e) scale, TF32-rounded; its transpose is the K-major operand of d feat
        wplain, _ = fold_queries_f32(w_out, b_out, queries, scale=1.0 / LOG2E, round_tf32=True)
        wpt = wplain.transpose(1, 2).contiguous()                       # [B,128,256]
        if nhwc:
            gfeat = torch.empty((b, k, h, w), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
            gemm_nt(gl, wpt, out=gfeat.permute(0, 2, 3, 1).reshape(b, p, k))   # [B,P,128] = gl [B,P,256] . W'^T
            feat_t = nhwc_to_nchw(feat).reshape(b, k, p)
            feat_t = round_tf32_(feat_t)
        else:
            gfeat = torch.empty((b, k, h, w), dtype=torch.float32, device=dev)
            gemm_nt(wpt, gl, out=gfeat.reshape(b, k, p))                        # [B,128,P] = W'^T . gl^T
            feat_t = round_tf32(feat.reshape(b, k, p))
        # d W'[j,k] = sum_p gl[p,j] feat[p,k]: K = P pixels, split over CTAs; both operands are TF32-rounded (RNA), so the
        # tensor core's operand truncation is exact
        # two 128-row output tiles per image: enough K splits to put about two CTAs on every SM
        splits = max(1, min(64, (2 * NUM_SMS) // (2 * b)))
        gwp = gemm_nt(glT, feat_t, splits=splits)
        wo = w_out.reshape(w_out.shape[0], -1)
        gw = torch.einsum("bjk,bnk->jn", gwp, queries)                   # d W_out
        gq = torch.matmul(wo.t().unsqueeze(0), gwp)                      # d Q_b = W_out^T d W'_b
        return gfeat, gq, gw.view_as(w_out), gb_img.sum(dim=0), gc


def head_chain_autograd(feat, queries, w_out, b_out, centers):
    """pred = fused chain(feat, queries, conv_out, centres) with gradients to all five inputs; feat may be NCHW or
    channels_last (consumed in place either way)."""
    return _HeadChainFn.apply(_f32(feat), _f32(queries).contiguous(), _f32(w_out), _f32(b_out), _f32(centers).contiguous())


def head_chain_supported(x, n_bins):
    return x.shape[1] == 128 and n_bins == 256 and (x.shape[2] * x.shape[3]) % 128 == 0


class _UpsampleConcat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, skip):
        lib = _lib.load()
        b, c1, h, w = x.shape
        _, c2, hh, ww = skip.shape
        out = torch.empty((b, c1 + c2, hh, ww), dtype=torch.float32, device=x.device)
        with timing("upsample_concat"):
            rc = lib.mde_upsample_concat_fwd(_p(x), _p(skip), _p(out), b, c1, c2, h, w, hh, ww, _s())
        _lib.check(rc, "mde_upsample_concat_fwd")
        ctx.shape = (b, c1, c2, h, w, hh, ww)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        b, c1, c2, h, w, hh, ww = ctx.shape
        gout = _f32(gout).contiguous()
        gx = torch.empty((b, c1, h, w), dtype=torch.float32, device=gout.device)
        ws = torch.empty(int(lib.mde_upsample_bwd_ws_bytes(h, w)), dtype=torch.uint8, device=gout.device)
        with timing("upsample_bwd"):
            rc = lib.mde_upsample_bwd(_p(gout), _p(gx), 0, b, c1, c1 + c2, h, w, hh, ww, _p(ws), _s())
        _lib.check(rc, "mde_upsample_bwd")
        return gx, gout[:, c1:]


def upsample_concat(x, skip):
    """cat((bilinear_align_corners(x -> skip's size), skip), dim=1) in one pass (DecoderBN's UpSampleBN input)."""
    _need_cuda(x, skip)
    return _UpsampleConcat.apply(x.contiguous().float(), skip.contiguous().float())


class _UpsampleConcatNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_cl, skip):
        out = _upsample_concat_nhwc_fwd(x_cl, skip)
        ctx.shape = (x_cl.shape[0], x_cl.shape[1], skip.shape[1], x_cl.shape[2], x_cl.shape[3], skip.shape[2], skip.shape[3])
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        b, c1, c2, h, w, hh, ww = ctx.shape
        gout = _f32(gout)
        if not gout.is_contiguous(memory_format=torch.channels_last):
            gout = gout.contiguous(memory_format=torch.channels_last)
        gx = torch.empty((b, c1, h, w), dtype=torch.float32, device=gout.device, memory_format=torch.channels_last)
        ws = torch.empty(int(lib.mde_upsample_bwd_ws_bytes(h, w)), dtype=torch.uint8, device=gout.device)
        with timing("upsample_bwd"):
            rc = lib.mde_upsample_bwd(_p(gout), _p(gx), 1, b, c1, c1 + c2, h, w, hh, ww, _p(ws), _s())
        _lib.check(rc, "mde_upsample_bwd")
        return gx, gout[:, c1:]


def upsample_concat_nhwc(x_cl, skip):
    """channels_last DecoderBN up-sampling step: bilinear(align_corners=True) resize of x_cl [B,C1,h,w] (channels_last)
    to skip's size, concatenated with skip (either memory format) -> channels_last [B,C1+C2,H,W]; differentiable."""
    _need_cuda(x_cl, skip)
    x_cl, skip = _f32(x_cl), _f32(skip)
    if torch.is_grad_enabled() and (x_cl.requires_grad or skip.requires_grad):
        return _UpsampleConcatNHWC.apply(x_cl, skip)
    return _upsample_concat_nhwc_fwd(x_cl, skip)


def upsample_concat_nhwc_pair(x_cl, skip, pad_to=1):
    """The inference form of upsample_concat_nhwc that writes its result as a SplitBF16 (the operand format of the conv3x3
    that follows): bilinear(align_corners=True) resize of x_cl to skip's size, concatenated with skip.  ``pad_to``: the channel
    count is rounded up to a multiple of it with zero channels (32 keeps the conv's 64-byte operand rows sector-aligned; the
    conv's filter is zero-padded to match, see prepare_conv3x3_weight(cin_pad_to=...))."""
    lib = _lib.load()
    _need_cuda(x_cl, skip)
    x_cl, skip = _f32(x_cl), _f32(skip)
    if not x_cl.is_contiguous(memory_format=torch.channels_last):
        x_cl = x_cl.contiguous(memory_format=torch.channels_last)
    if not skip.is_contiguous(memory_format=torch.channels_last):
        skip = skip.contiguous(memory_format=torch.channels_last)
    b, c1, h, w = x_cl.shape
    _, c2, hh, ww = skip.shape
    cp = -(-(c1 + c2) // pad_to) * pad_to
    planes = torch.empty((2, b, hh, ww, cp), dtype=torch.bfloat16, device=x_cl.device)
    with timing("upsample_concat_nhwc"):
        rc = lib.mde_upsample_concat_nhwc_pair_fwd(_p(x_cl), _p(skip), _p(planes), b, c1, c2, cp, h, w, hh, ww, _s())
    _lib.check(rc, "mde_upsample_concat_nhwc_pair_fwd")
    return SplitBF16(planes)


def _upsample_concat_nhwc_fwd(x_cl, skip):
    lib = _lib.load()
    if not x_cl.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("upsample_concat_nhwc expects a channels_last x")
    b, c1, h, w = x_cl.shape
    _, c2, hh, ww = skip.shape
    skip_cl = skip.is_contiguous(memory_format=torch.channels_last)
    if not skip_cl:
        skip = skip.contiguous()
    out = torch.empty((b, c1 + c2, hh, ww), dtype=torch.float32, device=x_cl.device, memory_format=torch.channels_last)
    with timing("upsample_concat_nhwc"):
        rc = lib.mde_upsample_concat_nhwc_fwd(_p(x_cl), _p(skip), 1 if skip_cl else 0, _p(out), b, c1, c2, h, w, hh, ww, _s())
    _lib.check(rc, "mde_upsample_concat_nhwc_fwd")
    return out


def to_channels_last(x):
    """NCHW-contiguous fp32 [B,C,H,W] -> the same logical tensor with channels_last strides (tiled transpose kernel)."""
    lib = _lib.load()
    _need_cuda(x)
    x = _f32(x)
    if x.is_contiguous(memory_format=torch.channels_last):
        return x
    x = x.contiguous()
    b, c, h, w = x.shape
    out = torch.empty_like(x, memory_format=torch.channels_last)
    with timing("nchw_to_nhwc"):
        rc = lib.mde_nchw_to_nhwc(_p(x), _p(out), b, c, h * w, _s())
    _lib.check(rc, "mde_nchw_to_nhwc")
    return out


def concat_channels_last(sources, pads=None):
    """torch.cat(sources, dim=1) of NCHW-contiguous fp32 tensors, produced directly in channels_last memory (one tiled
    transpose per source into its channel slice; no planar intermediate).  pads = (top, bottom, left, right): the result is
    additionally zero-padded spatially (the SAME padding of a stride-2 stem convolution)."""
    lib = _lib.load()
    _need_cuda(*sources)
    sources = [_f32(t).contiguous() for t in sources]
    b, _, h, w = sources[0].shape
    ctot = sum(t.shape[1] for t in sources)
    pt, pb, pl, pr = (0, 0, 0, 0) if pads is None else (int(v) for v in pads)
    out = torch.empty((b, ctot, h + pt + pb, w + pl + pr), dtype=torch.float32, device=sources[0].device,
                      memory_format=torch.channels_last)
    base = out.data_ptr()
    ch = 0
    with timing("nchw_to_nhwc"):
        for t in sources:
            dst = ctypes.c_void_p(base + 4 * ch)
            if pt or pb or pl or pr:
                rc = lib.mde_nchw_to_nhwc_slice_padded(_p(t), dst, b, t.shape[1], h, w, ctot, pt, pb, pl, pr, _s())
            else:
                rc = lib.mde_nchw_to_nhwc_slice(_p(t), dst, b, t.shape[1], h * w, ctot, _s())
            _lib.check(rc, "mde_nchw_to_nhwc_slice")
            ch += t.shape[1]
    if pt:
        out[:, :, :pt].zero_()
    if pb:
        out[:, :, h + pt:].zero_()
    if pl:
        out[:, :, :, :pl].zero_()
    if pr:
        out[:, :, :, w + pl:].zero_()
    return out


def relu_eps(x, eps=1e-4):
    lib = _lib.load()
    x = _f32(x).contiguous()
    y = torch.empty_like(x)
    _lib.check(lib.mde_relu_eps_fwd(_p(x), _p(y), x.numel(), float(eps), _s()), "mde_relu_eps_fwd")
    return y


# ------------------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------------------
class _SILog(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mask, interpolate):
        lib = _lib.load()
        b, _, h, w = pred.shape
        hh, ww = target.shape[-2:]
        ws = torch.empty(int(lib.mde_silog_ws_bytes()), dtype=torch.uint8, device=pred.device)
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        with timing("silog_fwd"):
            rc = lib.mde_silog_fwd(_p(pred), _p(target), _p(mask), b, h, w, hh, ww, 1 if interpolate else 0, _p(ws),
                                   _p(loss), _s())
        _lib.check(rc, "mde_silog_fwd")
        ctx.save_for_backward(pred, target, mask, ws)
        ctx.interpolate = interpolate
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        pred, target, mask, ws = ctx.saved_tensors
        b, _, h, w = pred.shape
        hh, ww = target.shape[-2:]
        gp = torch.empty_like(pred)
        g = g.contiguous().float()
        rc = lib.mde_silog_bwd(_p(pred), _p(target), _p(mask), b, h, w, hh, ww, 1 if ctx.interpolate else 0, _p(ws),
                               _p(g), _p(gp), _s())
        _lib.check(rc, "mde_silog_bwd")
        return gp, None, None, None


def silog(pred, target, mask=None, interpolate=True):
    _need_cuda(pred, target, mask)
    if pred.dim() != 4 or pred.shape[1] != 1 or target.dim() != 4 or target.shape[1] != 1:
        raise ValueError("silog expects pred [B,1,h,w] and target [B,1,H,W]")
    if mask is not None:
        if mask.dtype != torch.bool and mask.dtype != torch.uint8:
            raise ValueError("mask must be a bool tensor")
        mask = mask.expand_as(target).contiguous()
        mask = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
    return _SILog.apply(pred.contiguous().float(), target.contiguous().float(), mask, bool(interpolate))


class _Chamfer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, edges, target, min_target):
        lib = _lib.load()
        b, n1 = edges.shape
        hw = target.numel() // b
        want_grad = bool(ctx.needs_input_grad[0])  # per-centre sums are only collected for the backward
        ws = torch.empty(int(lib.mde_chamfer_ws_bytes(b, n1 - 1)), dtype=torch.uint8, device=edges.device)
        loss = torch.empty((), dtype=torch.float32, device=edges.device)
        with timing("chamfer_fwd"):
            rc = lib.mde_chamfer_fwd(_p(edges), _p(target), b, n1 - 1, hw, float(min_target), 1 if want_grad else 0, _p(ws),
                                     _p(loss), _s())
        _lib.check(rc, "mde_chamfer_fwd")
        ctx.save_for_backward(edges, ws)
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        edges, ws = ctx.saved_tensors
        b, n1 = edges.shape
        ge = torch.empty_like(edges)
        g = g.contiguous().float()
        _lib.check(lib.mde_chamfer_bwd(_p(edges), b, n1 - 1, _p(ws), _p(g), _p(ge), _s()), "mde_chamfer_bwd")
        return ge, None, None


def bins_chamfer(edges, target_depth_maps, min_target=1e-3):
    _need_cuda(edges, target_depth_maps)
    if edges.dim() != 2 or target_depth_maps.shape[0] != edges.shape[0]:
        raise ValueError("bins_chamfer expects edges [B,n_bins+1] and targets [B,...]")
    return _Chamfer.apply(edges.contiguous().float(), target_depth_maps.contiguous().float(), min_target)


class _DepthLosses(torch.autograd.Function):
    """SILog (mask = target > min_depth) and bins-chamfer in one pass over the target (mde_depth_losses_fwd)."""

    @staticmethod
    def forward(ctx, pred, edges, target, min_depth, min_target, interpolate):
        lib = _lib.load()
        b, _, h, w = pred.shape
        hh, ww = target.shape[-2:]
        n1 = edges.shape[1]
        want_grad = bool(ctx.needs_input_grad[1])
        ws_s = torch.empty(int(lib.mde_silog_ws_bytes()), dtype=torch.uint8, device=pred.device)
        ws_c = torch.empty(int(lib.mde_chamfer_ws_bytes(b, n1 - 1)), dtype=torch.uint8, device=pred.device)
        out = torch.empty(2, dtype=torch.float32, device=pred.device)
        with timing("loss_fused"):
            rc = lib.mde_depth_losses_fwd(_p(pred), _p(edges), _p(target), b, h, w, hh, ww, n1 - 1, 1 if interpolate else 0,
                                          float(min_depth), float(min_target), 1 if want_grad else 0, _p(ws_s), _p(ws_c),
                                          _p(out[0:1]), _p(out[1:2]), _s())
        _lib.check(rc, "mde_depth_losses_fwd")
        ctx.save_for_backward(pred, edges, target, ws_s, ws_c)
        ctx.cfg = (float(min_depth), bool(interpolate))
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_silog, g_chamfer):
        lib = _lib.load()
        pred, edges, target, ws_s, ws_c = ctx.saved_tensors
        min_depth, interpolate = ctx.cfg
        b, _, h, w = pred.shape
        hh, ww = target.shape[-2:]
        n1 = edges.shape[1]
        gp = ge = None
        if ctx.needs_input_grad[0]:
            gp = torch.empty_like(pred)
            rc = lib.mde_silog_bwd_thr(_p(pred), _p(target), min_depth, b, h, w, hh, ww, 1 if interpolate else 0, _p(ws_s),
                                       _p(g_silog.contiguous().float()), _p(gp), _s())
            _lib.check(rc, "mde_silog_bwd_thr")
        if ctx.needs_input_grad[1]:
            ge = torch.empty_like(edges)
            rc = lib.mde_chamfer_bwd(_p(edges), b, n1 - 1, _p(ws_c), _p(g_chamfer.contiguous().float()), _p(ge), _s())
            _lib.check(rc, "mde_chamfer_bwd")
        return gp, ge, None, None, None, None


def depth_losses(pred, edges, target, min_depth=1e-3, min_target=1e-3, interpolate=True):
    """(SILogLoss(pred, target, mask=target > min_depth, interpolate), BinsChamferLoss(edges, target)) -- the two criteria of
    the reference's training loop (train.py:414-419) -- from ONE kernel that reads the target once."""
    _need_cuda(pred, edges, target)
    if pred.dim() != 4 or pred.shape[1] != 1 or target.dim() != 4 or target.shape[1] != 1 or edges.dim() != 2:
        raise ValueError("depth_losses expects pred [B,1,h,w], edges [B,n_bins+1] and target [B,1,H,W]")
    return _DepthLosses.apply(pred.contiguous().float(), edges.contiguous().float(), target.contiguous().float(), min_depth,
                              min_target, bool(interpolate))


# ------------------------------------------------------------------------------------------------------------
# "next" row (f)2: evaluation epilogue + metrics
# ------------------------------------------------------------------------------------------------------------
METRIC_KEYS = ("a1", "a2", "a3", "abs_rel", "rmse", "log_10", "rmse_log", "silog", "sq_rel")


def eval_metrics(pred, gt, min_depth_eval, max_depth_eval, crop_box=None):
    """pred [B,1,h,w], gt [B,1,H,W] (CUDA float32) -> [B,10] float32: the reference's nine metrics per image
    (METRIC_KEYS order) + the number of valid pixels.  The bilinear up-sampling, the clipping of the prediction and the
    validity / crop mask of evaluate.py:59-71,128-150 are fused into the reduction.  crop_box = (y0, y1, x0, x1)."""
    lib = _lib.load()
    _need_cuda(pred, gt)
    if pred.dim() != 4 or gt.dim() != 4 or pred.shape[1] != 1 or gt.shape[1] != 1 or pred.shape[0] != gt.shape[0]:
        raise ValueError("eval_metrics expects pred [B,1,h,w] and gt [B,1,H,W]")
    pred, gt = pred.contiguous().float(), gt.contiguous().float()
    b, _, h, w = pred.shape
    hh, ww = gt.shape[-2:]
    y0, y1, x0, x1 = (0, hh, 0, ww) if crop_box is None else crop_box
    ws = torch.empty(int(lib.mde_eval_metrics_ws_bytes(b)), dtype=torch.uint8, device=pred.device)
    out = torch.empty((b, 10), dtype=torch.float32, device=pred.device)
    with timing("eval_metrics"):
        rc = lib.mde_eval_metrics_fwd(_p(pred), _p(gt), b, h, w, hh, ww, float(min_depth_eval), float(max_depth_eval),
                                      int(y0), int(y1), int(x0), int(x1), _p(ws), _p(out), _s())
    _lib.check(rc, "mde_eval_metrics_fwd")
    return out


def flip_average(pred, pred_of_flipped, lo, hi):
    """0.5 * (clip(pred) + clip(flip_w(pred_of_flipped)))  -- the mirror test-time augmentation of infer.py:108-118."""
    lib = _lib.load()
    _need_cuda(pred, pred_of_flipped)
    a, bf = pred.contiguous().float(), pred_of_flipped.contiguous().float()
    out = torch.empty_like(a)
    w = a.shape[-1]
    _lib.check(lib.mde_flip_average(_p(a), _p(bf), _p(out), a.numel() // w, w, float(lo), float(hi), _s()), "mde_flip_average")
    return out
