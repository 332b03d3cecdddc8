"""Tensor-level operators over the C ABI (libdad_sm100a.so) + their torch.library registration.

Functional layer (no autograd): augment, encoder_pool_fwd, encoder_pool_bwd, ema_, split_f16.
torch.library layer: `torch.ops.dad.encoder_pool` (autograd-aware; used by model.Emotion2VecEncoder so the
reference's module API -- enc(x, padding_mask) -> [B,256], differentiable w.r.t. W, b -- stays drop-in) and
`torch.ops.dad.augment`.  Nothing here has a CPU or eager-PyTorch fallback: CPU tensors raise DadError."""
import ctypes as C

import torch

from . import _lib
from ._lib import D_HID, D_IN, GEMM_FP32, GEMM_TC, NPARAM, DadError, check, lib, make_aug, ptr, stream_ptr

_ws_cache = {}


def _zeros_ws(nbytes, device, tag):
    """Zero-initialised, cached workspace (the ops keep their arrival counters consistent themselves)."""
    key = (tag, device.index, torch.cuda.current_stream().cuda_stream)
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = t
    return t


def default_gemm_mode(T):
    import os
    mode = os.environ.get("DAD_GEMM", "").lower()
    if mode == "fp32":
        return GEMM_FP32
    if mode in ("tc", "f16x3"):
        return GEMM_TC
    # tcgen05 path by default (fp32-grade 3-product fp16 split); utterances shorter than one 32-frame group use the CUDA-core kernel
    return GEMM_TC if T >= 32 else GEMM_FP32


def _as_u8_pad(pad, B, T, device):
    if pad is None:
        return None
    if pad.dtype == torch.bool:
        pad = pad.view(torch.uint8) if pad.is_contiguous() else pad.contiguous().view(torch.uint8)
    elif pad.dtype != torch.uint8:
        pad = (pad != 0).to(torch.uint8)
    if tuple(pad.shape) != (B, T):
        raise DadError(f"padding_mask shape {tuple(pad.shape)} != {(B, T)}")
    return pad.contiguous()


def augment(x, sigma, noise=None, seed=None, stream=0, feat_keep=None, t_start=None, mask_len=0, strong=False):
    """DataAugmentation.weak_augment / strong_augment on [B,T,768] or [T,768] (I/utils.py:328-375).  Materialises the view."""
    _lib.require_device()
    x3 = x if x.dim() == 3 else x.unsqueeze(0)
    if x3.dim() != 3 or x3.shape[-1] != D_IN:
        raise DadError(f"augment: expected [B,T,{D_IN}] or [T,{D_IN}], got {tuple(x.shape)}")
    x3 = x3.contiguous()
    B, T, _ = x3.shape
    out = torch.empty_like(x3)
    if noise is not None:
        noise = noise.reshape(B, T, D_IN).contiguous()
    a = make_aug(2 if strong else 1, noise=noise, sigma=sigma, feat_keep=feat_keep, t_start=t_start, mask_len=mask_len,
                 seed=seed, stream=stream)
    check(lib().dad_augment(ptr(x3, torch.float32, "x"), ptr(out), C.byref(a), B, T, stream_ptr()))
    return out if x.dim() == 3 else out[0]


def encoder_pool_fwd(x, pad, W, b, aug=None, gemm_mode=None, want_bits=True, W_hi=None, W_lo=None, emb_out=None, xs=None):
    """emb[B,256] = masked-mean_t relu(W aug(x) + b)  (I/model.py:18-41).  Returns (emb, relu_bits|None, lens).
    xs = (xs_hi,) or (xs_hi, xs_lo): optional fp16 [B*T,768] operand-cache buffers filled for encoder_pool_bwd (tcgen05 mode only;
    the default 2-product backward needs xs_hi alone)."""
    _lib.require_device()
    if x.dim() != 3 or x.shape[-1] != D_IN:
        raise DadError(f"encoder: expected x [B,T,{D_IN}], got {tuple(x.shape)}")
    x = x.contiguous()
    B, T, _ = x.shape
    dev = x.device
    pad8 = _as_u8_pad(pad, B, T, dev)
    mode = default_gemm_mode(T) if gemm_mode is None else gemm_mode
    emb = emb_out if emb_out is not None else torch.empty(B, D_HID, dtype=torch.float32, device=dev)
    bits = torch.empty(B * T, D_HID // 32, dtype=torch.int32, device=dev) if want_bits else None
    lens = torch.empty(B, dtype=torch.float32, device=dev)
    if B == 0:
        return emb, bits, lens
    if mode == GEMM_TC and (W_hi is None or W_lo is None):
        W_hi, W_lo = split_f16(W)
    nb = lib().dad_encoder_fwd_workspace_bytes(B, T, mode)
    ws = _zeros_ws(nb, dev, "enc_fwd")
    check(lib().dad_encoder_fwd(ptr(x, torch.float32, "x"), ptr(pad8), ptr(W.contiguous(), torch.float32, "W"),
                                ptr(b.contiguous(), torch.float32, "b"), ptr(W_hi), ptr(W_lo),
                                C.byref(aug) if aug is not None else None, ptr(emb), ptr(bits),
                                _xs(xs, 0), _xs(xs, 1), ptr(lens), ptr(ws), ws.numel(), B, T, mode, stream_ptr()))
    return emb, bits, lens


def _xs(xs, i):
    """i-th buffer of an operand cache given as (xs_hi,) or (xs_hi, xs_lo); None -> NULL."""
    if xs is None or len(xs) <= i or xs[i] is None:
        return None
    return ptr(xs[i], torch.float16, "xs_hi" if i == 0 else "xs_lo")


def aug_encoder_fwd(x, pad, Wt, bt, Ws, bs, aug_weak, aug_strong, want_bits=True, xs=None, Wt_split=None, Ws_split=None):
    """Fused noisy forward (I/train.py:405-409,439): ONE read of x -> teacher(weak(x)) and student(strong(x)) on the tcgen05 path.
    Returns (emb_teacher, emb_student, relu_bits_student|None, lens)."""
    _lib.require_device()
    if x.dim() != 3 or x.shape[-1] != D_IN:
        raise DadError(f"encoder: expected x [B,T,{D_IN}], got {tuple(x.shape)}")
    x = x.contiguous()
    B, T, _ = x.shape
    dev = x.device
    pad8 = _as_u8_pad(pad, B, T, dev)
    emb_t = torch.empty(B, D_HID, dtype=torch.float32, device=dev)
    emb_s = torch.empty(B, D_HID, dtype=torch.float32, device=dev)
    bits = torch.empty(B * T, D_HID // 32, dtype=torch.int32, device=dev) if want_bits else None
    lens = torch.empty(B, dtype=torch.float32, device=dev)
    if B == 0:
        return emb_t, emb_s, bits, lens
    t_hi, t_lo = Wt_split if Wt_split is not None else split_f16(Wt)
    s_hi, s_lo = Ws_split if Ws_split is not None else split_f16(Ws)
    ws = _zeros_ws(lib().dad_aug_encoder_fwd_workspace_bytes(B, T), dev, "aug_enc_fwd")
    check(lib().dad_aug_encoder_fwd(ptr(x, torch.float32, "x"), ptr(pad8), ptr(t_hi), ptr(t_lo), ptr(bt.contiguous(), torch.float32, "bt"),
                                    C.byref(aug_weak), ptr(emb_t), ptr(s_hi), ptr(s_lo), ptr(bs.contiguous(), torch.float32, "bs"),
                                    C.byref(aug_strong), ptr(emb_s), ptr(bits), _xs(xs, 0), _xs(xs, 1), ptr(lens), ptr(ws), ws.numel(),
                                    B, T, stream_ptr()))
    return emb_t, emb_s, bits, lens


def encoder_pool_bwd(x, pad, bits, g_emb, lens, aug=None, gemm_mode=None, xs=None):
    """(dW[256,768], db[256]) of encoder_pool_fwd given dL/d emb (autograd of I/model.py:28-39; no dX is needed).
    xs = the operand cache filled by encoder_pool_fwd (then x / aug are not read)."""
    _lib.require_device()
    x = x.contiguous()
    B, T, _ = x.shape
    dev = x.device
    pad8 = _as_u8_pad(pad, B, T, dev)
    mode = default_gemm_mode(T) if gemm_mode is None else gemm_mode
    S = lib().dad_encoder_bwd_splits(B, T, mode)
    dWp = torch.empty(S, D_HID, D_IN, dtype=torch.float32, device=dev)
    dbp = torch.empty(S, D_HID, dtype=torch.float32, device=dev)
    nb = lib().dad_encoder_bwd_workspace_bytes(B, T, mode)
    ws = _zeros_ws(nb, dev, "enc_bwd")
    check(lib().dad_encoder_bwd(ptr(x, torch.float32, "x"), ptr(pad8), C.byref(aug) if aug is not None else None,
                                ptr(bits), _xs(xs, 0), _xs(xs, 1),
                                ptr(g_emb.contiguous(), torch.float32, "g_emb"), ptr(lens, torch.float32, "lens"),
                                ptr(dWp), ptr(dbp), S, ptr(ws), ws.numel(), B, T, mode, stream_ptr()))
    g = torch.empty(NPARAM, dtype=torch.float32, device=dev)
    zc = torch.zeros(4 * D_HID + 4, dtype=torch.float32, device=dev)
    check(lib().dad_grad_reduce(ptr(dWp), ptr(dbp), S, None, None, 0, ptr(zc), zc.data_ptr() + 4 * D_HID * 4, ptr(g), stream_ptr()))
    return g[:_lib.OFF_B].view(D_HID, D_IN), g[_lib.OFF_B:_lib.OFF_WC]


def predict(x, pad, Ws, bs, Wcs, bcs, Wt=None, bt=None, Wct=None, bct=None, Ws_split=None, Wt_split=None):
    """Fused inference (dad_predict): ONE read of x [B,T,768] on the tcgen05 path -> student (and, with Wt given, teacher) encoder +
    classifier in eval mode.  Returns dict(logits_s, pred_s, emb_s[, logits_t, pred_t, emb_t]).  T >= 32."""
    _lib.require_device()
    if x.dim() != 3 or x.shape[-1] != D_IN:
        raise DadError(f"predict: expected x [B,T,{D_IN}], got {tuple(x.shape)}")
    x = x.contiguous().float()
    B, T, _ = x.shape
    dev = x.device
    two = Wt is not None
    pad8 = _as_u8_pad(pad, B, T, dev)
    f32 = lambda *sh: torch.empty(*sh, dtype=torch.float32, device=dev)
    i32 = lambda *sh: torch.empty(*sh, dtype=torch.int32, device=dev)
    out = {"logits_s": f32(B, _lib.N_CLS), "pred_s": i32(B), "emb_s": f32(B, D_HID)}
    if two:
        out.update({"logits_t": f32(B, _lib.N_CLS), "pred_t": i32(B), "emb_t": f32(B, D_HID)})
    if B == 0:
        return out
    s_hi, s_lo = Ws_split if Ws_split is not None else split_f16(Ws)
    t_hi, t_lo = (Wt_split if Wt_split is not None else split_f16(Wt)) if two else (None, None)
    ws = _zeros_ws(lib().dad_predict_workspace_bytes(B, T), dev, "predict")
    c = lambda t: ptr(t.contiguous().float(), torch.float32) if t is not None else None
    check(lib().dad_predict(ptr(x, torch.float32, "x"), ptr(pad8), ptr(s_hi), ptr(s_lo), c(bs), c(Wcs), c(bcs), ptr(t_hi), ptr(t_lo), c(bt),
                            c(Wct), c(bct), ptr(out["emb_s"]), ptr(out.get("emb_t")), ptr(out["logits_s"]), ptr(out.get("logits_t")),
                            ptr(out["pred_s"]), ptr(out.get("pred_t")), ptr(ws), ws.numel(), B, T, stream_ptr()))
    return out


def gather_pad(feats, offsets, sizes, T=None, x_out=None, pad_out=None):
    """Device-side collate (I/dataload_clean.py:177-193): utterances (offsets, sizes) of a resident [sum T, 768] feature array ->
    zero-padded x [B,T,768] and padding mask [B,T] (bool, True = pad).  offsets int64, sizes int32, all on the device."""
    _lib.require_device()
    B = int(offsets.numel())
    if T is None:
        T = int(sizes.max()) if B else 0
    x = x_out if x_out is not None else torch.empty(B, T, D_IN, dtype=torch.float32, device=feats.device)
    pad = pad_out if pad_out is not None else torch.empty(B, T, dtype=torch.uint8, device=feats.device)
    check(lib().dad_gather_pad(ptr(feats, torch.float32, "feats"), ptr(offsets, torch.int64, "offsets"), ptr(sizes, torch.int32, "sizes"),
                               ptr(x, torch.float32, "x"), ptr(pad, torch.uint8, "pad"), B, T, stream_ptr()))
    return x, pad.view(torch.bool)


def ema_(teacher_flat, student_flat, momentum):
    """In-place t = t*m + s*(1-m) with the reference's rounding (I/model.py:211-223)."""
    _lib.require_device()
    m32, om32 = float(momentum), float(1.0 - momentum)
    check(lib().dad_ema(ptr(teacher_flat, torch.float32, "teacher"), ptr(student_flat, torch.float32, "student"), m32, om32,
                        teacher_flat.numel(), stream_ptr()))
    return teacher_flat


def split_f16(W):
    """(hi, lo) fp16 tensors with hi = fp16(256 W), lo = fp16(256 W - hi): operands of the 3-product tcgen05 GEMM."""
    _lib.require_device()
    W = W.contiguous()
    hi = torch.empty(W.shape, dtype=torch.float16, device=W.device)
    lo = torch.empty_like(hi)
    check(lib().dad_split_f16(ptr(W, torch.float32, "W"), ptr(hi), ptr(lo), W.numel(), stream_ptr()))
    return hi, lo


# ------------------------------------------------------------------------------------------------ torch.library
_registered = False


def register_torch_ops():
    """Register torch.ops.dad.* once (custom_op + fake + autograd).  Safe to call repeatedly."""
    global _registered
    if _registered:
        return
    _registered = True

    @torch.library.custom_op("dad::encoder_pool", mutates_args=(), device_types="cuda")
    def encoder_pool(x: torch.Tensor, pad: torch.Tensor, W: torch.Tensor, b: torch.Tensor,
                     has_pad: bool) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        emb, bits, lens = encoder_pool_fwd(x, pad if has_pad else None, W, b, want_bits=True)
        return emb, bits, lens

    @encoder_pool.register_fake
    def _(x, pad, W, b, has_pad):
        B, T = x.shape[0], x.shape[1]
        return (x.new_empty(B, D_HID), x.new_empty(B * T, D_HID // 32, dtype=torch.int32), x.new_empty(B))

    @torch.library.custom_op("dad::encoder_pool_bwd", mutates_args=(), device_types="cuda")
    def encoder_pool_bwd_op(x: torch.Tensor, pad: torch.Tensor, bits: torch.Tensor, g_emb: torch.Tensor, lens: torch.Tensor,
                            has_pad: bool) -> tuple[torch.Tensor, torch.Tensor]:
        dW, db = encoder_pool_bwd(x, pad if has_pad else None, bits, g_emb, lens)
        return dW.clone(), db.clone()

    @encoder_pool_bwd_op.register_fake
    def _(x, pad, bits, g_emb, lens, has_pad):
        return x.new_empty(D_HID, D_IN), x.new_empty(D_HID)

    def setup_context(ctx, inputs, output):
        x, pad, W, b, has_pad = inputs
        _, bits, lens = output
        ctx.save_for_backward(x, pad, bits, lens)
        ctx.has_pad = has_pad

    def backward(ctx, g_emb, _g_bits, _g_lens):
        x, pad, bits, lens = ctx.saved_tensors
        dW, db = torch.ops.dad.encoder_pool_bwd(x, pad, bits, g_emb.contiguous(), lens, ctx.has_pad)
        return None, None, dW, db, None  # inputs never need a gradient on this path (I/train.py: feats are data)

    encoder_pool.register_autograd(backward, setup_context=setup_context)

    @torch.library.custom_op("dad::augment", mutates_args=(), device_types="cuda")
    def augment_op(x: torch.Tensor, noise: torch.Tensor, sigma: float, feat_keep: torch.Tensor, t_start: torch.Tensor,
                   mask_len: int, strong: bool) -> torch.Tensor:
        return augment(x, sigma, noise=noise, feat_keep=feat_keep if strong else None, t_start=t_start if strong else None,
                       mask_len=mask_len, strong=strong)

    @augment_op.register_fake
    def _(x, noise, sigma, feat_keep, t_start, mask_len, strong):
        return torch.empty_like(x)
