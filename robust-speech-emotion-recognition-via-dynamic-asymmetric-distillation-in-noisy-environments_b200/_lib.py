"""ctypes binding of libdad_sm100a.so (C ABI: /include/dad_sm100a.h).

There is no fallback: if the library is missing or the device is not sm_100, every entry point raises.
Wrappers take torch CUDA tensors, check dtype/contiguity host-side, and enqueue on the CURRENT torch stream."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdad_sm100a.so")

D_IN, D_HID, N_CLS = 768, 256, 4
NPARAM = 197892
OFF_W, OFF_B, OFF_WC, OFF_BC = 0, 196608, 196864, 197888
GEMM_FP32, GEMM_TC = 0, 1

EXPORTS = (
    "dad_abi_version", "dad_last_error", "dad_device_check", "dad_augment", "dad_encoder_fwd_workspace_bytes",
    "dad_encoder_fwd", "dad_encoder_bwd_splits", "dad_encoder_bwd", "dad_encoder_bwd_workspace_bytes",
    "dad_pack_rank_words", "dad_head_grid", "dad_head_fwd", "dad_dacp_mask", "dad_head_bwd",
    "dad_ecda_workspace_bytes", "dad_ecda", "dad_grad_reduce", "dad_grad_norm", "dad_adam_ema", "dad_ema",
    "dad_split_f16", "dad_step_rng", "dad_certainty_scores", "dad_scl_workspace_bytes", "dad_scl",
    "dad_peer_push", "dad_peer_barrier", "dad_grad_sum_ranks", "dad_max_world", "dad_gather_pad",
    "dad_aug_encoder_fwd", "dad_aug_encoder_fwd_workspace_bytes", "dad_tc_set_cta_pair", "dad_tc_get_cta_pair",
    "dad_tc_set_bwd_products", "dad_tc_get_bwd_products", "dad_predict", "dad_predict_workspace_bytes",
    "dad_peer_push_barrier", "dad_grad_reduce_xchg", "dad_grad_norm_xchg",
)


class DadError(RuntimeError):
    pass


class Aug(C.Structure):
    _fields_ = [("mode", C.c_int32), ("noise_src", C.c_int32), ("noise", C.c_void_p), ("seed", C.c_uint64),
                ("stream", C.c_uint64), ("sigma", C.c_float), ("feat_keep", C.c_void_p), ("t_start", C.c_void_p),
                ("mask_len", C.c_int32), ("step_ptr", C.c_void_p), ("frame_offset", C.c_int64)]


class Pack(C.Structure):
    _fields_ = [("base", C.c_void_p), ("B_local", C.c_int32), ("world", C.c_int32)]


class HeadFwdArgs(C.Structure):
    _fields_ = [("pack", Pack), ("rank", C.c_int32), ("B_local", C.c_int32), ("noisy", C.c_int32), ("use_entropy", C.c_int32),
                ("emb_teacher", C.c_void_p), ("Wc", C.c_void_p), ("bc", C.c_void_p), ("tWc", C.c_void_p), ("tbc", C.c_void_p),
                ("keep_clean", C.c_void_p), ("keep_strong", C.c_void_p), ("drop_scale", C.c_float),
                ("label_smoothing", C.c_float), ("inv_B_global", C.c_float),
                ("dz_clean", C.c_void_p), ("probs", C.c_void_p), ("logq", C.c_void_p), ("ce_part", C.c_void_p),
                ("n_clean", C.c_int32), ("n_noisy", C.c_int32)]


class DacpArgs(C.Structure):
    _fields_ = [("pack", Pack), ("use_dacp", C.c_int32), ("npow2_", C.c_int32), ("gamma", C.c_float), ("sens_k", C.c_float),
                ("calib_lambda", C.c_float), ("alpha", C.c_float), ("one_minus_alpha", C.c_float), ("fixed_thr", C.c_float),
                ("quality", C.c_void_p), ("anchors", C.c_void_p), ("ema_thr", C.c_void_p), ("mask", C.c_void_p),
                ("out_w", C.c_void_p), ("out_thr", C.c_void_p), ("out_nmask", C.c_void_p), ("epoch_sum", C.c_void_p),
                ("epoch_cnt", C.c_void_p)]


class HeadBwdArgs(C.Structure):
    _fields_ = [("pack", Pack), ("rank", C.c_int32), ("B_local", C.c_int32), ("noisy", C.c_int32), ("n_ce_part", C.c_int32),
                ("Wc", C.c_void_p), ("keep_clean", C.c_void_p), ("keep_strong", C.c_void_p), ("drop_scale", C.c_float),
                ("w_kd", C.c_float), ("inv_B_global", C.c_float), ("dz_clean", C.c_void_p), ("probs", C.c_void_p),
                ("logq", C.c_void_p), ("ce_part", C.c_void_p), ("mask", C.c_void_p), ("nmask", C.c_void_p),
                ("g_clean", C.c_void_p), ("g_strong", C.c_void_p), ("dWc", C.c_void_p), ("dbc", C.c_void_p),
                ("part", C.c_void_p), ("loss_out", C.c_void_p), ("counter", C.c_void_p)]


class EcdaArgs(C.Structure):
    _fields_ = [("pack", Pack), ("rank", C.c_int32), ("class_aware", C.c_int32), ("mask", C.c_void_p), ("nmask", C.c_void_p),
                ("w_class", C.c_void_p), ("lambda_cls", C.c_float), ("gamma_comp", C.c_float), ("delta_rep", C.c_float),
                ("w_ecda", C.c_float), ("g_clean", C.c_void_p), ("g_strong", C.c_void_p), ("loss_out", C.c_void_p),
                ("ws", C.c_void_p), ("ws_bytes", C.c_size_t), ("phases", C.c_int32)]


class SclArgs(C.Structure):
    _fields_ = [("pack", Pack), ("rank", C.c_int32), ("noisy", C.c_int32), ("mask", C.c_void_p), ("temperature", C.c_float),
                ("w_scl", C.c_float), ("g_clean", C.c_void_p), ("g_strong", C.c_void_p), ("loss_out", C.c_void_p),
                ("ws", C.c_void_p), ("ws_bytes", C.c_size_t)]


_lib = None


def lib():
    """Load the shared library once; raise loudly if it is absent (no CPU / PyTorch fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DadError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). dad_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.dad_last_error.restype = C.c_char_p
    L.dad_abi_version.restype = C.c_int
    for name in ("dad_encoder_fwd_workspace_bytes", "dad_encoder_bwd_workspace_bytes", "dad_ecda_workspace_bytes", "dad_scl_workspace_bytes",
                 "dad_aug_encoder_fwd_workspace_bytes", "dad_predict_workspace_bytes"):
        getattr(L, name).restype = C.c_size_t
    L.dad_pack_rank_words.restype = C.c_int64
    L.dad_encoder_fwd_workspace_bytes.argtypes = [C.c_int64, C.c_int64, C.c_int32]
    L.dad_encoder_bwd_workspace_bytes.argtypes = [C.c_int64, C.c_int64, C.c_int32]
    L.dad_encoder_bwd_splits.argtypes = [C.c_int64, C.c_int64, C.c_int32]
    L.dad_ecda_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    L.dad_pack_rank_words.argtypes = [C.c_int32]
    L.dad_head_grid.argtypes = [C.c_int32]
    vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float
    L.dad_augment.argtypes = [vp, vp, C.POINTER(Aug), i64, i64, vp]
    L.dad_encoder_fwd.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(Aug), vp, vp, vp, vp, vp, vp, C.c_size_t, i64, i64, i32, vp]
    L.dad_aug_encoder_fwd_workspace_bytes.argtypes = [i64, i64]
    L.dad_aug_encoder_fwd.argtypes = [vp, vp, vp, vp, vp, C.POINTER(Aug), vp, vp, vp, vp, C.POINTER(Aug), vp, vp, vp, vp, vp, vp, C.c_size_t,
                                      i64, i64, vp]
    L.dad_predict_workspace_bytes.argtypes = [i64, i64]
    L.dad_predict.argtypes = [vp] * 19 + [C.c_size_t, i64, i64, vp]
    L.dad_tc_set_cta_pair.argtypes = [i32]
    L.dad_tc_set_cta_pair.restype = None
    L.dad_tc_set_bwd_products.argtypes = [i32]
    L.dad_tc_set_bwd_products.restype = None
    L.dad_encoder_bwd.argtypes = [vp, vp, C.POINTER(Aug), vp, vp, vp, vp, vp, vp, vp, i32, vp, C.c_size_t, i64, i64, i32, vp]
    L.dad_head_fwd.argtypes = [C.POINTER(HeadFwdArgs), vp]
    L.dad_dacp_mask.argtypes = [C.POINTER(DacpArgs), vp]
    L.dad_head_bwd.argtypes = [C.POINTER(HeadBwdArgs), vp]
    L.dad_ecda.argtypes = [C.POINTER(EcdaArgs), vp]
    L.dad_scl_workspace_bytes.argtypes = [i32, i32]
    L.dad_scl.argtypes = [C.POINTER(SclArgs), vp]
    L.dad_peer_push.argtypes = [vp, C.POINTER(vp), i32, i32, i32, C.c_size_t, C.c_size_t, vp]
    L.dad_peer_barrier.argtypes = [C.POINTER(vp), i32, i32, C.c_size_t, vp, vp]
    L.dad_grad_sum_ranks.argtypes = [vp, i32, i64, vp, vp]
    L.dad_peer_push_barrier.argtypes = [vp, C.POINTER(vp), i32, i32, C.c_size_t, C.c_size_t, C.c_size_t, vp, vp, vp]
    L.dad_grad_reduce_xchg.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp, vp, C.POINTER(vp), i32, i32, C.c_size_t, C.c_size_t, vp, vp, vp]
    L.dad_grad_norm_xchg.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, f32, vp]
    L.dad_gather_pad.argtypes = [vp, vp, vp, vp, vp, i64, i64, vp]
    L.dad_grad_reduce.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp, vp, vp]
    L.dad_grad_norm.argtypes = [vp, vp, vp, vp, vp, vp, f32, vp]
    L.dad_adam_ema.argtypes = [vp, vp, vp, vp, vp, vp, f32, f32, f32, i32, vp, vp, vp, vp, vp, vp, vp, f32, f32, f32, vp]
    L.dad_ema.argtypes = [vp, vp, f32, f32, i64, vp]
    L.dad_split_f16.argtypes = [vp, vp, vp, i64, vp]
    L.dad_certainty_scores.argtypes = [vp, i32, i32, vp, vp, vp]
    L.dad_step_rng.argtypes = [C.c_uint64, vp, vp, vp, i64, f32, vp, f32, vp, i64, i32, i64, vp]
    if L.dad_abi_version() != 3:
        raise DadError("libdad_sm100a.so ABI version mismatch")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise DadError(lib().dad_last_error().decode() + f" (code {rc})")


_device_ok = False


def require_device():
    """Raise unless the current CUDA device is a compute-capability 10.x part."""
    global _device_ok
    if _device_ok:
        return
    if not torch.cuda.is_available():
        raise DadError("dad_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    check(lib().dad_device_check())
    _device_ok = True


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t, dtype=None, name="tensor"):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise DadError(f"{name}: expected a CUDA tensor")
    if not t.is_contiguous():
        raise DadError(f"{name}: expected a contiguous tensor")
    if dtype is not None and t.dtype != dtype:
        raise DadError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t.data_ptr()


STREAM_WEAK, STREAM_STRONG = 2, 3  # Philox streams of the two views of a noisy batch: a (2k, 2k+1) pair shares its Philox blocks


def make_aug(mode=0, noise=None, sigma=0.0, feat_keep=None, t_start=None, mask_len=0, seed=None, stream=0, step_ptr=None, frame_offset=0):
    """Build a dad_aug.  noise tensor -> parity mode (noise_src 0); seed -> in-kernel Philox (noise_src 1)."""
    a = Aug()
    a.mode = int(mode)
    if mode == 0:
        return a
    if noise is not None:
        a.noise_src, a.noise = 0, ptr(noise, torch.float32, "aug.noise")
    else:
        if seed is None:
            raise DadError("augmentation needs either a supplied noise tensor or a Philox seed")
        a.noise_src, a.seed, a.stream = 1, int(seed) & (2 ** 64 - 1), int(stream)
        a.step_ptr = step_ptr
        a.frame_offset = int(frame_offset)
    a.sigma = float(sigma)
    if mode == 2:
        a.feat_keep = ptr(feat_keep, torch.float32, "aug.feat_keep")
        a.t_start = ptr(t_start, torch.int32, "aug.t_start")
        a.mask_len = int(mask_len) if t_start is not None else 0
    return a


def pack_rank_words(B_local):
    return int(B_local) * (2 * D_HID + 3)


def pack_offsets(B_local):
    """Word offsets of the fields inside ONE rank's segment of the exchange buffer (include/dad_sm100a.h `dad_pack`):
    [emb_clean B*H f32 | emb_strong B*H f32 | score B f32 | pred B i32 | label B i32]."""
    B = int(B_local)
    return {"emb_clean": 0, "emb_strong": B * D_HID, "score": 2 * B * D_HID, "pred": 2 * B * D_HID + B,
            "label": 2 * B * D_HID + 2 * B, "words": pack_rank_words(B)}


def pack_global_views(pack_f32, B_local, world):
    """Global-batch views (rank-major sample order) of an all-gathered exchange buffer (any device):
    -> dict(emb_clean [Bg,H] f32, emb_strong [Bg,H] f32, score [Bg] f32, pred [Bg] i32, label [Bg] i32).
    These are strided views of `pack_f32` (no copy), so writes land in the buffer."""
    B, o = int(B_local), pack_offsets(B_local)
    seg = pack_f32.view(world, o["words"])
    segi = pack_f32.view(torch.int32).view(world, o["words"])
    return {
        "emb_clean": seg[:, o["emb_clean"]:o["emb_clean"] + B * D_HID].unflatten(1, (B, D_HID)),
        "emb_strong": seg[:, o["emb_strong"]:o["emb_strong"] + B * D_HID].unflatten(1, (B, D_HID)),
        "score": seg[:, o["score"]:o["score"] + B],
        "pred": segi[:, o["pred"]:o["pred"] + B],
        "label": segi[:, o["label"]:o["label"] + B],
    }
