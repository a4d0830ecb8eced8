"""ctypes binding of libmars5_b200.so (include/mars5_b200.h).  No torch types cross this boundary: tensors are passed
as raw ``data_ptr()`` integers.  The library is mandatory: there is no Python/CPU fallback for any entry point."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmars5_b200.so")

M5_OK = 0
MEM_HOST, MEM_DEVICE = 0, 1
DT_F16, DT_F32, DT_U8 = 0, 1, 2
OUT_F32, OUT_F16, OUT_SWIGLU_F16, OUT_F16_SPLIT, OUT_SWIGLU_F16_SPLIT = 0, 1, 2, 3, 4
ACT_NONE, ACT_GELU, ACT_SILU = 0, 1, 2
NUM_FAST, NUM_PRECISE, NUM_MIXED, NUM_MIXED8, NUM_MIXED8K = 0, 1, 2, 3, 4   # m5_nar_cfg.precise (NAR numerics)
NUM_NAMES = {"fast": NUM_FAST, "precise": NUM_PRECISE, "mixed": NUM_MIXED, "mixed8": NUM_MIXED8, "mixed8k": NUM_MIXED8K}
NUM_DEFAULT = NUM_MIXED8K   # what Mars5TTS and bench.py run: holds the 1e-3 max-abs logit bound at full dims


class ModelCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("ar_dim", "ar_heads", "ar_layers", "ar_hidden", "ar_vocab", "ar_text_vocab",
                                         "ar_spk_layers", "ar_spk_ff")] + [("ar_norm_eps", C.c_float), ("ar_pos_alpha", C.c_float)] + \
               [(n, C.c_int32) for n in ("nar_dim", "nar_heads", "nar_enc_layers", "nar_dec_layers", "nar_spk_layers",
                                         "nar_ff", "nar_text_vocab", "n_classes", "n_quant")] + \
               [("ln_eps", C.c_float), ("head_ln_eps", C.c_float), ("nar_pos_alpha", C.c_float),
                ("nar_cond_alpha", C.c_float), ("nar_ref_alpha", C.c_float)] + \
               [(n, C.c_int32) for n in ("voc_feat", "voc_dim", "voc_inter", "voc_layers", "voc_nfft", "voc_hop",
                                         "voc_n_bw", "voc_codebook", "max_pos")]


class Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ptr", C.c_void_p), ("numel", C.c_int64), ("dtype", C.c_int32)]


class ArCfg(C.Structure):
    _fields_ = [("temperature", C.c_float), ("top_k", C.c_int32), ("top_p", C.c_float),
                ("alpha_frequency", C.c_float), ("alpha_presence", C.c_float), ("penalty_window", C.c_int32),
                ("eos_penalty_decay", C.c_float), ("eos_penalty_factor", C.c_float), ("max_len", C.c_int32),
                ("eos_id", C.c_int32), ("force_len", C.c_int32), ("sync_every", C.c_int32), ("typical_p", C.c_float)]


class NarCfg(C.Structure):
    _fields_ = [("T", C.c_int32), ("x0_temp", C.c_float), ("guidance_w", C.c_float),
                ("q0_override_steps", C.c_int32), ("deep_clone", C.c_int32), ("precise", C.c_int32),
                ("schedule", C.c_void_p), ("jump_len", C.c_int32), ("jump_n_sample", C.c_int32), ("scaled_forward", C.c_int32)]


_P = C.c_void_p
_I = C.c_int32
_SIGS = {
    "m5_create": (_I, [_I, C.POINTER(ModelCfg), C.POINTER(Tensor), _I, C.POINTER(_P)]),
    "m5_destroy": (None, [_P]),
    "m5_last_error": (C.c_char_p, [_P]),
    "m5_sync": (_I, [_P]),
    "m5_launch_count": (C.c_int64, [_P]),
    "m5_num_sms": (_I, [_P]),
    "m5_stream": (_P, [_P]),
    "m5_profile_enable": (_I, [_P, _I]),
    "m5_profile_read": (_I, [_P, _I, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "m5_ar_generate": (_I, [_P, _I, _P, _P, _P, _P, _P, C.POINTER(ArCfg), _I, _P, _I, C.c_uint64, _P, _P, _P, _P, _P, _I]),
    "m5_nar_infer": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, C.POINTER(NarCfg), _I, _P, _P, C.c_uint64, _P, _P]),
    "m5_nar_forward": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "m5_ar_forward": (_I, [_P, _I, _P, _P, _P, _P, _I, _P]),
    "m5_vocode": (_I, [_P, _I, _P, _P, _I, _I, _P]),
    "m5_vocode_trim": (_I, [_P, _I, _P, _P, _I, _I, C.c_float, _I, _I, _P, _P, _P]),
    "m5_encodec_encode": (_I, [_P, _I, _P, _P, _I, _I, _P]),
    "m5_dbg_gemm": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I]),
    "m5_dbg_gemm_f8lo": (_I, [_P, _P, _I, _P, _P, _P, _I, _I, _I, _P, _I]),
    "m5_dbg_skinny": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _I, _I, _I]),
    "m5_dbg_norm": (_I, [_P, _P, _I, _I, _P, _P, C.c_float, _I, _P, _P]),
    "m5_dbg_attn": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _I, _I]),
    "m5_dbg_attn_split": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I]),
    "m5_dbg_decode_attn": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _I]),
    "m5_dbg_sample": (_I, [_P, _P, _I, _I, C.POINTER(ArCfg), _I, _P, _I, _P, _P, _P, C.c_uint64, _P, _P]),
    "m5_dbg_posterior": (_I, [_P, _P, _P, _I, _I, _P, C.c_float, C.c_float, _P, _P, _P, _P, _P, C.c_uint64, _P]),
    "m5_dbg_istft": (_I, [_P, _P, _I, _P, _P]),
    "m5_trim_bounds": (_I, [_I, _P, _P, C.c_float, _I, _I, _P, _P, _I]),
    "m5_bpe_create": (_I, [_I, _P, _I, C.POINTER(_P)]),
    "m5_bpe_destroy": (None, [_P]),
    "m5_bpe_encode": (_I, [_P, _P, _P, _I, _P, _P, _I]),
    "m5_bpe_expand": (C.c_int64, [_P, _P, _P, _I, _P, _I, _P, _P, C.c_int64]),
}

_lib = None


def exported_symbols():
    """Names declared in include/mars5_b200.h (used by the CPU test that checks the .so exports them all)."""
    return sorted(_SIGS)


def load(path=None):
    """Load the shared library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("M5_LIB_PATH") or LIB_PATH   # M5_LIB_PATH: an alternative build of the same ABI (tools/ A/B runs)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'` first. "
                           "There is no CPU/PyTorch fallback for the MARS5 hot path.")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError => symbol missing => hard failure
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(t):
    """Raw pointer of a torch tensor / numpy array / None.  A CUDA tensor was produced on torch's current stream and is
    about to be consumed on the context's own (non-blocking) stream: torch's stream is drained before the address is handed
    out, so that no C-ABI call can overtake the kernels that fill its inputs."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        if getattr(t, "is_cuda", False):
            import torch
            torch.cuda.current_stream(t.device).synchronize()
        return C.c_void_p(t.data_ptr())
    if hasattr(t, "ctypes"):
        return C.c_void_p(t.ctypes.data)
    raise TypeError(type(t))


class M5Error(RuntimeError):
    pass


def check(ctx, rc, what):
    if rc != M5_OK:
        msg = load().m5_last_error(ctx).decode() if ctx else ""
        raise M5Error(f"{what} failed with code {rc}: {msg}")
