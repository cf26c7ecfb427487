"""ctypes binding of libm2m_b200.so (C ABI: include/m2m_b200.h).

There is no CPU fallback anywhere in this package: if the shared library is missing, or no
sm_100 device is visible, the hot-path entry points raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libm2m_b200.so")

M2M_FP32, M2M_BF16 = 0, 1
FLAG_GRAPH, FLAG_TIME_ATTN, FLAG_SKIP_FINISHED = 1, 2, 4


class M2MError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libm2m_b200 error {status}: {message}")
        self.status = status


class Config(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32), ("d_model", C.c_int32), ("d_kv", C.c_int32), ("n_heads", C.c_int32),
        ("d_ff", C.c_int32), ("vocab", C.c_int32), ("n_buckets", C.c_int32), ("n_fft", C.c_int32),
        ("hop", C.c_int32), ("n_cond", C.c_int32), ("max_positions", C.c_int32), ("max_enc_len", C.c_int32),
        ("pad_id", C.c_int32), ("bos_id", C.c_int32), ("eos_id", C.c_int32), ("precision", C.c_int32),
        ("ln_eps", C.c_float),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_int64), ("decode_steps", C.c_int64), ("last_attn_ms", C.c_double),
        ("last_generate_ms", C.c_double), ("last_attn_launches", C.c_int64), ("attn_bytes", C.c_int64),
        ("cross_attn_bytes", C.c_int64), ("class_ms", C.c_double * 16), ("class_launches", C.c_int64 * 16),
    ]


# m2m_kernel_class (include/m2m_b200.h): index into Stats.class_ms / class_launches
KERNEL_CLASSES = ("mel_frame", "mel_dft", "mel_band", "cond", "enc_norm", "enc_gemm", "enc_attn", "cross_kv",
                  "dec_self_attn", "dec_cross_attn", "dec_chain", "dec_select")


# every symbol include/m2m_b200.h declares: (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "m2m_abi_version": (C.c_int, []),
    "m2m_last_error": (C.c_char_p, []),
    "m2m_device_count": (C.c_int, []),
    "m2m_default_config": (None, [C.POINTER(Config)]),
    "m2m_ctx_create": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(_P)]),
    "m2m_ctx_destroy": (C.c_int, [_P]),
    "m2m_set_tensor": (C.c_int, [_P, C.c_char_p, _P, C.c_int64, C.c_int]),
    "m2m_set_bucket_luts": (C.c_int, [_P, _P, C.c_int, _P, C.c_int]),
    "m2m_finalize_weights": (C.c_int, [_P]),
    "m2m_logmel": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "m2m_condition": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "m2m_encode": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "m2m_generate_from_embeds": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.POINTER(C.c_int), _P]),
    "m2m_generate": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, C.POINTER(C.c_int), _P]),
    "m2m_decoder_forward": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_int, _P, _P]),
    "m2m_transcribe_host": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, C.c_int, C.c_int, _P, _P]),
    "m2m_tokens_to_notes": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64,
                                      C.POINTER(C.c_int64)]),
    "m2m_tokens_to_notes_batch": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                            _P, C.c_int64, _P, C.POINTER(C.c_int64)]),
    "m2m_stats_reset": (C.c_int, [_P]),
    "m2m_stats_get": (C.c_int, [_P, C.POINTER(Stats)]),
    "m2m_set_flags": (C.c_int, [_P, C.c_uint32]),
    "m2m_debug_chain_trace": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "m2m_debug_gemm_bf16": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P]),
}

_lib: Optional[C.CDLL] = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Loads the shared library; raises if it has not been built (python -m music2midi_b200.build)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
                "Build it with `python -m music2midi_b200.build`."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.m2m_abi_version() != 1:
            raise ImportError(f"ABI version mismatch: {lib.m2m_abi_version()}")
        _lib = lib
        return lib


NOTES_LIB_PATH = os.path.join(HERE, "libm2m_notes.so")
_notes_lib: Optional[C.CDLL] = None


def load_notes() -> C.CDLL:
    """The token -> notes state machine (csrc/notes.cpp).  It is part of libm2m_b200.so; where that library has not been
    built (no nvcc: dataset tooling, CI) the same source compiled with g++ alone (libm2m_notes.so) is used, so the
    tokenizer - integer CPU code in the reference too - does not depend on the CUDA toolchain."""
    global _notes_lib
    if _lib is not None:
        return _lib
    if os.path.exists(LIB_PATH):
        return load()
    with _lock:
        if _notes_lib is None:
            if not os.path.exists(NOTES_LIB_PATH):
                from . import build as _build

                _build.build_notes()
            lib = C.CDLL(NOTES_LIB_PATH)
            for name in ("m2m_tokens_to_notes", "m2m_tokens_to_notes_batch"):
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = SYMBOLS[name]
            lib.m2m_notes_last_error.restype = C.c_char_p
            lib.m2m_last_error = lib.m2m_notes_last_error
            _notes_lib = lib
        return _notes_lib


def check(status: int) -> None:
    if status != 0:
        lib = _lib if _lib is not None else (_notes_lib if _notes_lib is not None else load())
        raise M2MError(status, lib.m2m_last_error().decode("utf-8", "replace"))


def default_config() -> Config:
    cfg = Config()
    load().m2m_default_config(C.byref(cfg))
    return cfg
