"""Host-side owner of one libm2m_b200 context (one per GPU): weight upload and typed wrappers
around the C-ABI entry points for torch CUDA tensors.  PyTorch is used for device memory and
streams only; all arithmetic of the hot path happens inside the library's CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import math
import threading
from typing import Dict, Mapping, Optional

import numpy as np
import torch

from . import _lib
from ._lib import M2M_BF16, M2M_FP32, M2MError, check

PRECISIONS = {"fp32": M2M_FP32, "bf16": M2M_BF16}


def relative_position_bucket(rel: torch.Tensor, bidirectional: bool, num_buckets: int = 32, max_distance: int = 128):
    """T5 bucket function with the same fp32 tensor ops as HF transformers
    (models/t5/modeling_t5.py `_relative_position_bucket`), evaluated on the host once; the
    kernels only ever see the resulting look-up tables (SURVEY.md §7 hard part 5)."""
    ret = torch.zeros_like(rel)
    if bidirectional:
        num_buckets //= 2
        ret = ret + (rel > 0).to(torch.long) * num_buckets
        rel = torch.abs(rel)
    else:
        rel = -torch.min(rel, torch.zeros_like(rel))
    max_exact = num_buckets // 2
    is_small = rel < max_exact
    large = max_exact + (
        torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)
    ).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return ret + torch.where(is_small, rel, large)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


class Engine:
    """One context on one device.  Not re-entrant: calls are serialised with a lock (the
    reference's web UI shares one model between Flask threads, webui.py:61,90-93)."""

    def __init__(self, device: torch.device, precision: str = "fp32", max_enc_len: int = 512,
                 overrides: Optional[Mapping[str, int]] = None, max_distance: int = 128):
        self.lib = _lib.load()
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(
                f"music2midi_b200 runs on CUDA (sm_100a) only; got device '{device}'. There is no CPU fallback."
            )
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {list(PRECISIONS)}")
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        self.precision = precision
        cfg = _lib.default_config()
        cfg.precision = PRECISIONS[precision]
        cfg.max_enc_len = max_enc_len
        for k, v in (overrides or {}).items():
            setattr(cfg, k, v)
        self.cfg = cfg
        self._ctx = C.c_void_p()
        check(self.lib.m2m_ctx_create(C.byref(cfg), self.device.index, C.byref(self._ctx)))
        self._lock = threading.RLock()
        self.model_ready = False
        self.max_distance = int(max_distance)  # T5Config.relative_attention_max_distance (bucket LUTs are built with it)

    # ------------------------------------------------------------------ lifetime / weights
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.m2m_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, sd: Mapping[str, torch.Tensor]) -> None:
        """Uploads every tensor of a reference-format state dict (keys of SURVEY.md §5, with or
        without the Lightning ``model.`` prefix) and packs the weight arena."""
        with self._lock:
            has_model = False
            for k, v in sd.items():
                a = v.detach().to(device="cpu", dtype=torch.float32).contiguous().numpy()
                check(self.lib.m2m_set_tensor(self._ctx, k.encode(), a.ctypes.data_as(C.c_void_p), a.size, 0))
                has_model = has_model or "transformer." in k
            g = self.cfg
            n_enc = 2 * g.max_enc_len - 1
            rel = torch.arange(n_enc, dtype=torch.long) - (g.max_enc_len - 1)
            enc = relative_position_bucket(rel, True, g.n_buckets, self.max_distance).to(torch.int32).contiguous().numpy()
            dec = relative_position_bucket(-torch.arange(g.max_positions, dtype=torch.long), False, g.n_buckets,
                                           self.max_distance).to(torch.int32).contiguous().numpy()
            check(self.lib.m2m_set_bucket_luts(self._ctx, enc.ctypes.data_as(C.c_void_p), enc.size,
                                               dec.ctypes.data_as(C.c_void_p), dec.size))
            check(self.lib.m2m_finalize_weights(self._ctx))
            self.model_ready = has_model

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _in(self, t: torch.Tensor, dtype, name: str) -> torch.Tensor:
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if t.device.type != "cuda":
            raise RuntimeError(f"{name} is on '{t.device}': music2midi_b200 has no CPU path; move it to {self.device}")
        if t.device != self.device:
            raise RuntimeError(f"{name} is on {t.device} but the engine lives on {self.device}")
        return t.to(dtype).contiguous()

    # ------------------------------------------------------------------ hot path
    def logmel(self, wave: torch.Tensor) -> torch.Tensor:
        """[..., S] -> [..., 1 + S // hop, d_model] fp32 (music2midi/input.py:33-41)."""
        w = self._in(wave, torch.float32, "waveform")
        lead, S = w.shape[:-1], w.shape[-1]
        w2 = w.reshape(-1, S)
        B = w2.shape[0]
        T = 1 + S // self.cfg.hop
        out = torch.empty(B, T, self.cfg.d_model, dtype=torch.float32, device=self.device)
        with self._lock, torch.cuda.device(self.device):
            check(self.lib.m2m_logmel(self._ctx, _ptr(w2), B, S, _ptr(out), self._stream()))
        return out.reshape(*lead, T, self.cfg.d_model)

    def condition(self, feature: torch.Tensor, cond_index: torch.Tensor) -> torch.Tensor:
        f = self._in(feature, torch.float32, "feature")
        ci = self._in(cond_index, torch.int64, "cond_index")
        B, T, D = f.shape
        if ci.shape != (B, self.cfg.n_cond):
            raise ValueError(f"cond_index must have shape ({B}, {self.cfg.n_cond}), got {tuple(ci.shape)}")
        out = torch.empty(B, T + self.cfg.n_cond, D, dtype=torch.float32, device=self.device)
        with self._lock, torch.cuda.device(self.device):
            check(self.lib.m2m_condition(self._ctx, _ptr(f), _ptr(ci), B, T, _ptr(out), self._stream()))
        return out

    def encode(self, embeds: torch.Tensor) -> torch.Tensor:
        e = self._in(embeds, torch.float32, "inputs_embeds")
        B, L, D = e.shape
        out = torch.empty_like(e)
        with self._lock, torch.cuda.device(self.device):
            check(self.lib.m2m_encode(self._ctx, _ptr(e), B, L, _ptr(out), self._stream()))
        return out

    def generate_from_embeds(self, embeds: torch.Tensor, max_length: int, forced: Optional[torch.Tensor] = None,
                             return_logits: bool = False):
        e = self._in(embeds, torch.float32, "inputs_embeds")
        B, L, D = e.shape
        tokens = torch.empty(B, max_length, dtype=torch.int64, device=self.device)
        f = None if forced is None else self._in(forced, torch.int64, "forced tokens")
        if f is not None and f.shape != (B, max_length):
            raise ValueError("forced tokens must be [B, max_length]")
        logits = None
        if return_logits:
            logits = torch.zeros(B, max(max_length - 1, 0), self.cfg.vocab, dtype=torch.float32, device=self.device)
        n = C.c_int(0)
        with self._lock, torch.cuda.device(self.device):
            check(self.lib.m2m_generate_from_embeds(self._ctx, _ptr(e), B, L, max_length, _ptr(f), _ptr(tokens),
                                                    _ptr(logits), C.byref(n), self._stream()))
        tokens = tokens[:, : n.value]
        return (tokens, logits) if return_logits else tokens

    def generate(self, wave: torch.Tensor, cond_index: torch.Tensor, max_length: int) -> torch.Tensor:
        """music2midi/transformer.py:41-45 as one library call."""
        w = self._in(wave, torch.float32, "waveform")
        ci = self._in(cond_index, torch.int64, "cond_index")
        B, S = w.shape
        tokens = torch.empty(B, max_length, dtype=torch.int64, device=self.device)
        n = C.c_int(0)
        with self._lock, torch.cuda.device(self.device):
            check(self.lib.m2m_generate(self._ctx, _ptr(w), _ptr(ci), B, S, max_length, _ptr(tokens), C.byref(n),
                                        self._stream()))
        return tokens[:, : n.value]

    def decoder_forward(self, enc: torch.Tensor, dec_in: torch.Tensor) -> torch.Tensor:
        e = self._in(enc, torch.float32, "encoder output")
        d = self._in(dec_in, torch.int64, "decoder_input_ids")
        B, L, _ = e.shape
        Ld = d.shape[1]
        logits = torch.empty(B, Ld, self.cfg.vocab, dtype=torch.float32, device=self.device)
        with self._lock, torch.cuda.device(self.device):
            check(self.lib.m2m_decoder_forward(self._ctx, _ptr(e), B, L, _ptr(d), Ld, _ptr(logits), self._stream()))
        return logits

    def transcribe_host(self, wave: np.ndarray, cond: Optional[np.ndarray], max_length: int, device_batch: int):
        """Host buffers in, host tokens out (H2D and D2H inside the call)."""
        w = np.ascontiguousarray(wave, dtype=np.float32)
        n, S = w.shape
        tokens = np.empty((n, max_length), dtype=np.int64)
        lens = np.empty((n,), dtype=np.int32)
        c = None if cond is None else np.ascontiguousarray(cond, dtype=np.int64)
        with self._lock, torch.cuda.device(self.device):
            check(self.lib.m2m_transcribe_host(
                self._ctx, w.ctypes.data_as(C.c_void_p), n, S,
                C.c_void_p(0) if c is None else c.ctypes.data_as(C.c_void_p), max_length, device_batch,
                tokens.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p)))
        return tokens, lens

    def debug_gemm_bf16(self, A: torch.Tensor, W: torch.Tensor, path: int) -> torch.Tensor:
        """Test hook: fp32 C = A @ W.T for bf16 CUDA operands through one GEMM kernel (see the header)."""
        A = self._in(A, torch.bfloat16, "A")
        W = self._in(W, torch.bfloat16, "W")
        M, K = A.shape
        N = W.shape[0]
        out = torch.empty(M, N, dtype=torch.float32, device=self.device)
        with self._lock, torch.cuda.device(self.device):
            check(self.lib.m2m_debug_gemm_bf16(self._ctx, _ptr(A), _ptr(W), M, N, K, _ptr(out), path, self._stream()))
        return out

    # ------------------------------------------------------------------ introspection
    def set_flags(self, graph: bool = True, time_classes: bool = False, skip_finished: bool = True,
                  no_tensor_cores: bool = False, mel: str = "auto", no_tc_attention: bool = False,
                  no_chain: bool = False, no_persistent_gemm: bool = False, no_f32_tc: bool = False):
        """mel: "auto" / "tc" (tcgen05 three-term split DFT, both precisions) or "simt" (fp32 CUDA-core DFT).
        no_f32_tc: fp32 contexts run their GEMMs on the CUDA-core FFMA kernel instead of the tcgen05 split-product one.
        time_classes: instrumented pass (CUDA events around every launch group, summed per kernel class).
        no_chain: bf16 contexts run the decode step as separate RMSNorm / GEMM launches (A/B testing)."""
        f = (1 if graph else 0) | (2 if time_classes else 0) | (4 if skip_finished else 0) | (8 if no_tensor_cores else 0)
        f |= {"auto": 0, "simt": 16, "tc": 32}[mel] | (64 if no_tc_attention else 0) | (128 if no_chain else 0) | (256 if no_persistent_gemm else 0) | (512 if no_f32_tc else 0)
        check(self.lib.m2m_set_flags(self._ctx, f))

    def stats(self, reset: bool = False) -> Dict[str, float]:
        s = _lib.Stats()
        check(self.lib.m2m_stats_get(self._ctx, C.byref(s)))
        if reset:
            check(self.lib.m2m_stats_reset(self._ctx))
        out = {k: getattr(s, k) for k, _ in s._fields_ if not k.startswith("class_")}
        out["class_ms"] = {n: s.class_ms[i] for i, n in enumerate(_lib.KERNEL_CLASSES)}
        out["class_launches"] = {n: int(s.class_launches[i]) for i, n in enumerate(_lib.KERNEL_CLASSES)}
        return out


__all__ = ["Engine", "M2MError", "relative_position_bucket"]
