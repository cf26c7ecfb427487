"""Audio features and conditioning, B200-native.

Drop-in for the reference's ``music2midi/input.py`` (``ModelInputs`` :9-12, ``LogMelSpectrogram``
:15-41, ``Conditioning`` :44-59): same constructor arguments, same buffer / parameter names in the
state dict (``melspectrogram.spectrogram.window``, ``melspectrogram.mel_scale.fb``,
``embeds.{i}.weight``), same outputs.  The arithmetic runs in libm2m_b200's CUDA kernels (framing +
window + DFT + |.|^2, banded mel, clamp, log); inputs must live on a CUDA device.
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional

import numpy as np
import torch
import torch.nn as nn

from . import synthetic
from .engine import Engine


class ModelInputs(NamedTuple):
    input_waveform: torch.Tensor
    notes_batch: Optional[tuple] = None
    cond_index: Optional[torch.Tensor] = None


class _Buffer(nn.Module):
    def __init__(self, name: str, value: torch.Tensor):
        super().__init__()
        self.register_buffer(name, value)


class _MelBuffers(nn.Module):
    """Holds the two registered buffers under torchaudio's names so checkpoints load unchanged."""

    def __init__(self, sample_rate, n_fft, f_min, n_mels):
        super().__init__()
        self.spectrogram = _Buffer("window", synthetic.hann_window(n_fft))
        self.mel_scale = _Buffer("fb", synthetic.mel_filterbank(sample_rate, n_fft, n_mels, f_min))


class LogMelSpectrogram(nn.Module):
    def __init__(self, sample_rate: int, n_fft: int, hop_length: int, f_min: float, n_mels: int):
        super().__init__()
        self.sample_rate, self.n_fft, self.hop_length, self.f_min, self.n_mels = sample_rate, n_fft, hop_length, f_min, n_mels
        self.melspectrogram = _MelBuffers(sample_rate, n_fft, f_min, n_mels)
        self._engine: Optional[Engine] = None  # injected by T5Transformer, else a frontend-only engine
        self._engine_key = None

    def _frontend_engine(self, device: torch.device) -> Engine:
        win, fb = self.melspectrogram.spectrogram.window, self.melspectrogram.mel_scale.fb
        key = (str(device), win._version, fb._version, win.data_ptr(), fb.data_ptr())
        if self._engine is None or self._engine_key != key:
            eng = Engine(device, "fp32", overrides=dict(n_fft=self.n_fft, hop=self.hop_length, d_model=self.n_mels))
            eng.load_state_dict({
                "spectrogram.melspectrogram.spectrogram.window": win,
                "spectrogram.melspectrogram.mel_scale.fb": fb,
            })
            self._engine, self._engine_key = eng, key
        return self._engine

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: waveform (..., sample) -> log-mel (..., frame, n_mels), float32, no grad."""
        if not x.is_cuda:
            raise RuntimeError("LogMelSpectrogram: input must be a CUDA tensor (B200-native path, no CPU fallback)")
        owner = getattr(self, "_owner_engine", None)
        eng = owner() if owner is not None else self._frontend_engine(x.device)
        with torch.no_grad():
            return eng.logmel(x)


class Conditioning(nn.Module):
    def __init__(self, n_dim: int, num_embeds: List[int]):
        super().__init__()
        self.n_dim = n_dim
        self.embeds = nn.ModuleList([nn.Embedding(num, n_dim) for num in num_embeds])
        self._engine: Optional[Engine] = None
        self._engine_key = None

    def _own_engine(self, device: torch.device) -> Engine:
        ws = [e.weight for e in self.embeds]
        key = (str(device),) + tuple((w._version, w.data_ptr()) for w in ws)
        if self._engine is None or self._engine_key != key:
            eng = Engine(device, "fp32", overrides=dict(d_model=self.n_dim, n_cond=len(ws)))
            sd = {f"conditioning.embeds.{i}.weight": w for i, w in enumerate(ws)}
            sd["spectrogram.melspectrogram.spectrogram.window"] = synthetic.hann_window(eng.cfg.n_fft)
            sd["spectrogram.melspectrogram.mel_scale.fb"] = synthetic.mel_filterbank(n_mels=self.n_dim)
            eng.load_state_dict(sd)
            self._engine, self._engine_key = eng, key
        return self._engine

    def forward(self, feature: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
        """feature (batch, L, n_dim), indices (batch, n_index) -> (batch, n_index + L, n_dim):
        the embedding rows are PREPENDED to the feature sequence."""
        if not feature.is_cuda:
            raise RuntimeError("Conditioning: input must be a CUDA tensor (B200-native path, no CPU fallback)")
        owner = getattr(self, "_owner_engine", None)
        eng = owner() if owner is not None else self._own_engine(feature.device)
        with torch.no_grad():
            return eng.condition(feature, indices.to(feature.device))
