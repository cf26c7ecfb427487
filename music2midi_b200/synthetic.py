"""Seeded synthetic inputs and random-init weights at config.yaml's dimensions.

There is no released checkpoint or dataset offline, so parity tests and benchmarks run on
synthetic audio and explicitly *set* weights.  Everything here is generated
on the CPU with a seeded ``torch.Generator`` so that the authoring container, the GPU box,
the oracle and the CUDA path all see bit-identical inputs.

Weights: every tensor of the reference's ``T5Transformer`` state dict is assigned (150 keys,
reference music2midi/transformer.py:11-26), including an **untied** ``lm_head`` because
``config.yaml:23`` sets ``tie_word_embeddings: false``.  The scales are larger
than T5's own initialiser on purpose: attention is peaked and logits are spread, so greedy
decoding is not a trivially constant stream and numerical mistakes show up as token flips.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional

import torch

SAMPLE_RATE = 16000
SEGMENT_SAMPLES = 48000  # int(16000 * 3)  (reference music2midi/model.py:86)

D_MODEL = 384
D_KV = 64
N_HEADS = 8
INNER = N_HEADS * D_KV
D_FF = 1152
VOCAB = 400
N_LAYERS = 6
N_BUCKETS = 32
N_FFT = 2048
N_FREQ = N_FFT // 2 + 1
COND_SIZES = (6, 3)  # genre, difficulty  (config.yaml:48-50)


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


# --------------------------------------------------------------------------- audio
def audio_noise(n_segments: int, seed: int = 0, samples: int = SEGMENT_SAMPLES) -> torch.Tensor:
    """Generator A: white noise 0.1*N(0,1), float32 [n_segments, samples]."""
    return 0.1 * torch.randn(n_segments, samples, generator=_gen(seed), dtype=torch.float32)


def audio_tones(n_segments: int, seed: int = 0, samples: int = SEGMENT_SAMPLES) -> torch.Tensor:
    """Generator B: 3-6 sinusoids in 55..4000 Hz, amplitudes 0.5..0.01, plus 1e-3 noise.

    High dynamic range across mel bands: stresses the log-mel tolerance.
    """
    g = _gen(1_000_003 + seed)
    t = torch.arange(samples, dtype=torch.float64) / SAMPLE_RATE
    out = torch.empty(n_segments, samples, dtype=torch.float32)
    for i in range(n_segments):
        k = int(torch.randint(3, 7, (1,), generator=g))
        logf = torch.rand(k, generator=g, dtype=torch.float64) * (math.log(4000.0) - math.log(55.0)) + math.log(55.0)
        freqs = torch.exp(logf)
        amps = torch.exp(
            torch.rand(k, generator=g, dtype=torch.float64) * (math.log(0.5) - math.log(0.01)) + math.log(0.01)
        )
        phase = torch.rand(k, generator=g, dtype=torch.float64) * 2 * math.pi
        y = (amps[:, None] * torch.sin(2 * math.pi * freqs[:, None] * t[None, :] + phase[:, None])).sum(0)
        y = y + 1e-3 * torch.randn(samples, generator=g, dtype=torch.float64)
        out[i] = y.to(torch.float32)
    return out


def audio_zeros(n_segments: int, samples: int = SEGMENT_SAMPLES) -> torch.Tensor:
    """Generator C: silence -> every log-mel value is log(1e-6)."""
    return torch.zeros(n_segments, samples, dtype=torch.float32)


# --------------------------------------------------------------------------- buffers
def hann_window(n_fft: int = N_FFT) -> torch.Tensor:
    """Periodic Hann, what torchaudio's Spectrogram registers as ``window``."""
    return torch.hann_window(n_fft, periodic=True, dtype=torch.float32)


def _hz_to_mel_htk(f: float) -> float:
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_filterbank(
    sample_rate: int = SAMPLE_RATE,
    n_fft: int = N_FFT,
    n_mels: int = D_MODEL,
    f_min: float = 20.0,
    f_max: Optional[float] = None,
) -> torch.Tensor:
    """HTK triangular filterbank [n_freq, n_mels], norm=None, in fp32 torch ops.

    Same sequence of fp32 tensor operations as torchaudio.functional.melscale_fbanks
    (site-packages/torchaudio/functional/functional.py, "htk", norm=None), so the result is
    bit-identical to the buffer the reference registers as ``mel_scale.fb``
    (tests/test_oracle_cpu.py checks that against torchaudio itself).
    """
    n_freqs = n_fft // 2 + 1
    if f_max is None:
        f_max = float(sample_rate // 2)
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = _hz_to_mel_htk(f_min)
    m_max = _hz_to_mel_htk(f_max)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    zero = torch.zeros(1)
    down_slopes = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up_slopes = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(zero, torch.min(down_slopes, up_slopes))
    return fb.contiguous()


# --------------------------------------------------------------------------- weights
def _attn_keys(prefix: str):
    return [f"{prefix}.{n}.weight" for n in ("q", "k", "v", "o")]


def state_dict_keys() -> list:
    """The 150 keys of the reference T5Transformer state dict, in its order."""
    keys = ["transformer.shared.weight", "transformer.encoder.embed_tokens.weight"]
    for l in range(N_LAYERS):
        p = f"transformer.encoder.block.{l}.layer"
        keys += _attn_keys(f"{p}.0.SelfAttention")
        if l == 0:
            keys.append(f"{p}.0.SelfAttention.relative_attention_bias.weight")
        keys.append(f"{p}.0.layer_norm.weight")
        keys += [f"{p}.1.DenseReluDense.{n}.weight" for n in ("wi_0", "wi_1", "wo")]
        keys.append(f"{p}.1.layer_norm.weight")
    keys.append("transformer.encoder.final_layer_norm.weight")
    keys.append("transformer.decoder.embed_tokens.weight")
    for l in range(N_LAYERS):
        p = f"transformer.decoder.block.{l}.layer"
        keys += _attn_keys(f"{p}.0.SelfAttention")
        if l == 0:
            keys.append(f"{p}.0.SelfAttention.relative_attention_bias.weight")
        keys.append(f"{p}.0.layer_norm.weight")
        keys += _attn_keys(f"{p}.1.EncDecAttention")
        keys.append(f"{p}.1.layer_norm.weight")
        keys += [f"{p}.2.DenseReluDense.{n}.weight" for n in ("wi_0", "wi_1", "wo")]
        keys.append(f"{p}.2.layer_norm.weight")
    keys.append("transformer.decoder.final_layer_norm.weight")
    keys.append("transformer.lm_head.weight")
    keys += [
        "spectrogram.melspectrogram.spectrogram.window",
        "spectrogram.melspectrogram.mel_scale.fb",
        "conditioning.embeds.0.weight",
        "conditioning.embeds.1.weight",
    ]
    return keys


def _shape_of(key: str):
    if key.endswith("relative_attention_bias.weight"):
        return (N_BUCKETS, N_HEADS)
    if key.endswith("layer_norm.weight"):
        return (D_MODEL,)
    if key.endswith(("shared.weight", "embed_tokens.weight", "lm_head.weight")):
        return (VOCAB, D_MODEL)
    if key.endswith((".q.weight", ".k.weight", ".v.weight")):
        return (INNER, D_MODEL)
    if key.endswith(".o.weight"):
        return (D_MODEL, INNER)
    if key.endswith(("wi_0.weight", "wi_1.weight")):
        return (D_FF, D_MODEL)
    if key.endswith("wo.weight"):
        return (D_MODEL, D_FF)
    if key.endswith("spectrogram.window"):
        return (N_FFT,)
    if key.endswith("mel_scale.fb"):
        return (N_FREQ, D_MODEL)
    if key.endswith("embeds.0.weight"):
        return (COND_SIZES[0], D_MODEL)
    if key.endswith("embeds.1.weight"):
        return (COND_SIZES[1], D_MODEL)
    raise KeyError(key)


def synthetic_state_dict(seed: int = 0, gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic fp32 weights for every key of the reference state dict.

    ``gain`` scales the attention-logit and output-logit spread (1.0 = the default used by
    the golden fixtures).
    """
    g = _gen(7_000_000 + seed)
    sd: Dict[str, torch.Tensor] = OrderedDict()

    def normal(shape, std):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    for key in state_dict_keys():
        shape = _shape_of(key)
        if key.endswith("spectrogram.window"):
            sd[key] = hann_window()
        elif key.endswith("mel_scale.fb"):
            sd[key] = mel_filterbank()
        elif key.endswith("embed_tokens.weight"):
            sd[key] = sd["transformer.shared.weight"]  # tied to `shared` in HF T5
        elif key.endswith("shared.weight"):
            sd[key] = normal(shape, 1.0)
        elif key.endswith("lm_head.weight"):
            sd[key] = normal(shape, gain * 2.0 * D_MODEL ** -0.5)
        elif key.endswith("layer_norm.weight"):
            sd[key] = 1.0 + normal(shape, 0.1)
        elif key.endswith("relative_attention_bias.weight"):
            sd[key] = normal(shape, gain * 1.0)
        elif key.endswith(".q.weight"):
            # T5 does not scale scores by 1/sqrt(d_kv); std here gives score std ~ gain*1.5
            sd[key] = normal(shape, gain * 1.5 * (D_MODEL ** -0.5) * (D_KV ** -0.25))
        elif key.endswith(".k.weight"):
            sd[key] = normal(shape, (D_MODEL ** -0.5) * (D_KV ** -0.25))
        elif key.endswith(".v.weight"):
            sd[key] = normal(shape, D_MODEL ** -0.5)
        elif key.endswith(".o.weight"):
            sd[key] = normal(shape, INNER ** -0.5)
        elif key.endswith(("wi_0.weight", "wi_1.weight")):
            sd[key] = normal(shape, D_MODEL ** -0.5)
        elif key.endswith("wo.weight"):
            sd[key] = normal(shape, D_FF ** -0.5)
        elif "conditioning.embeds" in key:
            sd[key] = normal(shape, 1.0)
        else:  # pragma: no cover
            raise KeyError(key)
    return sd
