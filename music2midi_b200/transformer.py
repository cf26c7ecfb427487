"""Encoder-decoder network wrapper, B200-native.

Drop-in for the reference's ``music2midi/transformer.py:10-45`` (``T5Transformer``): same
constructor (``config_path``), same attributes (``config, t5config, transformer, tokenizer,
spectrogram, conditioning``), same state-dict keys (so Lightning checkpoints of the reference load
unchanged), ``forward(ModelInputs) -> output with .loss/.logits`` and
``generate(ModelInputs, **kwargs) -> LongTensor`` with HF greedy semantics.

The reference delegates to ``transformers.T5ForConditionalGeneration``; here ``.transformer`` is a
parameter container with HF's module/parameter names and all arithmetic runs in libm2m_b200
(music2midi_b200/csrc): encoder, cross-KV, KV-cached greedy decode with on-device stop detection.
"""
from __future__ import annotations

import os
import threading
import weakref
from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .config import load_config
from .engine import Engine
from .input import Conditioning, LogMelSpectrogram, ModelInputs
from .tokenizer import MidiTokenizer


class _Linear(nn.Module):
    def __init__(self, n_in: int, n_out: int, std: float):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(n_out, n_in) * std)


class _Norm(nn.Module):
    def __init__(self, d: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))


class _Attention(nn.Module):
    def __init__(self, d_model, inner, d_kv, n_heads, n_buckets, has_bias):
        super().__init__()
        self.q = _Linear(d_model, inner, (d_model * d_kv) ** -0.5)
        self.k = _Linear(d_model, inner, d_model ** -0.5)
        self.v = _Linear(d_model, inner, d_model ** -0.5)
        self.o = _Linear(inner, d_model, inner ** -0.5)
        if has_bias:
            self.relative_attention_bias = nn.Embedding(n_buckets, n_heads)
            nn.init.normal_(self.relative_attention_bias.weight, std=d_model ** -0.5)


class _SelfAttnLayer(nn.Module):
    def __init__(self, *a):
        super().__init__()
        self.SelfAttention = _Attention(*a)
        self.layer_norm = _Norm(a[0])


class _CrossAttnLayer(nn.Module):
    def __init__(self, *a):
        super().__init__()
        self.EncDecAttention = _Attention(*a)
        self.layer_norm = _Norm(a[0])


class _GatedFF(nn.Module):
    def __init__(self, d_model, d_ff):
        super().__init__()
        self.wi_0 = _Linear(d_model, d_ff, d_model ** -0.5)
        self.wi_1 = _Linear(d_model, d_ff, d_model ** -0.5)
        self.wo = _Linear(d_ff, d_model, d_ff ** -0.5)


class _FFLayer(nn.Module):
    def __init__(self, d_model, d_ff):
        super().__init__()
        self.DenseReluDense = _GatedFF(d_model, d_ff)
        self.layer_norm = _Norm(d_model)


class _Block(nn.Module):
    def __init__(self, cfg, is_decoder, has_bias):
        super().__init__()
        a = (cfg.d_model, cfg.num_heads * cfg.d_kv, cfg.d_kv, cfg.num_heads, cfg.relative_attention_num_buckets)
        layers = [_SelfAttnLayer(*a, has_bias)]
        if is_decoder:
            layers.append(_CrossAttnLayer(*a, False))
        layers.append(_FFLayer(cfg.d_model, cfg.d_ff))
        self.layer = nn.ModuleList(layers)


class _Stack(nn.Module):
    def __init__(self, cfg, shared, is_decoder, n_layers):
        super().__init__()
        self.embed_tokens = shared  # same Embedding object as `shared`, as in HF
        self.block = nn.ModuleList([_Block(cfg, is_decoder, i == 0) for i in range(n_layers)])
        self.final_layer_norm = _Norm(cfg.d_model)


class T5Weights(nn.Module):
    """Parameter tree with the names of HF ``T5ForConditionalGeneration`` (state-dict compatible).
    ``lm_head`` is an independent parameter: ``tie_word_embeddings: false`` (config.yaml:23)."""

    def __init__(self, cfg):
        super().__init__()
        self.config = cfg
        self.shared = nn.Embedding(cfg.vocab_size, cfg.d_model)
        self.encoder = _Stack(cfg, self.shared, False, cfg.num_layers)
        self.decoder = _Stack(cfg, self.shared, True, cfg.num_decoder_layers)
        self.lm_head = _Linear(cfg.d_model, cfg.vocab_size, 1.0)

    @property
    def device(self) -> torch.device:
        return self.shared.weight.device

    @torch.no_grad()
    def generate(self, inputs_embeds: torch.Tensor = None, max_length: int = 20, **kwargs) -> torch.Tensor:
        """HF-style entry used as ``model.transformer.generate(inputs_embeds=..., max_length=...)``
        (reference transformer.py:44): greedy decoding from encoder input embeddings."""
        if inputs_embeds is None:
            raise ValueError("inputs_embeds is required (the reference never passes input_ids)")
        if kwargs.pop("do_sample", False) or kwargs.pop("num_beams", 1) != 1:
            raise NotImplementedError("only greedy decoding is implemented (the reference never samples)")
        if "max_new_tokens" in kwargs:
            max_length = int(kwargs.pop("max_new_tokens")) + 1
        if kwargs:
            raise TypeError(f"unsupported generate() arguments: {sorted(kwargs)}")
        owner = getattr(self, "_owner_engine", None)
        if owner is None:
            raise RuntimeError("T5Weights.generate needs its owning T5Transformer (engine not attached)")
        eng = owner()
        return eng.generate_from_embeds(inputs_embeds.to(eng.device), int(max_length))


def _t5_config(node) -> SimpleNamespace:
    d = dict(d_kv=64, num_heads=8, relative_attention_max_distance=128, layer_norm_epsilon=1e-6,
             relative_attention_num_buckets=32, num_decoder_layers=None)  # HF T5Config defaults
    d.update(dict(node))
    if d["num_decoder_layers"] is None:
        d["num_decoder_layers"] = d["num_layers"]
    if d.get("feed_forward_proj", "gated-gelu") != "gated-gelu":
        raise ValueError("only feed_forward_proj='gated-gelu' is implemented (config.yaml:21)")
    if d["num_layers"] != d["num_decoder_layers"]:
        raise ValueError("encoder and decoder depth must match")
    return SimpleNamespace(**d)


class Seq2SeqOutput(SimpleNamespace):
    """``.loss`` / ``.logits`` like HF's Seq2SeqLMOutput (also indexable: out["loss"])."""

    def __getitem__(self, k):
        return getattr(self, k)


class T5Transformer(nn.Module):
    def __init__(self, config_path, precision: Optional[str] = None):
        super().__init__()
        self.config = load_config(config_path)
        self.t5config = _t5_config(self.config.model.t5)
        self.transformer = T5Weights(self.t5config)
        self.tokenizer = MidiTokenizer(self.config)
        self.spectrogram = LogMelSpectrogram(
            sample_rate=self.config.model.sample_rate, n_mels=self.config.model.t5.d_model, **self.config.spectrogram
        )
        self.conditioning = Conditioning(
            self.config.model.t5.d_model, [len(v) for v in self.config.conditioning.values()]
        )
        inf = self.config.get("inference", {}) or {}
        self.precision = precision or os.environ.get("M2M_PRECISION") or inf.get("precision", "fp32")
        self._engine: Optional[Engine] = None
        self._engine_key = None
        self._engine_lock = threading.Lock()  # the reference's web UI shares one model between Flask threads
        ref = weakref.ref(self)
        # the sub-modules share this model's context instead of building their own
        object.__setattr__(self.spectrogram, "_owner_engine", lambda: ref().engine())
        object.__setattr__(self.conditioning, "_owner_engine", lambda: ref().engine())
        object.__setattr__(self.transformer, "_owner_engine", lambda: ref().engine())
        self.eval()

    # ------------------------------------------------------------------ engine management
    def set_precision(self, precision: str) -> "T5Transformer":
        self.precision = precision
        return self

    def _weights_key(self):
        sd_items = list(self.state_dict(keep_vars=True).items())
        return (self.precision,) + tuple((k, v.data_ptr(), v._version) for k, v in sd_items)

    def engine(self) -> Engine:
        """The CUDA context for the device the parameters live on; weights are (re)uploaded when any
        parameter, buffer, device or precision changed since the last call."""
        dev = self.transformer.device
        if dev.type != "cuda":
            raise RuntimeError(
                "T5Transformer parameters are on the CPU; move the model to a B200 with .cuda() / .to('cuda'). "
                "There is no CPU fallback."
            )
        key = (str(dev),) + self._weights_key()
        with self._engine_lock:  # one creator / swapper at a time; an old context is closed only under its own lock
            if self._engine is None or self._engine_key != key:
                if self._engine is not None:
                    old = self._engine
                    with old._lock:  # no other thread is inside a library call on the old context
                        old.close()
                t5 = self.t5config
                inf = self.config.get("inference", {}) or {}
                eng = Engine(dev, self.precision, max_enc_len=int(inf.get("max_enc_len", 512)),
                             max_distance=t5.relative_attention_max_distance, overrides=dict(
                    n_layers=t5.num_layers, d_model=t5.d_model, d_kv=t5.d_kv, n_heads=t5.num_heads, d_ff=t5.d_ff,
                    vocab=t5.vocab_size, n_buckets=t5.relative_attention_num_buckets, max_positions=t5.n_positions,
                    n_fft=self.spectrogram.n_fft, hop=self.spectrogram.hop_length, n_cond=len(self.conditioning.embeds),
                    pad_id=t5.pad_token_id, bos_id=t5.decoder_start_token_id, eos_id=t5.eos_token_id,
                    ln_eps=t5.layer_norm_epsilon))
                eng.load_state_dict(self.state_dict())
                self._engine, self._engine_key = eng, key
            return self._engine

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def forward(self, inputs: ModelInputs, **kwargs) -> Seq2SeqOutput:
        """Teacher-forced forward + cross-entropy (inference-shape benchmark only: no autograd).
        Labels come from the tokenizer, PAD -> -100, and are shifted right with the start token."""
        if kwargs:
            raise TypeError(f"unsupported arguments: {sorted(kwargs)}")
        eng = self.engine()
        t5 = self.t5config
        labels = self.tokenizer(inputs.notes_batch)
        labels[labels == t5.pad_token_id] = -100
        labels = labels.to(eng.device)
        dec_in = torch.cat([labels.new_full((labels.shape[0], 1), t5.decoder_start_token_id), labels[:, :-1]], dim=1)
        dec_in = dec_in.masked_fill(dec_in == -100, t5.pad_token_id)
        wave = inputs.input_waveform.to(eng.device)
        enc_in = eng.condition(eng.logmel(wave), inputs.cond_index.to(eng.device))
        enc = eng.encode(enc_in)
        logits = eng.decoder_forward(enc, dec_in)
        loss = F.cross_entropy(logits.view(-1, logits.size(-1)), labels.view(-1), ignore_index=-100)
        return Seq2SeqOutput(loss=loss, logits=logits, encoder_last_hidden_state=enc)

    @torch.no_grad()
    def generate(self, inputs: ModelInputs, **kwargs) -> torch.Tensor:
        """Greedy decoding (the only mode the reference uses).  ``max_length`` defaults to HF's 20."""
        max_length = kwargs.pop("max_length", None)
        max_new = kwargs.pop("max_new_tokens", None)
        if kwargs.pop("do_sample", False) or kwargs.pop("num_beams", 1) != 1:
            raise NotImplementedError("only greedy decoding is implemented (the reference never samples)")
        if kwargs:
            raise TypeError(f"unsupported generate() arguments: {sorted(kwargs)}")
        if max_new is not None:
            max_length = int(max_new) + 1
        if max_length is None:
            max_length = 20
        eng = self.engine()
        wave = inputs.input_waveform.to(eng.device)
        cond = inputs.cond_index
        if cond is None:
            raise ValueError("cond_index is required (the reference indexes it unconditionally, input.py:57)")
        return eng.generate(wave, cond.to(eng.device), int(max_length))
