"""TEST INFRASTRUCTURE ONLY — the reference's CPU code path without /root/reference.

The reference's hot path is four call sites into third-party libraries that ARE present in
this image (and on the GPU box, same image): ``torchaudio.transforms.MelSpectrogram``
(music2midi/input.py:25-31,39) and ``transformers.T5ForConditionalGeneration`` +
``.generate`` (music2midi/transformer.py:16,35-37,44).  /root/reference itself does not
travel to the GPU box, so this module re-creates those call sites, argument for argument, so
the *same library code the reference executes* can be timed on the box's host cores
(bench.py ``cpu_baseline`` / ``--impl reference``) and used as a second oracle next to
oracle/port.py.  tests/test_oracle_cpu.py proves (where /root/reference is present) that this
path is bit-identical to the real reference classes given the same state dict.

Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

T5_KW = dict(  # config.yaml:17-31 (model.t5)
    num_layers=6, num_decoder_layers=6, d_model=384, d_ff=1152, feed_forward_proj="gated-gelu",
    tie_word_embeddings=False, tie_encoder_decoder=False, vocab_size=400, n_positions=1024,
    relative_attention_num_buckets=32, pad_token_id=0, bos_token_id=1, eos_token_id=2,
    decoder_start_token_id=1,
)


class HFReferencePath(nn.Module):
    """Same submodule names as the reference T5Transformer so its state dict loads 1:1."""

    def __init__(self):
        super().__init__()
        import torchaudio
        from transformers import T5Config, T5ForConditionalGeneration

        self.t5config = T5Config(**T5_KW)  # transformer.py:14
        self.transformer = T5ForConditionalGeneration(self.t5config)  # transformer.py:16

        class _Spec(nn.Module):  # input.py:15-41
            def __init__(self):
                super().__init__()
                self.melspectrogram = torchaudio.transforms.MelSpectrogram(
                    sample_rate=16000, n_fft=2048, hop_length=256, f_min=20.0, n_mels=384
                )

            def forward(self, x):
                with torch.no_grad():
                    x = self.melspectrogram(x.float()).transpose(-2, -1)
                    x = x.clamp(min=1e-6).log()
                return x

        class _Cond(nn.Module):  # input.py:44-59
            def __init__(self):
                super().__init__()
                self.embeds = nn.ModuleList([nn.Embedding(6, 384), nn.Embedding(3, 384)])

            def forward(self, feature, indices):
                e = torch.stack([emb(indices[:, i]) for i, emb in enumerate(self.embeds)], dim=1)
                return torch.cat([e, feature], dim=1)

        self.spectrogram = _Spec()
        self.conditioning = _Cond()
        self.eval()

    def load_weights(self, sd: Dict[str, torch.Tensor]) -> "HFReferencePath":
        """Assign every tensor explicitly.  transformers>=5 force-ties lm_head to `shared`
        (SURVEY.md §0.5); assigning a fresh Parameter un-ties it and generate() keeps it."""
        own = self.state_dict()
        missing = [k for k in own if k not in sd]
        assert not missing, missing
        with torch.no_grad():
            for k, v in sd.items():
                if k == "transformer.lm_head.weight":
                    continue
                own[k].copy_(v)
            self.transformer.lm_head.weight = nn.Parameter(sd["transformer.lm_head.weight"].clone())
        self.eval()
        return self

    @torch.no_grad()
    def encode(self, wave, cond_index):
        x = self.conditioning(self.spectrogram(wave), cond_index)
        return self.transformer.encoder(inputs_embeds=x).last_hidden_state

    @torch.no_grad()
    def generate(self, wave: torch.Tensor, cond_index: torch.Tensor, **kwargs) -> torch.Tensor:
        """transformer.py:41-45."""
        x = self.conditioning(self.spectrogram(wave), cond_index)
        return self.transformer.generate(inputs_embeds=x, **kwargs)

    @torch.no_grad()
    def teacher_forced_logits(self, wave, cond_index, decoder_input_ids):
        x = self.conditioning(self.spectrogram(wave), cond_index)
        return self.transformer(inputs_embeds=x, decoder_input_ids=decoder_input_ids).logits

    @torch.no_grad()
    def forward_loss(self, wave, cond_index, labels):
        """transformer.py:28-39 with labels already tokenised and PAD -> -100."""
        x = self.conditioning(self.spectrogram(wave), cond_index)
        out = self.transformer(inputs_embeds=x, labels=labels)
        return out.loss, out.logits


def build(sd: Optional[Dict[str, torch.Tensor]] = None) -> HFReferencePath:
    m = HFReferencePath()
    if sd is not None:
        m.load_weights(sd)
    return m
