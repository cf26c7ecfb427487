"""Application model: audio in, MIDI out.

Drop-in for the inference surface of the reference's ``music2midi/model.py`` (``Music2MIDI``:
``__init__(config_path)``, ``generate`` :67-99, ``sample_tokens`` :101-140, ``evaluate_batch`` :55-65,
``load_from_checkpoint``, ``.model/.config/.device``).  The reference class is a
``pl.LightningModule``; training (``training_step``/``configure_optimizers``, :27-53) is out of scope
here, so this is a plain ``nn.Module`` that understands Lightning checkpoint files.

Differences that do not change results: segments are independent, so instead of
chunks of ``inference.batch_size`` (=128) the B200 path decodes ``inference.device_batch_size``
segments at a time (2560 when the config does not set it), and on several GPUs clips are sharded across ranks (music2midi_b200/distributed.py).
"""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional, Union

import numpy as np
import torch
import torch.nn as nn

from .config import load_config
from .evaluation import evaluate_batch
from .input import ModelInputs
from .transformer import T5Transformer
from .utils import numpy_to_midi


# Segments decoded as one device batch when the config has no ``inference.device_batch_size`` (the reference's
# ``inference.batch_size`` = 128 bounds the memory of its own GPUs; results do not depend on the batch size, and the
# decode step of this path is latency-bound below a few hundred rows): 2560 segments = 256 clips of 30 s, ~75 GB of KV
# cache in bf16.
DEFAULT_DEVICE_BATCH = 2560


def load_audio(path: Union[str, Path], sr: int) -> np.ndarray:
    """Mono float32 at ``sr`` Hz (the reference calls ``librosa.load(path, sr=sr)``, model.py:84)."""
    try:
        import librosa  # type: ignore

        return librosa.load(str(path), sr=sr)[0]
    except ImportError:
        pass
    import wave as _wave

    from scipy.signal import resample_poly

    with _wave.open(str(path), "rb") as f:
        n_ch, width, rate, n = f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()
        raw = f.readframes(n)
    if width == 2:
        y = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        y = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        y = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"unsupported WAV sample width {width}")
    y = y.reshape(-1, n_ch).mean(axis=1)
    if rate != sr:
        g = np.gcd(int(rate), int(sr))
        y = resample_poly(y, sr // g, rate // g).astype(np.float32)
    return y.astype(np.float32)


class Music2MIDI(nn.Module):
    def __init__(self, config_path: str, precision: Optional[str] = None):
        super().__init__()
        self.config = load_config(config_path)
        self.model = T5Transformer(config_path, precision=precision)
        self.hparams = {"config_path": config_path}
        self.eval()

    # ------------------------------------------------------------------ Lightning-compatible bits
    @property
    def device(self) -> torch.device:
        return self.model.transformer.device

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, config_path: Optional[str] = None, **kwargs):
        """Reads a Lightning ``.ckpt`` of the reference (keys ``model.`` + T5Transformer's state dict;
        ``hyper_parameters.config_path`` used when ``config_path`` is not given)."""
        ckpt = torch.load(str(checkpoint_path), map_location=map_location or "cpu", weights_only=False)
        if config_path is None:
            config_path = (ckpt.get("hyper_parameters") or {}).get("config_path")
        if config_path is None:
            raise ValueError("config_path is required (not stored in the checkpoint)")
        obj = cls(config_path, **kwargs)
        sd = ckpt.get("state_dict", ckpt)
        missing, unexpected = obj.load_state_dict(sd, strict=False)
        # tied aliases may be absent in checkpoints written by other transformers versions
        tied = {"model.transformer.encoder.embed_tokens.weight", "model.transformer.decoder.embed_tokens.weight"}
        missing = [k for k in missing if k not in tied]
        if missing == ["model.transformer.lm_head.weight"] and not unexpected:
            # config.yaml:23 sets tie_word_embeddings: false -> the released checkpoints carry their own lm_head.  One
            # without it was exported by a transformers version that force-ties the head; running with the randomly
            # initialised head instead would produce garbage silently.
            raise RuntimeError("checkpoint has no 'model.transformer.lm_head.weight' (untied lm_head required, "
                               "tie_word_embeddings: false); re-export it with the head, or copy "
                               "transformer.shared.weight into that key if the head really is tied")
        if missing or unexpected:
            raise RuntimeError(f"checkpoint mismatch: missing={missing} unexpected={list(unexpected)}")
        return obj

    # ------------------------------------------------------------------ inference
    @torch.no_grad()
    def evaluate_batch(self, inputs: ModelInputs):
        """reference model.py:55-65: greedy generation capped at 4 tokens per label note, batched token decoding,
        melody chroma accuracy of the generated against the label MIDI -> (score, output_midi, label_midi)."""
        max_num_notes = max(len(notes) for notes in inputs.notes_batch)
        generated = self.model.generate(inputs, max_length=max_num_notes * 4)
        decoded = self.model.tokenizer.decode(generated, mode="batched")
        label_midi = [numpy_to_midi(n) for n in inputs.notes_batch]
        output_midi = [numpy_to_midi(n) for n in decoded]
        return evaluate_batch(label_midi, output_midi), output_midi, label_midi

    def generate(self, audio_path: Optional[Union[str, Path]] = None, audio_y: Optional[np.ndarray] = None,
                 sr: Optional[int] = None, cond_index: Optional[List[int]] = None):
        """Specify either audio_path or audio_y as input.  Returns a PrettyMIDI(-compatible) object."""
        if audio_path is None and audio_y is None:
            raise ValueError("Either audio_path or audio_y should be specified")
        if sr is None:
            sr = self.config.model.sample_rate
        else:
            assert sr == self.config.model.sample_rate
        if audio_y is None:
            audio_y = load_audio(audio_path, sr)
        split_size = int(sr * self.config.dataset.segment_duration)
        pad = int(np.ceil(len(audio_y) / split_size)) * split_size - len(audio_y)
        audio_y = np.pad(np.asarray(audio_y), (0, pad), "constant")
        waveform = torch.from_numpy(audio_y).to(self.device)
        notes = self.sample_tokens(waveform, split_size, split_duration=self.config.dataset.segment_duration,
                                   cond_index=cond_index)
        return numpy_to_midi(notes)

    @torch.no_grad()
    def generate_tokens(self, waveform: torch.Tensor, split_size: int, cond_index: Optional[List[int]] = None,
                        max_length: int = 1024) -> List[torch.Tensor]:
        """Token rows (one per segment, on the device) for a padded 1-D waveform."""
        n_embeds = len(self.model.conditioning.embeds)
        tail = waveform.numel() % split_size
        if tail:  # direct callers with an unpadded waveform: zero-pad the last segment (pad_sequence, model.py:119)
            waveform = torch.nn.functional.pad(waveform, (0, split_size - tail))
        segments = waveform.reshape(-1, split_size)
        inf = self.config.get("inference", {}) or {}
        chunk = int(inf.get("device_batch_size", DEFAULT_DEVICE_BATCH))
        rows: List[torch.Tensor] = []
        for i in range(0, segments.shape[0], chunk):
            wav = segments[i:i + chunk].to(self.device)
            cond = torch.zeros((wav.shape[0], n_embeds))
            if cond_index is not None:
                cond = cond + torch.Tensor(cond_index)
            cond = cond.long().to(self.device)
            tokens = self.model.generate(ModelInputs(input_waveform=wav, cond_index=cond), max_length=max_length)
            rows += [*tokens]
        return rows

    def _pinned_stage(self, n_floats: int) -> torch.Tensor:
        buf = getattr(self, "_stage_buf", None)
        if buf is None or buf.numel() < n_floats:
            buf = torch.empty(n_floats, dtype=torch.float32, pin_memory=torch.cuda.is_available())
            self._stage_buf = buf
        return buf[:n_floats]

    @torch.no_grad()
    def generate_many(self, audios, cond_index: Optional[List[int]] = None, distributed: bool = False):
        """Batch entry point (not in the reference, which transcribes one recording per call): transcribes a list
        of recordings (float32 arrays at the model sample rate) as ONE device batch of independent 3 s segments and
        returns one MIDI object per recording.  With ``distributed=True`` and an initialised ``torch.distributed``
        process group, every rank passes the same list, clips are sharded clip-wise over the ranks and the token
        streams are all-gathered (music2midi_b200/distributed.py)."""
        from . import distributed as dist_mod

        sr = self.config.model.sample_rate
        dur = self.config.dataset.segment_duration
        split = int(sr * dur)
        clips = [np.asarray(y, dtype=np.float32).reshape(-1) for y in audios]
        counts = [max(1, -(-len(y) // split)) for y in clips]
        if not clips:
            return []
        lo_clip, hi_clip = 0, len(clips)
        if distributed and torch.distributed.is_initialized():
            lo_clip, hi_clip = dist_mod.shard_range(len(clips), torch.distributed.get_rank(),
                                                    torch.distributed.get_world_size())
        local = clips[lo_clip:hi_clip]
        n_local = sum(counts[lo_clip:hi_clip])
        if local:
            # one pinned staging buffer for all local segments (kept between calls: page-locking half a gigabyte costs
            # more than the upload), zero-padded per clip in place, then the host-buffer entry point of the library:
            # the upload of device batch i + 1 overlaps the decode of batch i, tokens come back as int16
            stage = self._pinned_stage(n_local * split).view(n_local, split)
            flat, pos = stage.view(-1).numpy(), 0
            for y, n in zip(local, counts[lo_clip:hi_clip]):
                flat[pos:pos + len(y)] = y
                flat[pos + len(y):pos + n * split] = 0.0
                pos += n * split
            n_embeds = len(self.model.conditioning.embeds)
            cond = np.zeros((n_local, n_embeds), dtype=np.int64)
            if cond_index is not None:
                cond += np.asarray(torch.Tensor(cond_index).long().numpy(), dtype=np.int64)
            inf = self.config.get("inference", {}) or {}
            chunk = int(inf.get("device_batch_size", DEFAULT_DEVICE_BATCH))
            toks, _ = self.model.engine().transcribe_host(stage.numpy(), cond, 1024, device_batch=chunk)
            tok = torch.from_numpy(toks)
        else:
            tok = torch.zeros(0, 1024, dtype=torch.int64)
        if distributed and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size()
            rows = [sum(counts[slice(*dist_mod.shard_range(len(clips), r, world))]) for r in range(world)]
            tok = dist_mod.gather_tokens(tok.to(self.device), sum(counts), counts=rows)
        tok = tok.to(torch.int64).cpu()
        out, pos = [], 0
        for n in counts:
            notes = self.model.tokenizer.decode(tok[pos:pos + n], mode="sequential", duration_per_batch=dur)
            out.append(numpy_to_midi(notes))
            pos += n
        return out

    @torch.no_grad()
    def sample_tokens(self, waveform: torch.Tensor, split_size: int, split_duration: float,
                      cond_index: Optional[List[int]] = None) -> np.ndarray:
        """(N,4) float64 notes [onset_s, offset_s, pitch, velocity] of the whole recording."""
        rows = self.generate_tokens(waveform, split_size, cond_index, max_length=1024)
        return self.model.tokenizer.decode(rows, mode="sequential", duration_per_batch=split_duration)
