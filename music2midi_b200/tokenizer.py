"""MIDI token vocabulary: token streams <-> note arrays.

Drop-in for the reference's ``music2midi/tokenizer.py`` (``MidiTokenizer`` + the ``PAD BOS EOS
ONSET OFFSET`` constants; reference lines 11-222).  The token -> notes direction (the one on the
inference path, reference :143-200 and the njit helper :242-267) runs in the C++ state machine
``m2m_tokens_to_notes`` of libm2m_b200; the notes -> tokens direction (labels for the
teacher-forced forward, reference :98-141,202-222) is vectorised numpy.

Vocabulary (config.yaml:32-38): 0 PAD, 1 BOS, 2 EOS, 3 ONSET, 4 OFFSET, 5..132 pitch, 133..332
time in 50 ms steps.  Ids above 332 are never produced by tokenisation but decode as time
tokens (reference :187-190), so they are NOT clamped here either.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Union

import numpy as np
import torch

from . import _lib

PAD = 0
BOS = 1
EOS = 2
ONSET = 3
OFFSET = 4

_NAMES = {PAD: "PAD", BOS: "BOS", EOS: "EOS", ONSET: "ONSET", OFFSET: "OFFSET"}


def _get(cfg, name):
    return cfg[name] if isinstance(cfg, dict) else getattr(cfg, name)


class MidiTokenizer:
    def __init__(self, config):
        self.config = _get(config, "tokenizer")
        vs = _get(self.config, "vocab_size")
        self.time_step = _get(self.config, "midi_quantize_ms") / 1000
        self.pitch_token_offset = int(_get(vs, "special"))
        self.time_token_offset = self.pitch_token_offset + int(_get(vs, "pitch"))
        self.n_time = int(_get(vs, "time"))
        self.default_velocity = int(_get(self.config, "default_velocity"))

    # ------------------------------------------------------------------ strings
    def to_string(self, tokens) -> List[str]:
        out = []
        for t in tokens:
            t = int(t)
            if t in _NAMES:
                out.append(_NAMES[t])
            elif t >= self.time_token_offset:
                out.append(f"time_{t - self.time_token_offset}")
            elif t >= self.pitch_token_offset:
                out.append(f"note_{t - self.pitch_token_offset}")
            else:
                raise ValueError(f"Invalid token '{t}'")
        return out

    # ------------------------------------------------------------------ tokens -> notes
    def decode(self, tokens_batch: Iterable[Union[np.ndarray, torch.Tensor]], mode: str = "batched",
               duration_per_batch: Optional[float] = None, cutoff_time: Optional[int] = None):
        """``mode="batched"``: one (n,4) float64 array per row.  ``mode="sequential"``: rows are
        consecutive segments of one recording; segment i is shifted by i*duration_per_batch and all
        notes are concatenated.  Rows: [onset_s, offset_s, pitch, velocity]."""
        if mode not in ("batched", "sequential"):
            raise ValueError(f"Invalid argument mode={mode}")
        rows = self._rows_to_host(tokens_batch)
        if mode == "batched":
            return [self._decode(r, 0, cutoff_time) for r in rows]
        assert duration_per_batch is not None, 'duration_per_batch is required for mode="sequential"'
        n_steps = round(duration_per_batch / self.time_step)
        if len(rows) > 1 and len({r.shape for r in rows}) == 1 and rows[0].ndim == 1:
            # equal-length rows (a token matrix): one C call for the whole recording
            return self._finish(self._decode_matrix(np.stack(rows), n_steps), cutoff_time)
        out = [self._decode(r, i * n_steps, cutoff_time) for i, r in enumerate(rows)]
        return np.concatenate(out)  # raises on an empty batch, like the reference

    @staticmethod
    def _rows_to_host(tokens_batch) -> List[np.ndarray]:
        # one device->host copy for a whole tensor batch instead of one sync per row
        if isinstance(tokens_batch, torch.Tensor):
            return list(tokens_batch.detach().cpu().numpy())
        tokens_batch = list(tokens_batch)
        if tokens_batch and all(isinstance(t, torch.Tensor) and t.is_cuda for t in tokens_batch):
            lens = [int(t.numel()) for t in tokens_batch]
            if len(set(lens)) == 1:
                return list(torch.stack(tokens_batch).cpu().numpy())
        return [t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t) for t in tokens_batch]

    def _decode(self, tokens, start_idx: int = 0, cutoff_time: Optional[int] = None) -> np.ndarray:
        if isinstance(tokens, torch.Tensor):
            tokens = tokens.detach().cpu().numpy()
        return self._finish(self._decode_tokens(np.asarray(tokens), start_idx), cutoff_time)

    def _finish(self, notes: np.ndarray, cutoff_time: Optional[int]) -> np.ndarray:
        notes = notes[notes[:, 1] != -1]  # un-closed notes are dropped
        notes[:, :2] = notes[:, :2] * self.time_step
        if cutoff_time is not None:
            notes = notes[notes[:, 0] < cutoff_time]
            notes[:, 1] = np.where(notes[:, 1] > cutoff_time, cutoff_time, notes[:, 1])
        return notes

    def _decode_tokens(self, tokens: np.ndarray, start_idx: int) -> np.ndarray:
        """(n,4) float64 rows [onset_idx, offset_idx | -1, pitch, velocity] in time-step units."""
        toks = np.ascontiguousarray(tokens, dtype=np.int64).reshape(-1)
        cap = max(int(toks.size), 1)
        rows = np.empty((cap, 4), dtype=np.int64)
        n = C.c_int64(0)
        _lib.check(_lib.load_notes().m2m_tokens_to_notes(
            toks.ctypes.data_as(C.c_void_p), toks.size, int(start_idx), self.pitch_token_offset,
            self.time_token_offset, self.default_velocity, rows.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
        return rows[: n.value].astype(np.float64)

    def _decode_matrix(self, tokens: np.ndarray, steps_per_row: int, start_idx: int = 0) -> np.ndarray:
        """All rows of a [n_rows, L] token matrix in one call; row i starts at start_idx + i * steps_per_row."""
        toks = np.ascontiguousarray(tokens, dtype=np.int64)
        cap = max(int(toks.size), 1)
        rows = np.empty((cap, 4), dtype=np.int64)
        n = C.c_int64(0)
        _lib.check(_lib.load_notes().m2m_tokens_to_notes_batch(
            toks.ctypes.data_as(C.c_void_p), toks.shape[0], toks.shape[1], int(start_idx), int(steps_per_row),
            self.pitch_token_offset, self.time_token_offset, self.default_velocity, rows.ctypes.data_as(C.c_void_p), cap,
            None, C.byref(n)))
        return rows[: n.value].astype(np.float64)

    # ------------------------------------------------------------------ notes -> tokens
    def __call__(self, notes_batch: Iterable[np.ndarray], cutoff_time: Optional[int] = None) -> torch.Tensor:
        assert isinstance(notes_batch, Iterable), "notes should be passed in batch"
        rows = [self._tokenize(n, cutoff_time) for n in notes_batch]
        width = max((r.numel() for r in rows), default=0)
        out = torch.full((len(rows), width), PAD, dtype=torch.long)
        for i, r in enumerate(rows):
            out[i, : r.numel()] = r
        return out

    def _tokenize(self, notes: np.ndarray, cutoff_time: Optional[int] = None) -> torch.Tensor:
        """Per unique time index: [time, ONSET, onset pitches.., OFFSET, offset pitches..]; EOS last.
        Notes last at least one step; indices are rint(nextafter(t / step)) clipped to the vocabulary."""
        toks: List[int] = []
        if len(notes) > 0:
            n = np.array(notes, dtype=np.float64, copy=True)
            if cutoff_time is not None:
                n = n[n[:, 0] < cutoff_time]
            n[:, 1] = np.maximum(n[:, 1], n[:, 0] + self.time_step)
            idx = n[:, :2] / self.time_step
            idx = np.minimum(np.rint(np.nextafter(idx, idx + 1)), self.n_time - 1)
            pitch_tok = n[:, 2] + self.pitch_token_offset
            for t in np.unique(idx):
                on = pitch_tok[idx[:, 0] == t]
                off = pitch_tok[idx[:, 1] == t]
                toks.append(int(t + self.time_token_offset))
                if on.size:
                    toks.append(ONSET)
                    toks.extend(on.tolist())
                if off.size:
                    toks.append(OFFSET)
                    toks.extend(off.tolist())
        toks.append(EOS)
        return torch.tensor(toks, dtype=torch.float32).long()  # reference builds via torch.Tensor(...).long()
