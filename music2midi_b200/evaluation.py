"""Melody chroma accuracy of generated MIDI against label MIDI.

Drop-in for the reference's ``music2midi/evaluation.py:10-75`` (called from ``Music2MIDI.evaluate_batch``,
``music2midi/model.py:55-65``).  The reference computes the metric with pretty_midi piano rolls, librosa's
``midi_to_hz`` and mir_eval's ``to_cent_voicing`` / ``raw_chroma_accuracy``; none of those packages is a dependency
here, so their published arithmetic is restated in numpy (pretty_midi 0.2.10 ``get_piano_roll``, librosa 0.10.1
``midi_to_hz``, mir_eval 0.6 ``melody.hz2cents`` / ``raw_chroma_accuracy``).  CPU code by nature (a quality metric
over a few thousand piano-roll frames), off the GPU hot path.

One documented difference: for a frame in which no note sounds, the reference's numba kernel reads
``onset_pitches[-1]`` of an EMPTY array (``evaluation.py:16-18``) - an out-of-bounds read whose result is undefined.
Here such frames are unvoiced (pitch -1 -> frequency 0), which is what the ``np.nan`` assignment two lines earlier
evidently intends.
"""
from __future__ import annotations

from typing import Iterable, Tuple

import numpy as np

UNVOICED = -1


def get_highest_pitches_from_piano_roll(piano_roll: np.ndarray) -> np.ndarray:
    """(128, n_frames) piano roll -> highest sounding pitch per frame (int), UNVOICED where nothing sounds."""
    piano_roll = np.asarray(piano_roll)
    n_pitch, n_frames = piano_roll.shape
    on = piano_roll != 0
    highest = n_pitch - 1 - np.argmax(on[::-1, :], axis=0) if n_frames else np.zeros(0, dtype=np.int64)
    return np.where(on.any(axis=0), highest, UNVOICED).astype(np.int64)


def extract_midi_melody(target, output, fs: int = 100) -> Tuple[np.ndarray, np.ndarray]:
    """Two arrays: per frame (1/fs seconds) the highest sounding pitch of `target` and of `output`
    (reference evaluation.py:23-44; both rolls are sampled on the same time base up to the later end time)."""
    end_time = max(output.get_end_time(), target.get_end_time())
    times = np.arange(0, end_time, 1 / fs)
    t = get_highest_pitches_from_piano_roll(target.get_piano_roll(fs=fs, times=times))
    o = get_highest_pitches_from_piano_roll(output.get_piano_roll(fs=fs, times=times))
    if len(t) == 0 and len(o) > 0:
        t = np.zeros_like(o)
    if len(o) == 0 and len(t) > 0:
        o = np.zeros_like(t)
    return t, o


def midi_to_hz(pitch: np.ndarray) -> np.ndarray:
    """librosa.midi_to_hz; UNVOICED (-1) frames map to 0 Hz (= unvoiced for mir_eval)."""
    pitch = np.asarray(pitch, dtype=np.float64)
    return np.where(pitch < 0, 0.0, 440.0 * (2.0 ** ((pitch - 69.0) / 12.0)))


def hz2cents(freq_hz: np.ndarray, base_frequency: float = 10.0) -> np.ndarray:
    """mir_eval.melody.hz2cents: 1200 log2(f / 10 Hz), 0 where f == 0."""
    freq_hz = np.asarray(freq_hz, dtype=np.float64)
    cents = np.zeros(freq_hz.shape[0])
    nz = np.flatnonzero(freq_hz)
    cents[nz] = 1200.0 * np.log2(np.abs(freq_hz[nz]) / base_frequency)
    return cents


def raw_chroma_accuracy(ref_voicing, ref_cent, est_voicing, est_cent, cent_tolerance: float = 50.0) -> float:
    """mir_eval.melody.raw_chroma_accuracy: share of reference-voiced frames whose estimate is within
    `cent_tolerance` cents of the reference after folding the difference to the nearest octave."""
    ref_voicing = np.asarray(ref_voicing).astype(bool)
    est_voicing = np.asarray(est_voicing).astype(bool)
    if ref_voicing.size == 0 or est_voicing.size == 0 or ref_cent.size == 0 or est_cent.size == 0:
        return 0.0
    if ref_voicing.sum() == 0:
        return 0.0
    matching = ref_voicing & (est_cent > 0)
    diff = np.abs(ref_cent - est_cent)[matching]
    octave = 1200.0 * np.floor(diff / 1200.0 + 0.5)
    correct = np.abs(diff - octave) < cent_tolerance
    return float(np.sum(correct) / float(ref_voicing.sum()))


def melody_chroma_accuracy(ref_pitch: np.ndarray, est_pitch: np.ndarray, fs: int = 100) -> float:
    """reference evaluation.py:47-62 (identical time bases, so mir_eval's to_cent_voicing does not resample)."""
    ref_pitch, est_pitch = np.asarray(ref_pitch), np.asarray(est_pitch)
    assert ref_pitch.shape[0] == len(ref_pitch)
    assert ref_pitch.shape == est_pitch.shape
    ref_freq, est_freq = midi_to_hz(ref_pitch), midi_to_hz(est_pitch)
    return raw_chroma_accuracy(ref_freq > 0, hz2cents(ref_freq), est_freq > 0, hz2cents(est_freq))


def evaluate_batch(targets: Iterable, outputs: Iterable) -> float:
    """reference evaluation.py:65-75: melodies of all pairs concatenated, one accuracy for the batch."""
    data = [extract_midi_melody(t, o) for t, o in zip(targets, outputs)]
    if not data:
        return 0.0
    t, o = zip(*data)
    return melody_chroma_accuracy(np.concatenate(t), np.concatenate(o))
