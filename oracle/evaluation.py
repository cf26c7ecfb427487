"""TEST INFRASTRUCTURE ONLY - plain-loop restatement of the reference's melody chroma metric
(music2midi/evaluation.py:10-75) and of the third-party arithmetic it calls.

PARITY UNPINNED: pretty_midi 0.2.10, librosa 0.10.1 and mir_eval 0.6 (pinned in the reference's environment.yaml)
are absent from this image and from /root/reference, and the reference has no test or golden vector for the metric,
so this restates their PUBLISHED algorithms (pretty_midi Instrument.get_piano_roll, librosa.midi_to_hz,
mir_eval.melody.hz2cents / raw_chroma_accuracy) element by element, deliberately written differently from the
vectorised product code (music2midi_b200/evaluation.py) it checks.  Notes are (onset_s, offset_s, pitch, velocity).
"""
from __future__ import annotations

import math
from typing import List, Sequence


def piano_roll(notes: Sequence[Sequence[float]], fs: int, times: List[float]) -> List[List[float]]:
    """pretty_midi get_piano_roll(fs, times): roll[p][int(start*fs):int(end*fs)] += velocity on a grid of
    int(fs * max(end_time, times[-1])) columns, then column n = mean of roll[:, round(t_n fs) : round(t_{n+1} fs)]
    (at least one column; the last output column stays zero)."""
    if len(notes) == 0:
        return [[] for _ in range(128)]
    end = max(n[1] for n in notes)
    if times and times[-1] > end:
        end = times[-1]
    width = int(fs * end)
    roll = [[0.0] * width for _ in range(128)]
    for on, off, pitch, vel in notes:
        for c in range(int(on * fs), min(int(off * fs), width)):
            roll[int(pitch)][c] += vel
    ticks = [int(round(t * fs)) for t in times]  # python round == numpy round (half to even)
    out = [[0.0] * len(times) for _ in range(128)]
    for n in range(len(times) - 1):
        s, e = ticks[n], ticks[n + 1]
        if s < width:
            if s == e:
                e = s + 1
            e = min(e, width)
            for p in range(128):
                seg = roll[p][s:e]
                out[p][n] = sum(seg) / len(seg)
    return out


def highest_pitches(roll: List[List[float]]) -> List[int]:
    n = len(roll[0]) if roll else 0
    res = []
    for c in range(n):
        best = -1
        for p in range(128):
            if roll[p][c] != 0:
                best = p
        res.append(best)
    return res


def melody_pair(target_notes, output_notes, fs: int = 100):
    end_t = max((n[1] for n in target_notes), default=0.0)
    end_o = max((n[1] for n in output_notes), default=0.0)
    end = max(end_t, end_o)
    step = 1 / fs
    times = [k * step for k in range(max(0, math.ceil(end / step)))]  # numpy.arange(0, end, step), value by value
    t = highest_pitches(piano_roll(target_notes, fs, times))
    o = highest_pitches(piano_roll(output_notes, fs, times))
    if len(t) == 0 and len(o) > 0:
        t = [0] * len(o)
    if len(o) == 0 and len(t) > 0:
        o = [0] * len(t)
    return t, o


def chroma_accuracy(ref: Sequence[int], est: Sequence[int], tol: float = 50.0) -> float:
    def cents(p):
        if p < 0:
            return 0.0, False
        hz = 440.0 * 2.0 ** ((p - 69.0) / 12.0)
        return 1200.0 * math.log2(hz / 10.0), True

    voiced, correct = 0, 0
    for r, e in zip(ref, est):
        rc, rv = cents(r)
        ec, ev = cents(e)
        if not rv:
            continue
        voiced += 1
        if ec > 0:
            d = abs(rc - ec)
            d -= 1200.0 * math.floor(d / 1200.0 + 0.5)
            if abs(d) < tol:
                correct += 1
    return correct / voiced if voiced else 0.0


def evaluate_batch(target_notes_batch, output_notes_batch) -> float:
    T, O = [], []
    for t, o in zip(target_notes_batch, output_notes_batch):
        a, b = melody_pair(t, o)
        T += a
        O += b
    return chroma_accuracy(T, O)
