"""A small PrettyMIDI-compatible container + Standard MIDI File writer/reader.

The reference returns ``pretty_midi.PrettyMIDI`` objects (music2midi/utils.py:5-20) and its callers
use ``.instruments[i].notes``, ``.write(path)``, ``.remove_invalid_notes()`` and ``.get_end_time()``
(webui.py:62, demo.ipynb).  pretty_midi is not installed in this image, so ``utils.numpy_to_midi``
falls back to these classes, which mirror that part of its interface (same attribute names, same
tick quantisation: tick = round(seconds * resolution * tempo / 60)).
"""
from __future__ import annotations

import struct
from typing import List

import numpy as np


class Note:
    def __init__(self, velocity: int, pitch: int, start: float, end: float):
        self.velocity, self.pitch, self.start, self.end = velocity, pitch, start, end

    def get_duration(self) -> float:
        return self.end - self.start

    @property
    def duration(self) -> float:
        return self.end - self.start

    def __repr__(self):
        return f"Note(start={self.start:f}, end={self.end:f}, pitch={self.pitch}, velocity={self.velocity})"


class Instrument:
    def __init__(self, program: int, is_drum: bool = False, name: str = ""):
        self.program, self.is_drum, self.name = program, is_drum, name
        self.notes: List[Note] = []
        self.pitch_bends: list = []
        self.control_changes: list = []

    def remove_invalid_notes(self) -> None:
        self.notes = [n for n in self.notes if n.end > n.start]

    def get_end_time(self) -> float:
        return max((n.end for n in self.notes), default=0.0)

    def get_piano_roll(self, fs: int = 100, times=None, pedal_threshold=64) -> np.ndarray:
        """(128, n) velocity piano roll like pretty_midi's (no pedal / pitch-bend handling: this container never
        holds control changes).  Without `times`: one column per 1/fs s.  With `times`: column n is the mean of the
        fs-grid columns in [round(times[n] fs), round(times[n+1] fs)); the last column stays zero."""
        if not self.notes:
            return np.zeros((128, 0))
        end_time = self.get_end_time()
        if times is not None and len(times) and times[-1] > end_time:
            end_time = times[-1]
        roll = np.zeros((128, int(fs * end_time)))
        for n in self.notes:
            roll[int(n.pitch), int(n.start * fs):int(n.end * fs)] += n.velocity
        if times is None:
            return roll
        times = np.asarray(times, dtype=np.float64)
        out = np.zeros((128, times.shape[0]))
        ticks = np.array(np.round(times * fs), dtype=np.int64)
        for i, (a, b) in enumerate(zip(ticks[:-1], ticks[1:])):
            if a < roll.shape[1]:
                if a == b:
                    b = a + 1
                out[:, i] = np.mean(roll[:, a:b], axis=1)
        return out

    def fluidsynth(self, fs: int = 44100, sf2_path=None):
        raise RuntimeError(_NO_SYNTH)

    def synthesize(self, fs: int = 44100, wave=None):
        raise RuntimeError(_NO_SYNTH)

    def __repr__(self):
        return f'Instrument(program={self.program}, is_drum={self.is_drum}, name="{self.name}")'


def _vlq(n: int) -> bytes:
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    return bytes(reversed(out))


_NO_SYNTH = ("audio synthesis needs the real pretty_midi package (and pyfluidsynth + a SoundFont for fluidsynth()); "
             "music2midi_b200.midi is the note container / Standard-MIDI-File writer used when pretty_midi is not "
             "installed.  Write the file with .write(path) and render it with any synthesiser, or "
             "`pip install pretty_midi pyfluidsynth` and numpy_to_midi() will return real PrettyMIDI objects.")


class PrettyMIDI:
    def __init__(self, midi_file=None, resolution: int = 220, initial_tempo: float = 120.0):
        self.resolution = resolution
        self.initial_tempo = float(initial_tempo)
        self.instruments: List[Instrument] = []
        if midi_file is not None:
            self._read(midi_file)

    # ---- pretty_midi surface used by the reference's callers
    def remove_invalid_notes(self) -> None:
        for inst in self.instruments:
            inst.remove_invalid_notes()

    def get_end_time(self) -> float:
        return max((i.get_end_time() for i in self.instruments), default=0.0)

    def get_piano_roll(self, fs: int = 100, times=None, pedal_threshold=64) -> np.ndarray:
        """Sum of the instruments' piano rolls (pretty_midi.PrettyMIDI.get_piano_roll)."""
        if not self.instruments:
            return np.zeros((128, 0))
        rolls = [i.get_piano_roll(fs=fs, times=times, pedal_threshold=pedal_threshold) for i in self.instruments]
        out = np.zeros((128, max(r.shape[1] for r in rolls)))
        for r in rolls:
            out[:, : r.shape[1]] += r
        return out

    def fluidsynth(self, fs: int = 44100, sf2_path=None):
        """webui.py:66 / demo.ipynb:77 call this right after generate(); it needs the real pretty_midi."""
        raise RuntimeError(_NO_SYNTH)

    def synthesize(self, fs: int = 44100, wave=None):
        raise RuntimeError(_NO_SYNTH)

    def time_to_tick(self, t: float) -> int:
        return int(round(t * self.resolution * self.initial_tempo / 60.0))

    def tick_to_time(self, tick: int) -> float:
        return tick * 60.0 / (self.initial_tempo * self.resolution)

    def write(self, filename) -> None:
        tracks = [self._tempo_track()]
        channels = [c for c in range(16) if c != 9]
        for i, inst in enumerate(self.instruments):
            ch = 9 if inst.is_drum else channels[i % len(channels)]
            ev = []
            for n in inst.notes:
                # (tick, order, bytes): note-offs sort before note-ons on the same tick
                ev.append((self.time_to_tick(n.start), 1, bytes([0x90 | ch, int(n.pitch) & 0x7F, int(n.velocity) & 0x7F])))
                ev.append((self.time_to_tick(n.end), 0, bytes([0x80 | ch, int(n.pitch) & 0x7F, 0])))
            ev.sort(key=lambda e: (e[0], e[1]))
            body = bytearray()
            if inst.name:
                nm = inst.name.encode("latin-1", "replace")
                body += b"\x00\xff\x03" + _vlq(len(nm)) + nm
            body += b"\x00" + bytes([0xC0 | ch, int(inst.program) & 0x7F])
            last = 0
            for tick, _, msg in ev:
                body += _vlq(max(tick - last, 0)) + msg
                last = max(tick, last)
            body += b"\x01\xff\x2f\x00"
            tracks.append(bytes(body))
        data = b"MThd" + struct.pack(">IHHH", 6, 1, len(tracks), self.resolution)
        for t in tracks:
            data += b"MTrk" + struct.pack(">I", len(t)) + t
        if hasattr(filename, "write"):
            filename.write(data)
        else:
            with open(filename, "wb") as f:
                f.write(data)

    def _tempo_track(self) -> bytes:
        us = int(round(60_000_000 / self.initial_tempo))
        return (b"\x00\xff\x51\x03" + us.to_bytes(3, "big") + b"\x00\xff\x58\x04\x04\x02\x18\x08" + b"\x01\xff\x2f\x00")

    # ---- minimal reader (round-trip tests, and so that write() output can be inspected offline)
    def _read(self, midi_file) -> None:
        data = midi_file.read() if hasattr(midi_file, "read") else open(midi_file, "rb").read()
        assert data[:4] == b"MThd"
        _, _fmt, ntrk, div = struct.unpack(">IHHH", data[4:14])
        self.resolution = div
        pos = 14
        for _ in range(ntrk):
            assert data[pos:pos + 4] == b"MTrk"
            (ln,) = struct.unpack(">I", data[pos + 4:pos + 8])
            trk = data[pos + 8:pos + 8 + ln]
            pos += 8 + ln
            self._read_track(trk)

    def _read_track(self, trk: bytes) -> None:
        i, tick, status = 0, 0, 0
        inst, open_notes, name = None, {}, ""

        def vlq():
            nonlocal i
            v = 0
            while True:
                b = trk[i]
                i += 1
                v = (v << 7) | (b & 0x7F)
                if not b & 0x80:
                    return v

        while i < len(trk):
            tick += vlq()
            b = trk[i]
            if b == 0xFF:
                typ = trk[i + 1]
                i += 2
                ln = vlq()
                payload = trk[i:i + ln]
                i += ln
                if typ == 0x51:
                    self.initial_tempo = 60_000_000 / int.from_bytes(payload, "big")
                elif typ == 0x03:
                    name = payload.decode("latin-1")
                continue
            if b & 0x80:
                status = b
                i += 1
            kind = status & 0xF0
            if kind == 0xC0:
                inst = Instrument(trk[i], is_drum=(status & 0x0F) == 9, name=name)
                self.instruments.append(inst)
                i += 1
            elif kind in (0x80, 0x90):
                pitch, vel = trk[i], trk[i + 1]
                i += 2
                if inst is None:
                    inst = Instrument(0, name=name)
                    self.instruments.append(inst)
                if kind == 0x90 and vel > 0:
                    open_notes.setdefault(pitch, []).append((tick, vel))
                elif open_notes.get(pitch):
                    t0, v0 = open_notes[pitch].pop(0)
                    inst.notes.append(Note(v0, pitch, self.tick_to_time(t0), self.tick_to_time(tick)))
            elif kind in (0xA0, 0xB0, 0xE0):
                i += 2
            elif kind == 0xD0:
                i += 1
            else:  # sysex etc.
                i += 1
                i += vlq()
