"""notes array -> MIDI object.  Drop-in for the reference's ``music2midi/utils.py:5-20``."""
from __future__ import annotations

import numpy as np

try:  # the real thing when it is installed (the reference depends on pretty-midi 0.2.10)
    import pretty_midi as _pm
except ImportError:  # not in this image: same interface, own Standard-MIDI-File writer
    from . import midi as _pm


def numpy_to_midi(notes: np.ndarray):
    """(N,4) rows [onset_s, offset_s, pitch, velocity] -> PrettyMIDI(resolution=384, 120 bpm) with one
    "Piano" instrument (program 0); notes with end <= start are removed."""
    midi_data = _pm.PrettyMIDI(resolution=384, initial_tempo=120.0)
    inst = _pm.Instrument(program=0, name="Piano")
    inst.notes = [
        _pm.Note(start=float(on), end=float(off), pitch=int(pitch), velocity=int(vel)) for on, off, pitch, vel in notes
    ]
    midi_data.instruments.append(inst)
    midi_data.remove_invalid_notes()
    return midi_data
