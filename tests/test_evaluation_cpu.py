"""Melody chroma metric and the PrettyMIDI-compatible surface it needs: the numpy product code
(music2midi_b200/evaluation.py, midi.py) against the plain-loop restatement in oracle/evaluation.py on seeded
random note sets, plus hand-derived known answers.  Parity with pretty_midi / mir_eval themselves is unpinned:
neither is installed here (oracle/evaluation.py header)."""
import io

import numpy as np
import pytest

from music2midi_b200 import evaluation as ev
from music2midi_b200.midi import PrettyMIDI
from music2midi_b200.utils import numpy_to_midi
from oracle import evaluation as oev


def random_notes(rng, n, span=6.0):
    on = np.round(rng.uniform(0, span, n) / 0.05) * 0.05
    dur = np.round(rng.uniform(0.05, 1.0, n) / 0.05) * 0.05
    pitch = rng.integers(30, 100, n)
    return np.stack([on, on + dur, pitch, np.full(n, 80)], 1)


@pytest.mark.parametrize("seed", range(6))
def test_piano_roll_and_melody_match_the_loop_oracle(seed):
    rng = np.random.default_rng(seed)
    a, b = random_notes(rng, 5 + 7 * seed), random_notes(rng, 3 + 5 * seed, span=7.5)
    ma, mb = numpy_to_midi(a), numpy_to_midi(b)
    end = max(ma.get_end_time(), mb.get_end_time())
    times = np.arange(0, end, 0.01)
    roll = ma.get_piano_roll(fs=100, times=times)
    ref = np.array(oev.piano_roll(a.tolist(), 100, times.tolist()))
    assert roll.shape == ref.shape and np.array_equal(roll, ref)
    t, o = ev.extract_midi_melody(ma, mb)
    rt, ro = oev.melody_pair(a.tolist(), b.tolist())
    assert t.tolist() == rt and o.tolist() == ro
    assert ev.melody_chroma_accuracy(t, o) == pytest.approx(oev.chroma_accuracy(rt, ro), abs=1e-12)


def test_evaluate_batch_matches_oracle_and_known_answers():
    rng = np.random.default_rng(42)
    tg = [random_notes(rng, 12), random_notes(rng, 20), random_notes(rng, 1)]
    out = [random_notes(rng, 10), tg[1].copy(), random_notes(rng, 4)]
    score = ev.evaluate_batch([numpy_to_midi(n) for n in tg], [numpy_to_midi(n) for n in out])
    assert score == pytest.approx(oev.evaluate_batch([n.tolist() for n in tg], [n.tolist() for n in out]), abs=1e-12)
    assert 0.0 < score < 1.0
    # identical transcription -> 1; one octave up -> chroma still 1; one semitone up -> 0
    base = np.array([[0.0, 1.0, 60, 80], [1.0, 2.0, 64, 80], [2.0, 3.0, 67, 80]])
    up12, up1 = base.copy(), base.copy()
    up12[:, 2] += 12
    up1[:, 2] += 1
    m = numpy_to_midi
    assert ev.evaluate_batch([m(base)], [m(base)]) == 1.0
    assert ev.evaluate_batch([m(base)], [m(up12)]) == 1.0
    assert ev.evaluate_batch([m(base)], [m(up1)]) == 0.0
    # estimate covers only the first of three seconds: 100 of the ~300 voiced reference frames are right
    assert ev.evaluate_batch([m(base)], [m(base[:1])]) == pytest.approx(1 / 3, abs=0.01)
    # nothing transcribed at all: the reference substitutes pitch 0 (8.18 Hz) for the empty output -> 0
    assert ev.evaluate_batch([m(base)], [m(np.zeros((0, 4)))]) == 0.0


def test_highest_pitch_and_silent_frames():
    roll = np.zeros((128, 4))
    roll[60, 0] = roll[72, 0] = 80
    roll[40, 2] = 1
    assert ev.get_highest_pitches_from_piano_roll(roll).tolist() == [72, ev.UNVOICED, 40, ev.UNVOICED]
    assert ev.midi_to_hz(np.array([69, 57, ev.UNVOICED])).tolist() == [440.0, 220.0, 0.0]
    assert ev.hz2cents(np.array([10.0, 20.0, 0.0])).tolist() == [0.0, 1200.0, 0.0]


def test_standard_midi_file_bytes_golden():
    """Hand-derived SMF bytes of a 3-note case at resolution 384, 120 bpm (tick = seconds * 768): format 1, a tempo
    track (500000 us per quarter, 4/4) and one "Piano" track with program 0 on channel 0."""
    notes = np.array([[0.0, 0.5, 60, 80], [0.5, 1.0, 64, 80], [0.25, 0.25, 70, 80], [1.0, 2.0, 67, 100]])
    midi = numpy_to_midi(notes)  # the zero-length note is removed (remove_invalid_notes)
    assert [(n.pitch, n.start, n.end) for n in midi.instruments[0].notes] == [(60, 0.0, 0.5), (64, 0.5, 1.0), (67, 1.0, 2.0)]
    buf = io.BytesIO()
    midi.write(buf)
    tempo = bytes.fromhex("00ff5103" "07a120" "00ff5804" "04021808" "01ff2f00")
    piano = (bytes.fromhex("00ff0305") + b"Piano" + bytes.fromhex("00c000")
             + bytes.fromhex("00903c50")            # t=0     note on 60 vel 80
             + bytes.fromhex("8300803c00")          # +384    note off 60            (VLQ 384 = 83 00)
             + bytes.fromhex("00904050")            # +0      note on 64 (offs sort before ons on a tick)
             + bytes.fromhex("8300804000")          # +384    note off 64
             + bytes.fromhex("00904364")            # +0      note on 67 vel 100
             + bytes.fromhex("8600804300")          # +768    note off 67            (VLQ 768 = 86 00)
             + bytes.fromhex("01ff2f00"))
    want = (b"MThd" + bytes.fromhex("00000006" "0001" "0002" "0180")
            + b"MTrk" + len(tempo).to_bytes(4, "big") + tempo + b"MTrk" + len(piano).to_bytes(4, "big") + piano)
    assert buf.getvalue() == want
    back = PrettyMIDI(io.BytesIO(want))
    assert back.resolution == 384 and back.initial_tempo == 120.0
    assert [(n.pitch, n.velocity, n.start, n.end) for n in back.instruments[0].notes] == [
        (60, 80, 0.0, 0.5), (64, 80, 0.5, 1.0), (67, 100, 1.0, 2.0)]


def test_synthesis_entry_points_raise_a_clear_error():
    midi = numpy_to_midi(np.array([[0.0, 0.5, 60, 80]]))
    for call in (midi.fluidsynth, midi.synthesize, midi.instruments[0].fluidsynth):
        with pytest.raises(RuntimeError, match="pretty_midi"):
            call()
