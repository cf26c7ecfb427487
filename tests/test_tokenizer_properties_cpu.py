"""Property tests (hypothesis) of the token <-> note code: the C++ state machine behind MidiTokenizer against the
oracle restatement of the reference tokenizer (itself pinned to the reference's outputs in tests/golden), on
arbitrary token streams and note sets."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from music2midi_b200.config import load_config
from music2midi_b200.tokenizer import MidiTokenizer
from oracle import port

TK = MidiTokenizer(load_config())

tokens_rows = st.lists(st.lists(st.integers(0, 399), min_size=0, max_size=120), min_size=1, max_size=6)
# streams biased towards the grammar (time, ONSET/OFFSET, pitches) so that notes actually open and close
grammar = st.lists(
    st.one_of(st.integers(133, 200), st.sampled_from([3, 4]), st.integers(60, 72), st.sampled_from([0, 1, 2])),
    min_size=0, max_size=200)


@settings(max_examples=150, deadline=None)
@given(tokens_rows)
def test_decode_equals_oracle_on_arbitrary_streams(rows):
    rows = [np.asarray(r, dtype=np.int64) for r in rows]
    for a, b in zip(TK.decode(rows, mode="batched"), port.decode(rows, mode="batched")):
        assert a.dtype == np.float64 and np.array_equal(a, b)
    assert np.array_equal(TK.decode(rows, mode="sequential", duration_per_batch=3),
                          port.decode(rows, mode="sequential", duration_per_batch=3))
    assert np.array_equal(TK.decode(rows, mode="sequential", duration_per_batch=3, cutoff_time=5),
                          port.decode(rows, mode="sequential", duration_per_batch=3, cutoff_time=5))


@settings(max_examples=150, deadline=None)
@given(grammar, st.integers(0, 500))
def test_decode_equals_oracle_on_grammar_streams(row, start):
    row = np.asarray(row, dtype=np.int64)
    assert np.array_equal(TK._decode(row, start), port.decode_row(row, start))


notes_strategy = st.lists(
    st.tuples(st.integers(0, 190), st.integers(0, 60), st.integers(21, 108)), min_size=0, max_size=40)


@settings(max_examples=150, deadline=None)
@given(notes_strategy)
def test_tokenize_equals_oracle_and_round_trips(raw):
    notes = np.array([[on * 0.05, (on + dur) * 0.05, p, 80] for on, dur, p in raw], dtype=np.float64).reshape(-1, 4)
    mine = TK((notes,)).numpy()
    assert np.array_equal(mine, port.tokenize((notes,)).numpy())
    assert mine[0, -1] == 2  # EOS terminated
    # decoding the labels gives back every note whose (quantised) duration is positive and whose offset is in range
    back = TK.decode([mine[0]])[0]
    assert np.array_equal(back, port.decode([mine[0]])[0])
    assert back.shape[0] <= max(len(raw), 0)
