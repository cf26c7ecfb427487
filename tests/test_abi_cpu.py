"""CPU tests of the boundary: the C-ABI library loads and exports every symbol of include/m2m_b200.h,
fails loudly without a GPU, and the CPU-side integer work (token -> notes) matches the reference's
tokenizer on the golden fixtures.  No device compute is called here."""
import ctypes as C
import io
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, golden
from music2midi_b200 import _lib
from music2midi_b200.config import load_config
from music2midi_b200.tokenizer import BOS, EOS, OFFSET, ONSET, PAD, MidiTokenizer


@pytest.fixture(scope="module")
def lib():
    from music2midi_b200 import build

    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "m2m_b200.h")).read()
    declared = set(re.findall(r"\b(m2m_[a-z0-9_]+)\s*\(", header))
    declared -= {"m2m_ctx"}
    assert len(declared) >= 18
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.m2m_abi_version() == 1


def test_default_config_matches_config_yaml(lib):
    cfg = _lib.default_config()
    y = load_config()
    t5 = y.model.t5
    assert (cfg.n_layers, cfg.d_model, cfg.d_ff, cfg.vocab) == (t5.num_layers, t5.d_model, t5.d_ff, t5.vocab_size)
    assert (cfg.n_fft, cfg.hop) == (y.spectrogram.n_fft, y.spectrogram.hop_length)
    assert (cfg.pad_id, cfg.bos_id, cfg.eos_id) == (t5.pad_token_id, t5.decoder_start_token_id, t5.eos_token_id)
    assert cfg.max_positions == t5.n_positions and cfg.n_cond == len(y.conditioning)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly(lib):
    assert lib.m2m_device_count() == 0
    cfg = _lib.default_config()
    ctx = C.c_void_p()
    rc = lib.m2m_ctx_create(C.byref(cfg), 0, C.byref(ctx))
    assert rc == 2 and not ctx.value  # M2M_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.m2m_last_error()
    from music2midi_b200.engine import Engine

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(torch.device("cpu"))
    with pytest.raises(_lib.M2MError):
        Engine(torch.device("cuda", 0))


def test_bad_config_rejected(lib):
    cfg = _lib.default_config()
    cfg.d_kv = 32
    ctx = C.c_void_p()
    assert lib.m2m_ctx_create(C.byref(cfg), 0, C.byref(ctx)) == 1
    assert b"d_kv" in lib.m2m_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "music2midi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "/root/reference" not in src, f


# ------------------------------------------------------------------ tokenizer (C++ state machine)
@pytest.fixture(scope="module")
def tk(lib):
    return MidiTokenizer(load_config())


def test_vocabulary_constants(tk):
    assert (PAD, BOS, EOS, ONSET, OFFSET) == (0, 1, 2, 3, 4)
    assert tk.pitch_token_offset == 5 and tk.time_token_offset == 133 and tk.time_step == 0.05
    g = golden("tokenizer.npz")
    assert tk.to_string(np.array([0, 1, 2, 3, 4, 5, 132, 133, 332, 399])) == g["strings"].tolist()
    with pytest.raises(ValueError):
        tk.to_string([-1])


def test_decode_known_answers(tk):
    notes = tk.decode([np.array([1, 133, 3, 65, 69, 143, 4, 65, 153, 4, 69, 3, 77, 2, 0, 0])])[0]
    assert notes.tolist() == [[0, 0.5, 60, 80], [0, 1.0, 64, 80]]  # the un-closed note 72 is dropped
    quirk = np.array([1, 138, 65, 3, 142, 4, 65, 2])  # pitch BEFORE the ONSET marker
    assert tk.decode([quirk])[0].tolist() == [[0.25, 0.45, 60, 80]]
    seq = tk.decode([quirk, quirk], mode="sequential", duration_per_batch=3)
    assert seq.tolist() == [[0.25, 0.45, 60, 80], [3.25, 3.45, 60, 80]]
    with pytest.raises(ValueError):
        tk.decode([quirk], mode="nope")
    with pytest.raises(AssertionError):
        tk.decode([quirk], mode="sequential")


def test_decode_matches_reference_fixture(tk):
    g = golden("tokenizer.npz")
    rows = [g["tokens"][i, : g["lens"][i]].astype(np.int64) for i in range(int(g["n_rows"]))]
    for i, b in enumerate(tk.decode(rows, mode="batched")):
        assert b.dtype == np.float64 and b.shape[1] == 4
        assert np.array_equal(b, g[f"batched_{i}"]), i
    assert np.array_equal(tk.decode(rows, mode="sequential", duration_per_batch=3), g["sequential"])
    assert np.array_equal(tk.decode(rows[:8], mode="sequential", duration_per_batch=3, cutoff_time=4),
                          g["sequential_cut4"])
    # torch tensors (as produced by generate) decode identically
    t_rows = [torch.from_numpy(r) for r in rows]
    assert np.array_equal(tk.decode(t_rows, mode="sequential", duration_per_batch=3), g["sequential"])


def test_tokenize_matches_reference_fixture(tk):
    g = golden("tokenizer.npz")
    notes = tuple(g[f"notes_{i}"] for i in range(int(g["n_notes_cases"])))
    before = [n.copy() for n in notes]
    labels = tk(notes)
    assert labels.dtype == torch.int64 and np.array_equal(labels.numpy(), g["labels"].astype(np.int64))
    assert np.array_equal(tk(notes, cutoff_time=2).numpy(), g["labels_cut2"].astype(np.int64))
    assert all(np.array_equal(a, b) for a, b in zip(notes, before))  # inputs are not mutated
    rt = tk.decode([labels[0]])[0]
    assert rt.tolist() == [[0, 0.5, 60, 80], [0, 1, 64, 80], [1, 1.5, 72, 80]]
    with pytest.raises(AssertionError):
        tk(5)


def test_tokens_to_notes_capacity_error(lib):
    toks = np.array([133, 3, 65, 66, 67], dtype=np.int64)
    rows = np.empty((1, 4), dtype=np.int64)
    n = C.c_int64(0)
    rc = lib.m2m_tokens_to_notes(toks.ctypes.data_as(C.c_void_p), toks.size, 0, 5, 133, 80,
                                 rows.ctypes.data_as(C.c_void_p), 1, C.byref(n))
    assert rc == 1 and b"capacity" in lib.m2m_last_error()


def test_many_open_notes_decode_is_linear_time(tk):
    # 20k onsets then 20k offsets: the reference's vstack/np.where path is quadratic; this must be fast
    on = [133, 3] + [5 + (i % 128) for i in range(20000)]
    off = [134, 4] + [5 + i for i in range(128)]
    notes = tk.decode([np.array(on + off + [2])])[0]
    assert notes.shape == (20000, 4) and np.all(notes[:, 1] == 0.05)


# ------------------------------------------------------------------ notes -> MIDI
def test_numpy_to_midi_and_smf_round_trip():
    from music2midi_b200 import midi
    from music2midi_b200.utils import numpy_to_midi

    notes = np.array([[0.0, 0.5, 60, 80], [0.0, 1.0, 64, 80], [1.0, 1.0, 72, 80], [2.0, 1.5, 50, 80]])
    m = numpy_to_midi(notes)
    assert m.resolution == 384 and len(m.instruments) == 1
    inst = m.instruments[0]
    assert inst.program == 0 and inst.name == "Piano" and not inst.is_drum
    assert [(n.start, n.end, n.pitch, n.velocity) for n in inst.notes] == [(0.0, 0.5, 60, 80), (0.0, 1.0, 64, 80)]
    assert m.get_end_time() == 1.0
    buf = io.BytesIO()
    m.write(buf)
    data = buf.getvalue()
    assert data[:4] == b"MThd" and data[12:14] == (384).to_bytes(2, "big")
    buf.seek(0)
    back = midi.PrettyMIDI(buf)
    assert [(n.start, n.end, n.pitch, n.velocity) for n in back.instruments[0].notes] == \
        [(0.0, 0.5, 60, 80), (0.0, 1.0, 64, 80)]
    assert back.initial_tempo == pytest.approx(120.0)
