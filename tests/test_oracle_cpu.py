"""CPU tests that pin the oracle (oracle/port.py, oracle/hf_path.py) against the golden fixtures
recorded from the live reference, and — where /root/reference is present — against the live
reference classes themselves."""
import numpy as np
import pytest
import torch

from conftest import golden
from music2midi_b200 import synthetic as syn
from music2midi_b200.engine import relative_position_bucket
from oracle import port, reference_shim


def candidate_inputs():
    wave = torch.cat([syn.audio_noise(8, 0), syn.audio_tones(8, 0)])
    cond = torch.stack([torch.arange(16) % 6, torch.arange(16) % 3], 1)
    return wave, cond


def test_synthetic_state_dict_layout(state_dict):
    keys = list(state_dict.keys())
    assert len(keys) == 150 and keys == syn.state_dict_keys()
    assert state_dict["transformer.lm_head.weight"] is not state_dict["transformer.shared.weight"]
    assert state_dict["transformer.encoder.embed_tokens.weight"] is state_dict["transformer.shared.weight"]
    fb = state_dict["spectrogram.melspectrogram.mel_scale.fb"]
    assert fb.shape == (1025, 384) and int((fb != 0).sum()) == 2034 and float(fb.max()) == pytest.approx(0.9992, abs=1e-4)
    assert int((fb != 0).sum(0).max()) <= 14 and int((fb != 0).sum(0).min()) >= 1  # banded, no empty filter
    w = state_dict["spectrogram.melspectrogram.spectrogram.window"]
    assert float(w[0]) == 0.0 and float(w[1024]) == 1.0


def test_filterbank_and_window_equal_torchaudio():
    import torchaudio

    ms = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, hop_length=256, f_min=20.0, n_mels=384)
    assert torch.equal(ms.mel_scale.fb, syn.mel_filterbank())
    assert torch.equal(ms.spectrogram.window, syn.hann_window())


@pytest.mark.parametrize("case", ["noise", "tones", "zeros", "long", "short"])
def test_port_logmel_vs_golden(case):
    g = golden("mel.npz")
    wave = {"noise": lambda: syn.audio_noise(2, 11), "tones": lambda: syn.audio_tones(2, 11),
            "zeros": lambda: syn.audio_zeros(1), "long": lambda: syn.audio_noise(1, 12, samples=66150),
            "short": lambda: syn.audio_tones(1, 13, samples=5000)}[case]()
    assert float(wave.double().abs().sum()) == pytest.approx(float(g[f"{case}_insum"]), rel=1e-9)
    out = port.logmel(wave, syn.hann_window(), syn.mel_filterbank())
    ref = torch.from_numpy(g[f"{case}_mel"])
    assert out.shape == ref.shape
    # same formula, possibly a different FFT kernel on another host CPU: well inside 1e-4 normalised
    assert float((out - ref).abs().max() / ref.abs().max()) <= 1e-4
    if case == "zeros":
        assert float(ref.max()) == float(ref.min()) == pytest.approx(-13.815511, abs=1e-5)


def test_port_encoder_and_logits_vs_golden(oracle_weights):
    g = golden("generate.npz")
    wave, cond = candidate_inputs()
    assert float(wave.double().abs().sum()) == pytest.approx(float(g["insum"]), rel=1e-9)
    W = oracle_weights
    rows = g["enc_rows"].tolist()
    x = port.conditioning(port.logmel(wave[rows], W.window, W.fb), cond[rows], W.cond_embeds)
    enc = port.encoder(x, W)
    ref = torch.from_numpy(g["enc"])
    assert float((enc - ref).abs().max() / ref.abs().max()) <= 1e-4
    tokens = torch.from_numpy(g["tokens"].astype(np.int64))[rows]
    steps = [s for s in g["logit_steps"].tolist() if s < 100]
    logits = port.decoder(tokens[:, :100], enc, W)
    gl = torch.from_numpy(g["logits"])[rows][:, : len(steps)]
    assert float((logits[:, steps] - gl).abs().max()) <= 5e-4


def test_port_greedy_tokens_vs_golden(oracle_weights):
    g = golden("generate.npz")
    wave, cond = candidate_inputs()
    rows = [2, 9, 10, 13]  # golden top-2 gap >= 2e-3 along the whole path
    assert float(torch.from_numpy(g["gap"])[rows].min()) >= 2e-3
    out = port.generate(wave[rows], cond[rows], oracle_weights, max_length=80)
    assert torch.equal(out, torch.from_numpy(g["tokens"].astype(np.int64))[rows, :80])


def test_port_eos_semantics(state_dict, oracle_weights):
    g = golden("generate_eos.npz")
    sd = dict(state_dict)
    lm = sd["transformer.lm_head.weight"].clone()
    lm[2] *= float(g["eos_row_scale"])
    sd["transformer.lm_head.weight"] = lm
    W2 = port.Weights(sd)
    wave, cond = candidate_inputs()
    out = port.generate(wave[:8], cond[:8], W2, max_length=40)
    assert torch.equal(out, torch.from_numpy(g["tokens_cap40"].astype(np.int64)))


def test_hf_path_vs_golden(state_dict):
    from oracle import hf_path

    g = golden("generate.npz")
    wave, cond = candidate_inputs()
    rows = [2, 9]
    m = hf_path.build(state_dict)
    out = m.generate(wave[rows], cond[rows], max_length=24)
    assert torch.equal(out, torch.from_numpy(g["tokens"].astype(np.int64))[rows, :24])


def test_port_forward_loss_vs_golden(oracle_weights):
    g = golden("forward.npz")
    tk = golden("tokenizer.npz")
    wave = syn.audio_noise(3, 21)
    notes = tuple(tk[f"notes_{i}"] for i in g["notes_idx"].tolist())
    loss, logits = port.forward_loss(wave, notes, torch.from_numpy(g["cond"]), oracle_weights)
    assert logits.shape == g["logits"].shape
    assert float((logits - torch.from_numpy(g["logits"])).abs().max()) <= 5e-4
    assert float(loss) == pytest.approx(float(g["loss"]), abs=1e-4)


def test_port_tokenizer_vs_golden():
    g = golden("tokenizer.npz")
    rows = [g["tokens"][i, : g["lens"][i]].astype(np.int64) for i in range(int(g["n_rows"]))]
    batched = port.decode(rows, mode="batched")
    for i, b in enumerate(batched):
        assert b.dtype == np.float64 and np.array_equal(b, g[f"batched_{i}"]), i
    assert np.array_equal(port.decode(rows, mode="sequential", duration_per_batch=3), g["sequential"])
    assert np.array_equal(port.decode(rows[:8], mode="sequential", duration_per_batch=3, cutoff_time=4),
                          g["sequential_cut4"])
    notes = tuple(g[f"notes_{i}"] for i in range(int(g["n_notes_cases"])))
    assert np.array_equal(port.tokenize(notes).numpy(), g["labels"].astype(np.int64))
    assert np.array_equal(port.tokenize(notes, cutoff_time=2).numpy(), g["labels_cut2"].astype(np.int64))


def test_relative_position_buckets_known_answers():
    """SURVEY.md §4 table (measured from HF)."""
    enc = lambda r: int(relative_position_bucket(torch.tensor([r]), True))  # noqa: E731
    dec = lambda r: int(relative_position_bucket(torch.tensor([r]), False))  # noqa: E731
    assert [enc(-r) for r in range(8)] == list(range(8))
    for lo, hi, b in [(8, 11, 8), (12, 15, 9), (16, 22, 10), (23, 31, 11), (32, 45, 12), (46, 63, 13), (64, 90, 14),
                      (91, 500, 15)]:
        assert enc(-lo) == enc(-hi) == b and enc(lo) == enc(hi) == 16 + b
    assert enc(1) == 17 and enc(7) == 23
    assert [dec(-r) for r in range(16)] == list(range(16)) and dec(5) == 0
    for lo, hi, b in [(16, 18, 16), (19, 20, 17), (21, 23, 18), (24, 26, 19), (27, 30, 20), (31, 34, 21), (35, 39, 22),
                      (40, 45, 23), (46, 51, 24), (52, 58, 25), (59, 66, 26), (67, 76, 27), (77, 86, 28), (87, 98, 29),
                      (99, 112, 30), (113, 1023, 31)]:
        assert dec(-lo) == dec(-hi) == b, (lo, hi, b)
    # identical to the oracle's restatement
    rel = torch.arange(-600, 600)
    for bi in (True, False):
        assert torch.equal(relative_position_bucket(rel, bi), port.relative_position_bucket(rel, bi))


@pytest.mark.skipif(not reference_shim.available(), reason="/root/reference not present (GPU box)")
def test_live_reference_agrees_with_hf_path_and_port(state_dict, oracle_weights):
    from oracle import hf_path

    inp, tr, tok = reference_shim.import_reference()
    ref = tr.T5Transformer(reference_shim.reference_config_path()).eval()
    own = ref.state_dict()
    assert list(own.keys()) == list(state_dict.keys())
    with torch.no_grad():
        for k, v in state_dict.items():
            if k != "transformer.lm_head.weight":
                own[k].copy_(v)
        ref.transformer.lm_head.weight = torch.nn.Parameter(state_dict["transformer.lm_head.weight"].clone())
    wave = torch.cat([syn.audio_noise(1, 42), syn.audio_tones(1, 42)])
    cond = torch.tensor([[1, 2], [4, 0]])
    with torch.no_grad():
        r = ref.generate(inp.ModelInputs(input_waveform=wave, cond_index=cond), max_length=20)
        mel = ref.spectrogram(wave)
    h = hf_path.build(state_dict).generate(wave, cond, max_length=20)
    assert torch.equal(r, h)
    assert torch.equal(mel, hf_path.build().spectrogram(wave))
    p = port.generate(wave, cond, oracle_weights, max_length=20)
    assert torch.equal(r, p)
    assert float((port.logmel(wave, oracle_weights.window, oracle_weights.fb) - mel).abs().max()) < 1e-5
