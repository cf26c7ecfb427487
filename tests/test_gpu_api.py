"""GPU tests of the drop-in Python API (the calls demo.ipynb / webui.py / evaluate.py make), checked
against the CPU oracle and the golden fixtures recorded from the live reference."""
import numpy as np
import pytest
import torch

from conftest import golden
from music2midi_b200 import synthetic as syn
from music2midi_b200.config import DEFAULT_CONFIG_PATH
from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def m2m(state_dict):
    from music2midi.model import Music2MIDI

    m = Music2MIDI(DEFAULT_CONFIG_PATH)
    m.model.load_state_dict(state_dict)
    return m.to(DEV)


def test_music2midi_generate_matches_oracle_notes(m2m, oracle_weights):
    """Music2MIDI.generate(audio_y=...) (model.py:67-99): pad to a multiple of 3 s, split, greedy decode to
    1024 tokens per segment, sequential token->note decoding, MIDI object."""
    g = torch.Generator().manual_seed(5)
    audio = (0.1 * torch.randn(48000 * 2 + 12345, generator=g)).numpy()  # 2 full segments + a ragged tail
    midi = m2m.generate(audio_y=audio, cond_index=[2, 1])
    padded = np.pad(audio, (0, 3 * 48000 - len(audio)))
    notes, rows = port.sample_tokens(torch.from_numpy(padded), oracle_weights, cond_index=[2, 1])
    got = np.array([[n.start, n.end, n.pitch, n.velocity] for n in midi.instruments[0].notes]).reshape(-1, 4)
    exp = notes[notes[:, 1] > notes[:, 0]]  # remove_invalid_notes
    assert got.shape == exp.shape and np.array_equal(got, exp)
    assert midi.resolution == 384 and midi.instruments[0].name == "Piano"
    # token level too
    toks = m2m.generate_tokens(torch.from_numpy(padded).to(DEV), 48000, cond_index=[2, 1])
    assert all(torch.equal(a.cpu(), b) for a, b in zip(toks, rows))
    # sample_tokens with cond_index=None == zeros
    n0 = m2m.sample_tokens(torch.from_numpy(padded[:48000]).to(DEV), 48000, 3)
    e0, _ = port.sample_tokens(torch.from_numpy(padded[:48000]), oracle_weights)
    assert np.array_equal(n0, e0)


def test_t5transformer_forward_matches_reference(m2m):
    from music2midi.input import ModelInputs

    g = golden("forward.npz")
    tk = golden("tokenizer.npz")
    notes = tuple(tk[f"notes_{i}"] for i in g["notes_idx"].tolist())
    out = m2m.model(ModelInputs(input_waveform=syn.audio_noise(3, 21).to(DEV), notes_batch=notes,
                                cond_index=torch.from_numpy(g["cond"]).to(DEV)))
    assert float((out.logits.cpu() - torch.from_numpy(g["logits"])).abs().max()) <= 2e-3
    assert abs(float(out.loss) - float(g["loss"])) <= 1e-4
    assert out["loss"] is out.loss


def test_t5transformer_generate_kwargs(m2m):
    from music2midi.input import ModelInputs

    g = golden("generate.npz")
    wave = torch.cat([syn.audio_noise(8, 0), syn.audio_tones(8, 0)])[[2, 9]]
    cond = torch.from_numpy(g["cond"])[[2, 9]]
    exp = torch.from_numpy(g["tokens"].astype(np.int64))[[2, 9]]
    mi = ModelInputs(input_waveform=wave.to(DEV), cond_index=cond.to(DEV))
    out = m2m.model.generate(mi, max_length=50)
    assert out.device.type == "cuda" and out.dtype == torch.int64 and torch.equal(out.cpu(), exp[:, :50])
    assert m2m.model.generate(mi).shape == (2, 20)  # HF default max_length
    # the HF-style call of the reference (transformer.py:42-44) on the parameter container
    emb = m2m.model.conditioning(m2m.model.spectrogram(mi.input_waveform), mi.cond_index)
    assert torch.equal(m2m.model.transformer.generate(inputs_embeds=emb, max_length=50).cpu(), exp[:, :50])
    assert torch.equal(m2m.model.generate(mi, max_new_tokens=9).cpu(), exp[:, :10])
    # weights edited in place are picked up (engine re-upload keyed on parameter versions)
    with torch.no_grad():
        m2m.model.transformer.lm_head.weight[2] *= 2.0
    eos = golden("generate_eos.npz")
    out2 = m2m.model.generate(ModelInputs(torch.cat([syn.audio_noise(8, 0)]).to(DEV),
                                          None, torch.from_numpy(g["cond"])[:8].to(DEV)), max_length=40)
    with torch.no_grad():
        m2m.model.transformer.lm_head.weight[2] /= 2.0
    assert torch.equal(out2.cpu(), torch.from_numpy(eos["tokens_cap40"].astype(np.int64)))


def test_standalone_modules(state_dict):
    from music2midi.input import Conditioning, LogMelSpectrogram

    spec = LogMelSpectrogram(sample_rate=16000, n_fft=2048, hop_length=256, f_min=20.0, n_mels=384).to(DEV)
    assert list(spec.state_dict().keys()) == ["melspectrogram.spectrogram.window", "melspectrogram.mel_scale.fb"]
    w = syn.audio_noise(2, 11)
    out = spec(w.to(DEV))
    ref = torch.from_numpy(golden("mel.npz")["noise_mel"])
    assert out.shape == ref.shape and float((out.cpu() - ref).abs().max() / ref.abs().max()) <= 1e-4
    cond = Conditioning(384, [6, 3]).to(DEV)
    feat = torch.randn(3, 5, 384, device=DEV)
    idx = torch.tensor([[0, 0], [5, 2], [1, 1]], device=DEV)
    got = cond(feat, idx)
    exp = torch.cat([torch.stack([cond.embeds[0].weight[idx[:, 0]], cond.embeds[1].weight[idx[:, 1]]], 1), feat], 1)
    assert torch.equal(got, exp)


def test_long_teacher_forced_decoder_matches_oracle(engine_fp32, oracle_weights, report):
    """Decoder length > 256 exercises the key-tiled causal attention (config 3 uses up to 1024)."""
    g = torch.Generator().manual_seed(9)
    enc = torch.randn(2, 190, 384, generator=g)
    dec_in = torch.randint(0, 400, (2, 700), generator=g)
    dec_in[:, 0] = 1
    ref = port.decoder(dec_in, enc, oracle_weights)
    out = engine_fp32.decoder_forward(enc.to(DEV), dec_in.to(DEV)).cpu()
    d = float((out - ref).abs().max())
    report(test="decoder_forward_long_fp32", Ld=700, max_abs=d)
    assert d <= 2e-3


@pytest.mark.parametrize("Ld", [1, 100, 128, 129, 700, 1024])
def test_bf16_teacher_forced_decoder_tcgen05_attention(engine_bf16, oracle_weights, report, Ld):
    """bf16 teacher-forced decoder: key-tiled fused tcgen05 attention (causal self-attention with the bucket-bias
    LUT, cross-attention over head-major encoder K/V) against the fp32 oracle and against the CUDA-core kernel."""
    g = torch.Generator().manual_seed(10 + Ld)
    enc = torch.randn(2, 190, 384, generator=g)
    dec_in = torch.randint(0, 400, (2, Ld), generator=g)
    dec_in[:, 0] = 1
    ref = port.decoder(dec_in, enc, oracle_weights)
    out = engine_bf16.decoder_forward(enc.to(DEV), dec_in.to(DEV)).cpu()
    engine_bf16.set_flags(no_tc_attention=True)
    out2 = engine_bf16.decoder_forward(enc.to(DEV), dec_in.to(DEV)).cpu()
    engine_bf16.set_flags()
    d, d2 = (out - ref).abs(), (out2 - ref).abs()
    report(test="decoder_forward_bf16", Ld=Ld, tc_max=float(d.max()), tc_mean=float(d.mean()), simt_max=float(d2.max()),
           simt_mean=float(d2.mean()), tc_vs_simt_max=float((out - out2).abs().max()))
    assert float(d.max()) <= 0.25 and float(d.mean()) <= 0.04


def test_generate_many_equals_per_recording_generate(m2m):
    """Batch entry point: several recordings of different lengths as one device batch == one call per recording."""
    g = torch.Generator().manual_seed(77)
    audios = [(0.1 * torch.randn(n, generator=g)).numpy() for n in (48000, 100000, 20000)]
    many = m2m.generate_many(audios, cond_index=[1, 0])
    assert len(many) == 3
    for y, midi in zip(audios, many):
        single = m2m.generate(audio_y=y, cond_index=[1, 0])
        a = [(n.start, n.end, n.pitch, n.velocity) for n in midi.instruments[0].notes]
        b = [(n.start, n.end, n.pitch, n.velocity) for n in single.instruments[0].notes]
        assert a == b
    assert m2m.generate_many([]) == []


def test_checkpoint_to_gpu_inference(state_dict, tmp_path):
    """webui.py:90 / demo.ipynb:14: Music2MIDI.load_from_checkpoint(ckpt, config_path=...).cuda() -> generate.  A
    Lightning-style checkpoint ("model." prefix, hyper_parameters) written from the synthetic weights must reproduce
    the reference's golden tokens through the CUDA path; one without lm_head.weight (an HF >= 5 export, where the
    head is force-tied) must fail loudly instead of silently running with a random head."""
    from music2midi.input import ModelInputs
    from music2midi.model import Music2MIDI

    ck = tmp_path / "m.ckpt"
    torch.save({"state_dict": {"model." + k: v for k, v in state_dict.items()},
                "hyper_parameters": {"config_path": DEFAULT_CONFIG_PATH}}, ck)
    m = Music2MIDI.load_from_checkpoint(str(ck)).cuda()
    assert m.device.type == "cuda"
    g = golden("generate.npz")
    wave = torch.cat([syn.audio_noise(8, 0), syn.audio_tones(8, 0)])
    cond = torch.stack([torch.arange(16) % 6, torch.arange(16) % 3], 1)
    toks = m.model.generate(ModelInputs(input_waveform=wave.to(DEV), cond_index=cond.to(DEV)), max_length=160)
    assert torch.equal(toks.cpu(), torch.from_numpy(g["tokens"].astype(np.int64))[:, :160])
    # the public entry point on top of it
    midi = m.generate(audio_y=wave[0].numpy(), cond_index=[0, 0])
    assert midi.resolution == 384 and len(midi.instruments) == 1

    bad = tmp_path / "no_head.ckpt"
    torch.save({"state_dict": {"model." + k: v for k, v in state_dict.items() if k != "transformer.lm_head.weight"},
                "hyper_parameters": {"config_path": DEFAULT_CONFIG_PATH}}, bad)
    with pytest.raises(RuntimeError, match="lm_head"):
        Music2MIDI.load_from_checkpoint(str(bad))


def test_generate_many_uses_the_host_buffer_path(m2m):
    """generate_many stages all recordings in one pinned buffer and runs m2m_transcribe_host (double-buffered
    upload, int16 token read-back): same notes as transcribing the recordings one by one."""
    g = torch.Generator().manual_seed(11)
    recs = [(0.1 * torch.randn(n, generator=g)).numpy() for n in (48000, 48000 * 2 + 777, 30000)]
    many = m2m.generate_many(recs, cond_index=[1, 2])
    for rec, mm in zip(recs, many):
        one = m2m.generate(audio_y=rec, cond_index=[1, 2])
        a = [(n.start, n.end, n.pitch, n.velocity) for n in one.instruments[0].notes]
        b = [(n.start, n.end, n.pitch, n.velocity) for n in mm.instruments[0].notes]
        assert a == b
