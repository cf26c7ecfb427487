"""CPU tests of the drop-in Python API surface (names, signatures, state-dict layout, error
behaviour) mirroring how the reference's callers use it (demo.ipynb, webui.py, evaluate.py)."""
import os

import numpy as np
import pytest
import torch

from music2midi_b200 import synthetic as syn
from music2midi_b200.config import DEFAULT_CONFIG_PATH, load_config

REF_CFG = "/root/reference/config.yaml"


def test_alias_package_exports_reference_names():
    from music2midi.input import Conditioning, LogMelSpectrogram, ModelInputs  # noqa: F401
    from music2midi.model import Music2MIDI  # noqa: F401
    from music2midi.tokenizer import BOS, EOS, OFFSET, ONSET, PAD, MidiTokenizer  # noqa: F401
    from music2midi.transformer import T5Transformer  # noqa: F401
    from music2midi.utils import numpy_to_midi  # noqa: F401

    mi = ModelInputs(input_waveform=torch.zeros(1, 4))
    assert mi.notes_batch is None and mi.cond_index is None and mi._fields == ("input_waveform", "notes_batch", "cond_index")


@pytest.mark.parametrize("path", [DEFAULT_CONFIG_PATH] + ([REF_CFG] if os.path.exists(REF_CFG) else []))
def test_t5transformer_state_dict_is_reference_compatible(path, state_dict):
    from music2midi.transformer import T5Transformer

    m = T5Transformer(path)
    sd = m.state_dict()
    assert list(sd.keys()) == syn.state_dict_keys()
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(state_dict[k].shape), k
    assert sd["transformer.encoder.embed_tokens.weight"].data_ptr() == sd["transformer.shared.weight"].data_ptr()
    assert sd["transformer.lm_head.weight"].data_ptr() != sd["transformer.shared.weight"].data_ptr()
    assert torch.equal(sd["spectrogram.melspectrogram.mel_scale.fb"], state_dict["spectrogram.melspectrogram.mel_scale.fb"])
    m.load_state_dict(state_dict)
    assert torch.equal(m.transformer.lm_head.weight, state_dict["transformer.lm_head.weight"])
    for attr in ("config", "t5config", "transformer", "tokenizer", "spectrogram", "conditioning"):
        assert hasattr(m, attr)
    assert len(m.conditioning.embeds) == 2 and not m.training
    assert m.t5config.d_kv == 64 and m.t5config.num_heads == 8


def test_cpu_model_fails_loudly_instead_of_falling_back():
    from music2midi.input import LogMelSpectrogram, ModelInputs
    from music2midi.transformer import T5Transformer

    m = T5Transformer(DEFAULT_CONFIG_PATH)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.generate(ModelInputs(input_waveform=torch.zeros(1, 48000), cond_index=torch.zeros(1, 2).long()), max_length=4)
    with pytest.raises(RuntimeError, match="CUDA"):
        LogMelSpectrogram(16000, 2048, 256, 20.0, 384)(torch.zeros(1, 48000))
    with pytest.raises(NotImplementedError):
        m.generate(ModelInputs(torch.zeros(1, 48000), None, torch.zeros(1, 2).long()), num_beams=4)


def test_music2midi_argument_validation_and_checkpoint(tmp_path, state_dict):
    from music2midi.model import Music2MIDI

    m = Music2MIDI(DEFAULT_CONFIG_PATH)
    assert m.device.type == "cpu" and m.config.model.sample_rate == 16000 and hasattr(m, "model")
    with pytest.raises(ValueError, match="Either audio_path or audio_y"):
        m.generate()
    with pytest.raises(AssertionError):
        m.generate(audio_y=np.zeros(10, dtype=np.float32), sr=22050)
    # Lightning-style checkpoint: "model." prefix + hyper_parameters.config_path
    ck = tmp_path / "m.ckpt"
    torch.save({"state_dict": {"model." + k: v for k, v in state_dict.items()},
                "hyper_parameters": {"config_path": DEFAULT_CONFIG_PATH}}, ck)
    m2 = Music2MIDI.load_from_checkpoint(str(ck))
    assert torch.equal(m2.model.transformer.decoder.block[3].layer[1].EncDecAttention.k.weight,
                       state_dict["transformer.decoder.block.3.layer.1.EncDecAttention.k.weight"])
    m3 = Music2MIDI.load_from_checkpoint(str(ck), config_path=DEFAULT_CONFIG_PATH)
    assert torch.equal(m3.model.conditioning.embeds[1].weight, state_dict["conditioning.embeds.1.weight"])
    bad = tmp_path / "bad.ckpt"
    torch.save({"state_dict": {"model.nope": torch.zeros(1)}}, bad)
    with pytest.raises(RuntimeError, match="checkpoint mismatch"):
        Music2MIDI.load_from_checkpoint(str(bad), config_path=DEFAULT_CONFIG_PATH)


def test_load_audio_wav(tmp_path):
    import wave

    from music2midi_b200.model import load_audio

    sr = 32000
    t = np.arange(sr) / sr
    y = (0.5 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    p = tmp_path / "a.wav"
    with wave.open(str(p), "wb") as f:
        f.setnchannels(2)
        f.setsampwidth(2)
        f.setframerate(sr)
        st = np.stack([y, y], 1)
        f.writeframes((st * 32767).astype("<i2").tobytes())
    out = load_audio(p, 16000)
    assert out.dtype == np.float32 and abs(len(out) - 16000) <= 1
    assert abs(float(np.abs(out).max()) - 0.5) < 0.02


def test_config_loader_supports_reference_usage():
    cfg = load_config()
    assert cfg.model.t5.d_model == cfg["model"]["t5"]["d_model"] == 384
    assert dict(**cfg.spectrogram) == {"n_fft": 2048, "hop_length": 256, "f_min": 20.0}
    assert [len(v) for v in cfg.conditioning.values()] == [6, 3]


def test_resampler_accuracy_against_the_analytic_signal():
    """Audio ingest (reference model.py:84: librosa.load(path, sr=16000), i.e. soxr_hq resampling; neither librosa nor
    soxr is installed here, so parity with them is UNPINNED).  What can be stated is the accuracy of the polyphase
    resampler used instead: a band-limited test signal (tones below 0.4 x the target Nyquist) written as 44.1 kHz /
    48 kHz / 22.05 kHz 16-bit WAV comes back at 16 kHz within 2e-3 max abs (amplitude 0.5; 16-bit quantisation alone is
    3e-5) of the analytically evaluated signal, away from the edges of the filter."""
    import wave

    from music2midi_b200.model import load_audio

    freqs, amps = (110.0, 440.0, 1250.0, 3100.0), (0.2, 0.15, 0.1, 0.05)
    for sr_in in (44100, 48000, 22050):
        t = np.arange(sr_in * 2) / sr_in
        y = sum(a * np.sin(2 * np.pi * f * t + 0.3 * i) for i, (f, a) in enumerate(zip(freqs, amps)))
        import tempfile

        with tempfile.NamedTemporaryFile(suffix=".wav") as tmp:
            with wave.open(tmp.name, "wb") as f:
                f.setnchannels(1)
                f.setsampwidth(2)
                f.setframerate(sr_in)
                f.writeframes(np.round(y * 32767).astype("<i2").tobytes())
            out = load_audio(tmp.name, 16000)
        t16 = np.arange(len(out)) / 16000
        ref = sum(a * np.sin(2 * np.pi * f * t16 + 0.3 * i) for i, (f, a) in enumerate(zip(freqs, amps)))
        err = np.abs(out - ref)[800:-800].max()
        assert err <= 2e-3, (sr_in, err)
        assert abs(len(out) - 32000) <= 1
