"""Generates tests/golden/*.npz from the LIVE reference (run in the authoring container only).

    python tests/golden/make_golden.py

Imports the real ``T5Transformer`` / ``LogMelSpectrogram`` / ``MidiTokenizer`` classes from
/root/reference through oracle/reference_shim.py, loads the seeded synthetic state dict
(music2midi_b200/synthetic.py) into them, and records their outputs.  Inputs are NOT stored:
tests regenerate them from the recorded seeds and check a checksum recorded here.

The reference has no tests or golden vectors of its own; these fixtures are
"outputs of the reference itself run here".  Library versions at generation time are recorded
in each file (the reference pins torchaudio 2.1.0 / transformers 4.34.0; this image has
2.11.0 / 5.5.0 -- SURVEY.md §0.4).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from music2midi_b200 import synthetic as syn  # noqa: E402
from oracle import port, reference_shim as rs  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(os.cpu_count() or 1)


def versions():
    import torchaudio
    import transformers

    return np.array(
        [f"torch={torch.__version__}", f"torchaudio={torchaudio.__version__}", f"transformers={transformers.__version__}",
         f"numpy={np.__version__}"]
    )


def checksum(t: torch.Tensor) -> float:
    return float(t.double().abs().sum())


def load_reference(sd):
    inp, tr, tok = rs.import_reference()
    ref = tr.T5Transformer(rs.reference_config_path()).eval()
    own = ref.state_dict()
    assert list(own.keys()) == list(sd.keys())
    with torch.no_grad():
        for k, v in sd.items():
            if k != "transformer.lm_head.weight":
                own[k].copy_(v)
        # transformers>=5 force-ties lm_head to `shared`; config.yaml:23 wants it untied.
        ref.transformer.lm_head.weight = torch.nn.Parameter(sd["transformer.lm_head.weight"].clone())
    return ref, inp, tok


def candidate_inputs():
    wave = torch.cat([syn.audio_noise(8, 0), syn.audio_tones(8, 0)])
    cond = torch.stack([torch.arange(16) % 6, torch.arange(16) % 3], 1)
    return wave, cond


LOGIT_STEPS = list(range(32)) + list(range(100, 1001, 100)) + list(range(1015, 1023))


def make_mel(ref):
    cases = {}
    w_n = syn.audio_noise(2, 11)
    w_t = syn.audio_tones(2, 11)
    w_z = syn.audio_zeros(1)
    w_long = syn.audio_noise(1, 12, samples=66150)  # training shape 3 s @ 22050 Hz -> T=259
    w_short = syn.audio_tones(1, 13, samples=5000)  # ragged / short input -> T=20
    for name, w in [("noise", w_n), ("tones", w_t), ("zeros", w_z), ("long", w_long), ("short", w_short)]:
        with torch.no_grad():
            m = ref.spectrogram(w)
        f64 = port.logmel(w, syn.hann_window(), syn.mel_filterbank(), dtype=torch.float64)
        cases[f"{name}_mel"] = m.numpy()
        cases[f"{name}_insum"] = np.float64(checksum(w))
        cases[f"{name}_ref_vs_f64_maxabs"] = np.float64((m.double() - f64).abs().max())
    np.savez_compressed(os.path.join(OUT, "mel.npz"), versions=versions(), **cases)
    print("mel.npz", {k: v.shape for k, v in cases.items() if hasattr(v, "shape") and v.ndim > 0})


def make_generate(ref, inp, sd):
    wave, cond = candidate_inputs()
    W = port.Weights(sd)
    t0 = time.time()
    with torch.no_grad():
        toks = ref.generate(inp.ModelInputs(input_waveform=wave, cond_index=cond), max_length=1024)
        x = ref.conditioning(ref.spectrogram(wave), cond)
        enc = ref.transformer.encoder(inputs_embeds=x).last_hidden_state
        logits = ref.transformer(inputs_embeds=x, decoder_input_ids=toks[:, :-1]).logits  # [16, 1023, 400]
    print("reference generate 16x1024:", time.time() - t0, "s", toks.shape)
    assert toks.shape == (16, 1024)
    # the cached greedy path and the teacher-forced path agree on the argmax everywhere
    assert torch.equal(logits.argmax(-1), toks[:, 1:])
    top2 = logits.topk(2, dim=-1).values
    gap = top2[..., 0] - top2[..., 1]  # [16, 1023]
    min_gap = gap.min(1).values
    print("min top-2 gap per row:", min_gap)
    # the port agrees with the live reference (pins oracle/port.py)
    p_toks = port.generate(wave, cond, W, max_length=128)
    assert torch.equal(p_toks, toks[:, :128]), "oracle port diverges from the reference"
    np.savez_compressed(
        os.path.join(OUT, "generate.npz"),
        versions=versions(),
        weight_seed=0,
        insum=np.float64(checksum(wave)),
        cond=cond.numpy(),
        tokens=toks.numpy().astype(np.int16),
        gap=gap.numpy().astype(np.float32),
        logit_steps=np.array(LOGIT_STEPS),
        logits=logits[:, LOGIT_STEPS].numpy(),
        enc_rows=np.array([0, 9]),
        enc=enc[[0, 9]].numpy(),
    )

    # --- EOS case: lm_head row of EOS scaled x2 so rows finish at different steps ---------
    sd2 = dict(sd)
    lm = sd["transformer.lm_head.weight"].clone()
    lm[2] *= 2.0
    sd2["transformer.lm_head.weight"] = lm
    ref2, inp2, _ = load_reference(sd2)
    rows = [0, 2, 3, 4, 5, 6, 7]
    with torch.no_grad():
        t_all = ref2.generate(inp2.ModelInputs(input_waveform=wave[:8], cond_index=cond[:8]), max_length=1024)
        t_sub = ref2.generate(inp2.ModelInputs(input_waveform=wave[rows], cond_index=cond[rows]), max_length=1024)
        t_cap = ref2.generate(inp2.ModelInputs(input_waveform=wave[:8], cond_index=cond[:8]), max_length=40)
    print("eos case: all", t_all.shape, "subset", t_sub.shape, "capped", t_cap.shape)
    np.savez_compressed(
        os.path.join(OUT, "generate_eos.npz"),
        versions=versions(),
        weight_seed=0,
        eos_row_scale=2.0,
        rows=np.array(rows),
        tokens_all=t_all.numpy().astype(np.int16),
        tokens_subset=t_sub.numpy().astype(np.int16),
        tokens_cap40=t_cap.numpy().astype(np.int16),
    )
    return toks


def make_tokenizer(tok_mod, gen_tokens):
    inp, tr, tok = rs.import_reference()
    from omegaconf import OmegaConf

    tk = tok.MidiTokenizer(OmegaConf.load(rs.reference_config_path()))
    g = torch.Generator().manual_seed(1234)
    rows = []
    # (a) known answers from SURVEY.md §4
    rows.append([1, 133, 3, 65, 69, 143, 4, 65, 153, 4, 69, 3, 77, 2, 0, 0])
    rows.append([1, 138, 65, 3, 142, 4, 65, 2])
    rows.append([2])
    rows.append([1])
    rows.append([])
    # (b) grammar-shaped random streams: time, ONSET pitches, OFFSET pitches ...
    for _ in range(24):
        r = [1]
        t = 0
        for _ in range(int(torch.randint(1, 40, (1,), generator=g))):
            t += int(torch.randint(0, 12, (1,), generator=g))
            r.append(133 + min(t, 199))
            if torch.rand(1, generator=g) < 0.7:
                r.append(3)
                r += (torch.randint(40, 60, (int(torch.randint(1, 4, (1,), generator=g)),), generator=g) + 5).tolist()
            if torch.rand(1, generator=g) < 0.6:
                r.append(4)
                r += (torch.randint(40, 60, (int(torch.randint(1, 4, (1,), generator=g)),), generator=g) + 5).tolist()
        if torch.rand(1, generator=g) < 0.8:
            r.append(2)
        r += [0] * int(torch.randint(0, 5, (1,), generator=g))
        rows.append(r)
    # (c) garbage streams over the whole vocabulary, including ids 333..399 (never clamped)
    for _ in range(16):
        n = int(torch.randint(1, 300, (1,), generator=g))
        rows.append(torch.randint(0, 400, (n,), generator=g).tolist())
    # (d) real model output rows
    for r in gen_tokens[:6]:
        rows.append(r.tolist())

    L = max(len(r) for r in rows)
    lens = np.array([len(r) for r in rows])
    mat = np.zeros((len(rows), L), dtype=np.int16)
    for i, r in enumerate(rows):
        mat[i, : len(r)] = r
    batched = tk.decode([np.asarray(r, dtype=np.int64) for r in rows], mode="batched")
    seq = tk.decode([np.asarray(r, dtype=np.int64) for r in rows], mode="sequential", duration_per_batch=3)
    seq_cut = tk.decode([np.asarray(r, dtype=np.int64) for r in rows[:8]], mode="sequential", duration_per_batch=3,
                        cutoff_time=4)
    out = {f"batched_{i}": b for i, b in enumerate(batched)}
    # notes -> tokens (labels), incl. the round trip from SURVEY.md §4
    notes_cases = [
        np.array([[0, 0.5, 60, 80], [0, 1, 64, 80], [1, 1.5, 72, 80]], dtype=np.float64),
        np.zeros((0, 4)),
    ]
    for _ in range(10):
        n = int(torch.randint(1, 30, (1,), generator=g))
        on = torch.rand(n, generator=g).double() * 3.0
        dur = torch.rand(n, generator=g).double() * 1.0
        pitch = torch.randint(21, 109, (n,), generator=g).double()
        notes_cases.append(torch.stack([on, on + dur, pitch, torch.full((n,), 80.0).double()], 1).numpy())
    labels = tk(tuple(notes_cases)).numpy()
    labels_cut = tk(tuple(notes_cases), cutoff_time=2).numpy()
    strings = tk.to_string(np.array([0, 1, 2, 3, 4, 5, 132, 133, 332, 399]))
    for i, n in enumerate(notes_cases):
        out[f"notes_{i}"] = n
    np.savez_compressed(
        os.path.join(OUT, "tokenizer.npz"),
        versions=versions(), tokens=mat, lens=lens, n_rows=len(rows), sequential=seq, sequential_cut4=seq_cut,
        n_notes_cases=len(notes_cases), labels=labels.astype(np.int16), labels_cut2=labels_cut.astype(np.int16),
        strings=np.array(strings), **out,
    )
    print("tokenizer.npz rows", len(rows), "seq notes", seq.shape, "labels", labels.shape)
    return tk, notes_cases


def make_forward(ref, inp, notes_cases):
    """T5Transformer.forward (teacher-forced + CE loss), reference transformer.py:28-39."""
    wave = syn.audio_noise(3, 21)
    cond = torch.tensor([[1, 2], [0, 0], [4, 1]])
    notes = tuple(notes_cases[i] for i in (0, 3, 5))
    with torch.no_grad():
        out = ref(inp.ModelInputs(input_waveform=wave, notes_batch=notes, cond_index=cond))
    np.savez_compressed(
        os.path.join(OUT, "forward.npz"), versions=versions(), weight_seed=0, insum=np.float64(checksum(wave)),
        cond=cond.numpy(), notes_idx=np.array([0, 3, 5]), loss=np.float64(out.loss), logits=out.logits.numpy(),
    )
    print("forward.npz loss", float(out.loss), out.logits.shape)


def main():
    assert rs.available(), "needs /root/reference"
    sd = syn.synthetic_state_dict(0)
    ref, inp, tok = load_reference(sd)
    make_mel(ref)
    toks = make_generate(ref, inp, sd)
    tk, notes_cases = make_tokenizer(tok, toks)
    make_forward(ref, inp, notes_cases)


if __name__ == "__main__":
    main()
