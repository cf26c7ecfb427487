"""TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (music2midi_b200/) never does.

The reference (ytinyui/music2midi) authors no arithmetic of its own on this path: it calls
torchaudio 2.1.0 ``MelSpectrogram`` (music2midi/input.py:25-31,39) and HF transformers 4.34.0
``T5ForConditionalGeneration`` + ``generate`` (music2midi/transformer.py:16,35-37,44); neither
is vendored under /root/reference.  This file restates their published algorithms in plain
fp32 torch CPU ops (no torchaudio, no transformers), each function citing the reference call
site and the third-party source it follows.  Paths starting ``site-packages/`` refer to the
installed torchaudio 2.11.0 / transformers 5.5.0.

PARITY PINNING: the reference has no tests or golden vectors of its own, so
the port is pinned against the *live* reference classes imported from /root/reference through
oracle/reference_shim.py (tests/test_oracle_cpu.py, when that tree is present) and against
fixtures generated from the live reference by tests/golden/make_golden.py.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

PAD, BOS, EOS, ONSET, OFFSET = 0, 1, 2, 3, 4  # music2midi/tokenizer.py:11-15

N_LAYERS = 6
N_HEADS = 8
D_KV = 64
N_BUCKETS = 32
MAX_DISTANCE = 128
LN_EPS = 1e-6


# ============================================================================ log-mel
def logmel(wave: torch.Tensor, window: torch.Tensor, fb: torch.Tensor, hop: int = 256, dtype=torch.float32):
    """music2midi/input.py:33-41 -> torchaudio Spectrogram(power=2, center=True, reflect,
    periodic Hann, normalized=False) -> MelScale (matmul with fb) -> transpose -> clamp(1e-6).log().

    Follows site-packages/torchaudio/functional/functional.py:123-144 (torch.stft) and
    site-packages/torchaudio/transforms/_transforms.py:417 (``(spec.T @ fb).T``).
    wave [..., S] -> [..., 1 + S//hop, n_mels].  ``dtype=torch.float64`` gives the
    "formula in double" evaluation used to state the tolerance norm.
    """
    n_fft = window.numel()
    x = wave.to(dtype)
    lead = x.shape[:-1]
    x = x.reshape(-1, x.shape[-1])
    pad = n_fft // 2
    x = F.pad(x.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)  # center=True, pad_mode="reflect"
    frames = x.unfold(-1, n_fft, hop)  # [B, T, n_fft]
    spec = torch.fft.rfft(frames * window.to(dtype), dim=-1)  # onesided
    power = spec.real ** 2 + spec.imag ** 2  # |.|^power with power=2
    mel = power @ fb.to(dtype)  # [B, T, n_mels]
    out = mel.clamp(min=1e-6).log()
    return out.reshape(*lead, out.shape[-2], out.shape[-1])


def conditioning(feature: torch.Tensor, indices: torch.Tensor, embeds: Sequence[torch.Tensor]):
    """music2midi/input.py:50-59: embedding rows stacked on dim 1 and PREPENDED."""
    rows = [embeds[i][indices[:, i]] for i in range(len(embeds))]
    return torch.cat([torch.stack(rows, dim=1), feature], dim=1)


# ============================================================================ T5 pieces
def relative_position_bucket(rel: torch.Tensor, bidirectional: bool) -> torch.Tensor:
    """site-packages/transformers/models/t5/modeling_t5.py:187-234 (rel = key - query)."""
    num_buckets = N_BUCKETS
    ret = torch.zeros_like(rel)
    if bidirectional:
        num_buckets //= 2
        ret = ret + (rel > 0).to(torch.long) * num_buckets
        rel = rel.abs()
    else:
        rel = -torch.clamp(rel, max=0)
    max_exact = num_buckets // 2
    is_small = rel < max_exact
    large = max_exact + (
        torch.log(rel.float() / max_exact) / math.log(MAX_DISTANCE / max_exact) * (num_buckets - max_exact)
    ).to(torch.long)
    large = torch.clamp(large, max=num_buckets - 1)
    return ret + torch.where(is_small, rel, large)


def position_bias(table: torch.Tensor, q_pos: torch.Tensor, k_len: int, bidirectional: bool) -> torch.Tensor:
    """modeling_t5.py:236-251 compute_bias: [H, len(q_pos), k_len] from Embedding(32, H)."""
    rel = torch.arange(k_len)[None, :] - q_pos[:, None]
    return table[relative_position_bucket(rel, bidirectional)].permute(2, 0, 1)


def rmsnorm(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """modeling_t5.py:46-68 T5LayerNorm: no mean subtraction, no bias."""
    var = x.pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + LN_EPS))


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    """transformers activations NewGELUActivation (dense_act_fn 'gelu_new' for gated-gelu)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def _heads(x: torch.Tensor) -> torch.Tensor:
    b, l, _ = x.shape
    return x.view(b, l, N_HEADS, D_KV).transpose(1, 2)


def attention(q, k, v, bias) -> torch.Tensor:
    """modeling_t5.py:308-338: scores = q k^T (NO 1/sqrt(d) scale) + bias; fp32 softmax; @ v."""
    scores = torch.matmul(q, k.transpose(-1, -2))
    if bias is not None:
        scores = scores + bias
    p = torch.softmax(scores.float(), dim=-1)
    o = torch.matmul(p, v)
    b, h, l, d = o.shape
    return o.transpose(1, 2).reshape(b, l, h * d)


def ffn(x, wi0, wi1, wo):
    """modeling_t5.py:106-132 T5DenseGatedActDense."""
    return (gelu_new(x @ wi0.T) * (x @ wi1.T)) @ wo.T


class Weights:
    """Thin view over the reference state dict (keys in SURVEY.md §5)."""

    def __init__(self, sd: Dict[str, torch.Tensor]):
        self.sd = {k: v.detach().to(torch.float32) for k, v in sd.items()}

    def __getitem__(self, k):
        return self.sd["transformer." + k]

    @property
    def window(self):
        return self.sd["spectrogram.melspectrogram.spectrogram.window"]

    @property
    def fb(self):
        return self.sd["spectrogram.melspectrogram.mel_scale.fb"]

    @property
    def cond_embeds(self):
        return [self.sd["conditioning.embeds.0.weight"], self.sd["conditioning.embeds.1.weight"]]


def encoder(x: torch.Tensor, W: Weights) -> torch.Tensor:
    """modeling_t5.py T5Stack (encoder): 6 blocks + final RMSNorm; bias from block 0 reused."""
    L = x.shape[1]
    bias = position_bias(
        W["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"], torch.arange(L), L, True
    )[None]
    for l in range(N_LAYERS):
        p = f"encoder.block.{l}.layer"
        h = rmsnorm(x, W[f"{p}.0.layer_norm.weight"])
        a = f"{p}.0.SelfAttention"
        q, k, v = (_heads(h @ W[f"{a}.{n}.weight"].T) for n in "qkv")
        x = x + attention(q, k, v, bias) @ W[f"{a}.o.weight"].T
        h = rmsnorm(x, W[f"{p}.1.layer_norm.weight"])
        d = f"{p}.1.DenseReluDense"
        x = x + ffn(h, W[f"{d}.wi_0.weight"], W[f"{d}.wi_1.weight"], W[f"{d}.wo.weight"])
    return rmsnorm(x, W["encoder.final_layer_norm.weight"])


def cross_kv(enc: torch.Tensor, W: Weights):
    """Cross-attention K/V computed once from the encoder output (modeling_t5.py:281-305)."""
    out = []
    for l in range(N_LAYERS):
        a = f"decoder.block.{l}.layer.1.EncDecAttention"
        out.append((_heads(enc @ W[f"{a}.k.weight"].T), _heads(enc @ W[f"{a}.v.weight"].T)))
    return out


def decoder(tokens: torch.Tensor, enc: torch.Tensor, W: Weights, ckv=None) -> torch.Tensor:
    """Full-sequence causal decoder -> logits [B, L, V] (teacher-forced shape, a10).

    lm_head without the d_model**-0.5 rescale (tie_word_embeddings false; modeling_t5.py:1107).
    """
    B, L = tokens.shape
    x = W["shared.weight"][tokens]
    bias = position_bias(
        W["decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"], torch.arange(L), L, False
    )[None]
    causal = torch.full((L, L), float("-inf")).triu(1)
    bias = bias + causal
    if ckv is None:
        ckv = cross_kv(enc, W)
    for l in range(N_LAYERS):
        p = f"decoder.block.{l}.layer"
        h = rmsnorm(x, W[f"{p}.0.layer_norm.weight"])
        a = f"{p}.0.SelfAttention"
        q, k, v = (_heads(h @ W[f"{a}.{n}.weight"].T) for n in "qkv")
        x = x + attention(q, k, v, bias) @ W[f"{a}.o.weight"].T
        h = rmsnorm(x, W[f"{p}.1.layer_norm.weight"])
        a = f"{p}.1.EncDecAttention"
        q = _heads(h @ W[f"{a}.q.weight"].T)
        x = x + attention(q, ckv[l][0], ckv[l][1], None) @ W[f"{a}.o.weight"].T
        h = rmsnorm(x, W[f"{p}.2.layer_norm.weight"])
        d = f"{p}.2.DenseReluDense"
        x = x + ffn(h, W[f"{d}.wi_0.weight"], W[f"{d}.wi_1.weight"], W[f"{d}.wo.weight"])
    x = rmsnorm(x, W["decoder.final_layer_norm.weight"])
    return x @ W["lm_head.weight"].T


class DecoderState:
    """KV-cached single-step decoder (a8)."""

    def __init__(self, enc: torch.Tensor, W: Weights):
        self.W = W
        self.ckv = cross_kv(enc, W)
        self.k: List[Optional[torch.Tensor]] = [None] * N_LAYERS
        self.v: List[Optional[torch.Tensor]] = [None] * N_LAYERS
        self.t = 0
        self.table = W["decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]

    def step(self, tok: torch.Tensor) -> torch.Tensor:
        """tok int64 [B] -> logits fp32 [B, V]."""
        W = self.W
        x = W["shared.weight"][tok][:, None, :]
        bias = position_bias(self.table, torch.tensor([self.t]), self.t + 1, False)[None]
        for l in range(N_LAYERS):
            p = f"decoder.block.{l}.layer"
            h = rmsnorm(x, W[f"{p}.0.layer_norm.weight"])
            a = f"{p}.0.SelfAttention"
            q, k, v = (_heads(h @ W[f"{a}.{n}.weight"].T) for n in "qkv")
            self.k[l] = k if self.k[l] is None else torch.cat([self.k[l], k], dim=2)
            self.v[l] = v if self.v[l] is None else torch.cat([self.v[l], v], dim=2)
            x = x + attention(q, self.k[l], self.v[l], bias) @ W[f"{a}.o.weight"].T
            h = rmsnorm(x, W[f"{p}.1.layer_norm.weight"])
            a = f"{p}.1.EncDecAttention"
            q = _heads(h @ W[f"{a}.q.weight"].T)
            x = x + attention(q, self.ckv[l][0], self.ckv[l][1], None) @ W[f"{a}.o.weight"].T
            h = rmsnorm(x, W[f"{p}.2.layer_norm.weight"])
            d = f"{p}.2.DenseReluDense"
            x = x + ffn(h, W[f"{d}.wi_0.weight"], W[f"{d}.wi_1.weight"], W[f"{d}.wo.weight"])
        x = rmsnorm(x, W["decoder.final_layer_norm.weight"])
        self.t += 1
        return (x @ W["lm_head.weight"].T)[:, 0, :]


def greedy_generate(enc: torch.Tensor, W: Weights, max_length: int, forced: Optional[torch.Tensor] = None,
                    return_logits: bool = False):
    """HF greedy loop semantics (site-packages/transformers/generation/utils.py:2727-2809, a9).

    start token 1; next = argmax(logits); finished rows emit pad(0); finish on EOS(2); stop when
    all rows finished or length == max_length; output length is dynamic.
    ``forced`` [B, L] replaces the argmax feedback (teacher forcing through the cached path).
    """
    B = enc.shape[0]
    st = DecoderState(enc, W)
    out = torch.full((B, 1), BOS, dtype=torch.long)
    unfinished = torch.ones(B, dtype=torch.long)
    logits_all = []
    while out.shape[1] < max_length:
        logits = st.step(out[:, -1])
        if return_logits:
            logits_all.append(logits)
        nxt = torch.argmax(logits, dim=-1)
        nxt = nxt * unfinished + PAD * (1 - unfinished)
        if forced is not None:
            nxt = forced[:, out.shape[1]]
        out = torch.cat([out, nxt[:, None]], dim=1)
        unfinished = unfinished & (nxt != EOS).long()
        if forced is None and unfinished.max() == 0:
            break
    if return_logits:
        return out, torch.stack(logits_all, dim=1)
    return out


def generate(wave: torch.Tensor, cond_index: torch.Tensor, W: Weights, max_length: int = 1024) -> torch.Tensor:
    """music2midi/transformer.py:41-45."""
    x = conditioning(logmel(wave, W.window, W.fb), cond_index, W.cond_embeds)
    return greedy_generate(encoder(x, W), W, max_length)


# ============================================================================ tokenizer
TIME_STEP = 0.05  # midi_quantize_ms / 1000 (config.yaml:33)
PITCH_OFFSET = 5  # vocab_size.special
TIME_OFFSET = 133  # special + pitch
DEFAULT_VELOCITY = 80
N_TIME = 200


def decode_tokens(tokens: Sequence[int], start_idx: int = 0) -> np.ndarray:
    """music2midi/tokenizer.py:169-200 (_decode_tokens) + :242-267 (_tokens_to_note).

    Returns integer-valued float64 rows [onset_idx, offset_idx|-1, pitch, velocity].
    """
    notes: List[List[float]] = []
    cur_time, cur_on, cur_note = -1, -1, -1
    for tok in tokens:
        tok = int(tok)
        if tok == EOS:
            break
        if tok in (BOS, PAD):
            continue
        if tok == ONSET:
            cur_on = 1
        if tok == OFFSET:
            cur_on = 0
        if tok >= TIME_OFFSET:
            cur_time = start_idx + tok - TIME_OFFSET
            cur_on = -1
            cur_note = -1
        elif tok >= PITCH_OFFSET:
            cur_note = tok - PITCH_OFFSET
        if -1 in (cur_time, cur_on, cur_note):
            continue
        if cur_on:  # velocity != 0 -> onset row with dummy offset -1
            notes.append([cur_time, -1, cur_note, DEFAULT_VELOCITY])
        else:  # offset: closes ALL open rows of this pitch with onset strictly earlier
            for row in notes:
                if row[0] < cur_time and row[1] == -1 and row[2] == cur_note:
                    row[1] = cur_time
        cur_note = -1
    return np.asarray(notes, dtype=np.float64).reshape(-1, 4)


def decode_row(tokens, start_idx: int = 0, cutoff_time=None) -> np.ndarray:
    """music2midi/tokenizer.py:143-167 (_decode)."""
    notes = decode_tokens(np.asarray(tokens).tolist(), start_idx)
    notes = notes[notes[:, 1] != -1]
    notes[:, :2] = notes[:, :2] * TIME_STEP
    if cutoff_time is not None:
        notes = notes[notes[:, 0] < cutoff_time]
        notes[:, 1] = np.where(notes[:, 1] > cutoff_time, cutoff_time, notes[:, 1])
    return notes


def decode(tokens_batch, mode: str = "batched", duration_per_batch=None, cutoff_time=None):
    """music2midi/tokenizer.py:46-84."""
    if mode == "batched":
        return [decode_row(t, 0, cutoff_time) for t in tokens_batch]
    if mode == "sequential":
        assert duration_per_batch is not None
        n_steps = round(duration_per_batch / TIME_STEP)
        ret, start = [], 0
        for t in tokens_batch:
            ret.append(decode_row(t, start, cutoff_time))
            start += n_steps
        return np.concatenate(ret)
    raise ValueError(f"Invalid argument mode={mode}")


def tokenize_row(notes: np.ndarray, cutoff_time=None) -> List[int]:
    """music2midi/tokenizer.py:98-141,202-222 (_tokenize / _get_tokens)."""
    toks: List[int] = []
    if len(notes) > 0:
        notes = np.array(notes, dtype=np.float64, copy=True)
        if cutoff_time is not None:
            notes = notes[notes[:, 0] < cutoff_time]
        notes[:, 1] = np.maximum(notes[:, 1], notes[:, 0] + TIME_STEP)
        notes[:, :2] = notes[:, :2] / TIME_STEP
        notes[:, :2] = np.rint(np.nextafter(notes[:, :2], notes[:, :2] + 1))
        notes[:, :2] = np.minimum(notes[:, :2], N_TIME - 1)
        for idx in np.unique(notes[:, :2]):
            on = notes[notes[:, 0] == idx]
            off = notes[notes[:, 1] == idx]
            toks.append(int(idx + TIME_OFFSET))
            if len(on):
                toks += [ONSET] + [int(p + PITCH_OFFSET) for p in on[:, 2]]
            if len(off):
                toks += [OFFSET] + [int(p + PITCH_OFFSET) for p in off[:, 2]]
    toks.append(EOS)
    return toks


def tokenize(notes_batch, cutoff_time=None) -> torch.Tensor:
    """music2midi/tokenizer.py:86-96 (__call__): pad with PAD, int64."""
    rows = [tokenize_row(n, cutoff_time) for n in notes_batch]
    L = max(len(r) for r in rows)
    out = torch.zeros(len(rows), L, dtype=torch.long)
    for i, r in enumerate(rows):
        out[i, : len(r)] = torch.tensor(r, dtype=torch.long)
    return out


# ============================================================================ driver
def sample_tokens(waveform: torch.Tensor, W: Weights, split_size: int = 48000, split_duration: float = 3.0,
                  cond_index=None, batch_size: int = 128, max_length: int = 1024, generate_fn=None):
    """music2midi/model.py:101-140: split, chunk by batch_size, generate, sequential decode."""
    segs = torch.split(waveform, split_size)
    tokens_list = []
    for i in range(0, len(segs), batch_size):
        batch = torch.stack(segs[i : i + batch_size])
        ci = torch.zeros((len(batch), 2))
        if cond_index is not None:
            ci = ci + torch.tensor(cond_index, dtype=torch.float32)
        ci = ci.long()
        fn = generate_fn or (lambda w, c: generate(w, c, W, max_length))
        tokens_list += [*fn(batch, ci)]
    return decode(tokens_list, mode="sequential", duration_per_batch=split_duration), tokens_list


def forward_loss(wave, notes_batch, cond_index, W: Weights):
    """music2midi/transformer.py:28-39 + HF label shift (modeling_t5.py:595-614) + CE(ignore -100)."""
    labels = tokenize(notes_batch)
    labels = labels.masked_fill(labels == PAD, -100)
    x = conditioning(logmel(wave, W.window, W.fb), cond_index, W.cond_embeds)
    enc = encoder(x, W)
    dec_in = torch.cat([torch.full((labels.shape[0], 1), BOS, dtype=torch.long), labels[:, :-1]], dim=1)
    dec_in = dec_in.masked_fill(dec_in == -100, PAD)
    logits = decoder(dec_in, enc, W)
    loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1), ignore_index=-100)
    return loss, logits
