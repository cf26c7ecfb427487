"""Small driver for `ncu --set full` captures of every kernel class of the hot path: log-mel on a few
slabs, the encoder + cross-KV prefill and a handful of decode steps at the benchmark batch (2560 segments), with plain
launches (no CUDA graph) so that every kernel appears as its own launch.

    ncu --set full --clock-control none --import-source on -k regex:'<pattern>' -o gpurun_out/prof_x \
        python tools/profile_classes.py [--clips 256] [--max-length 4] [--precision bf16] [--mel-segments 64]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from music2midi_b200 import synthetic as syn
from music2midi_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=256)
ap.add_argument("--max-length", type=int, default=4)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--mel-segments", type=int, default=64)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--teacher-forced", type=int, default=0, help="also run a teacher-forced forward with this many rows")
args = ap.parse_args()

dev = torch.device("cuda", 0)
eng = Engine(dev, args.precision)
eng.load_state_dict(syn.synthetic_state_dict(0))
eng.set_flags(graph=args.graph)
n = args.clips * 10
wave = torch.cat([syn.audio_noise(min(640, n - i), seed=i) for i in range(0, n, 640)]).to(dev)
cond = torch.zeros(n, 2, dtype=torch.long, device=dev)
mel = eng.logmel(wave[: args.mel_segments])
emb = eng.condition(mel, cond[: args.mel_segments])
# prefill + decode steps at the benchmark batch: embeddings are reused (tiled) so that the mel kernels appear only once
reps = (n + args.mel_segments - 1) // args.mel_segments
emb_all = emb.repeat(reps, 1, 1)[:n].contiguous()
tok = eng.generate_from_embeds(emb_all, args.max_length)
if args.teacher_forced:
    nb = args.teacher_forced
    dec_in = torch.randint(5, 333, (nb, 1024), device=dev)
    dec_in[:, 0] = 1
    enc = eng.encode(emb_all[:nb])
    eng.decoder_forward(enc, dec_in)
torch.cuda.synchronize()
print("ok", tuple(mel.shape), tuple(tok.shape), int(eng.stats()["kernel_launches"]))
