"""Per-kernel device-time breakdown of the teacher-forced forward (BASELINE configs[2]) with torch.profiler."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
from torch.profiler import ProfilerActivity, profile

from music2midi_b200 import synthetic as syn
from music2midi_b200.engine import Engine

dev = torch.device("cuda", 0)
eng = Engine(dev, sys.argv[1] if len(sys.argv) > 1 else "bf16")
eng.load_state_dict(syn.synthetic_state_dict(0))
nb, ld = 32, 1024
wave = syn.audio_noise(nb, 5).to(dev)
cz = torch.zeros(nb, 2, dtype=torch.long, device=dev)
dec_in = torch.randint(5, 333, (nb, ld), device=dev)
dec_in[:, 0] = 1


def fwd():
    enc = eng.encode(eng.condition(eng.logmel(wave), cz))
    return eng.decoder_forward(enc, dec_in)


for _ in range(2):
    fwd()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    fwd()
e1.record()
torch.cuda.synchronize()
print("forward ms", e0.elapsed_time(e1) / 5)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fwd()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
    if t:
        rows.append((t, e.count, e.key[:120]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
for t, n, k in rows:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:4d} avg={t / n:8.1f}  {k}")
print("total device us", tot)
