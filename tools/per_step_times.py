"""Per-step device times of every decode launch class at the benchmark batch (instrumented pass, CUDA-event pairs on
the launching stream; M2M_TIMING_DUMP makes the library write one "class,step,ms" line per timed launch group).

    python tools/per_step_times.py [--segments 2560] [--max-length 1024] [--out gpurun_out/per_step.csv]

Prints, per class, the mean launch duration in windows of steps, and a least-squares fit  t = a + b * bytes  for the two
KV-cache attention kernels (a = fixed cost per launch, 1/b = streaming bandwidth)."""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--segments", type=int, default=2560)
ap.add_argument("--max-length", type=int, default=1024)
ap.add_argument("--out", default="gpurun_out/per_step.csv")
args = ap.parse_args()
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
if os.path.exists(args.out):
    os.remove(args.out)

import numpy as np
import torch

from music2midi_b200 import synthetic as syn
from music2midi_b200._lib import KERNEL_CLASSES
from music2midi_b200.engine import Engine

dev = torch.device("cuda", 0)
eng = Engine(dev, "bf16")
eng.load_state_dict(syn.synthetic_state_dict(0))
n = args.segments
wave = torch.cat([syn.audio_noise(min(640, n - i), seed=i) for i in range(0, n, 640)]).to(dev)
cond = torch.zeros(n, 2, dtype=torch.long, device=dev)
emb = eng.condition(eng.logmel(wave), cond)
eng.generate_from_embeds(emb, args.max_length)  # warm-up (graph path)
os.environ["M2M_TIMING_DUMP"] = args.out
eng.set_flags(time_classes=True)
eng.generate_from_embeds(emb, args.max_length)
os.environ.pop("M2M_TIMING_DUMP")
eng.set_flags()
cfg = eng.config() if hasattr(eng, "config") else None

rows = np.loadtxt(args.out, delimiter=",")
by = collections.defaultdict(list)
for cls, step, ms in rows:
    by[int(cls)].append((int(step), ms))
inner, L = 512, 190
for cls, lst in sorted(by.items()):
    name = KERNEL_CLASSES[cls] if cls < len(KERNEL_CLASSES) else str(cls)
    a = np.array(lst)
    if a[:, 0].max() == 0:
        print(f"{name}: {len(a)} launches, {a[:, 1].sum():.3f} ms")
        continue
    line = [f"{name}: total {a[:, 1].sum():.1f} ms, per launch by step window:"]
    for lo in (0, 32, 64, 128, 256, 512, 768, 960):
        hi = {0: 32, 32: 64, 64: 128, 128: 256, 256: 512, 512: 768, 768: 960, 960: 1 << 30}[lo]
        m = (a[:, 0] >= lo) & (a[:, 0] < hi)
        if m.any():
            line.append(f"[{lo},{min(hi, args.max_length)}) {1e3 * a[m, 1].mean():.1f} us")
    print(" ".join(line))
    if name in ("dec_self_attn", "dec_cross_attn"):
        per_key = n * 2 * inner * 2
        x = (a[:, 0] + 1) * per_key if name == "dec_self_attn" else np.full(len(a), L * per_key, dtype=np.float64)
        if name == "dec_self_attn":
            A = np.stack([np.ones_like(x), x], 1)
            coef, *_ = np.linalg.lstsq(A, a[:, 1] * 1e-3, rcond=None)
            print(f"    fit: {coef[0] * 1e6:.1f} us fixed + bytes / {1e-9 / coef[1]:.0f} GB/s")
            for t in (63, 255, 511, 1022):
                m = a[:, 0] == t
                if m.any():
                    print(f"    step {t}: {1e3 * a[m, 1].mean():.1f} us, {x[m][0] / 1e9 / (a[m, 1].mean() * 1e-3):.0f} GB/s")
        else:
            print(f"    {x[0] / 1e9 / (a[:, 1].mean() * 1e-3):.0f} GB/s")
