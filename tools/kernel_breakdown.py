"""Per-kernel device-time breakdown of the hot path with torch.profiler (CUPTI), low overhead compared with
ncu's per-launch interception.  Usage: python tools/kernel_breakdown.py [--clips 256] [--max-length 96]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
from torch.profiler import ProfilerActivity, profile

from music2midi_b200 import synthetic as syn
from music2midi_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=256)
ap.add_argument("--max-length", type=int, default=96)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--out", default="gpurun_out/kernel_breakdown.json")
args = ap.parse_args()

dev = torch.device("cuda", 0)
eng = Engine(dev, args.precision)
eng.load_state_dict(syn.synthetic_state_dict(0))
n = args.clips * 10
wave = torch.cat([syn.audio_noise(min(640, n - i), seed=i) for i in range(0, n, 640)]).to(dev)
cond = torch.zeros(n, 2, dtype=torch.long, device=dev)
eng.generate(wave, cond, args.max_length)  # warm-up (allocations, graph capture)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.generate(wave, cond, args.max_length)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
    if t:
        rows.append({"name": e.key[:110], "count": e.count, "total_us": t, "avg_us": t / e.count})
rows.sort(key=lambda r: -r["total_us"])
tot = sum(r["total_us"] for r in rows)
for r in rows:
    print(f"{r['total_us']:12.1f} us {100 * r['total_us'] / tot:5.1f}%  n={r['count']:6d} avg={r['avg_us']:9.1f}  {r['name']}")
print("total device time us", tot, "last_generate_ms", eng.stats()["last_generate_ms"])
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump({"args": vars(args), "total_us": tot, "rows": rows}, open(args.out, "w"), indent=1)
