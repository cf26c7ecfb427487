"""Runs and times the log-mel frontend alone (BASELINE configs[1]: 640 segments = 120,320 frames).  Also used under ncu
for the DFT-GEMM capture.  Usage: python tools/mel_only.py [bf16|fp32] [segments]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from music2midi_b200 import synthetic as syn
from music2midi_b200.engine import Engine

dev = torch.device("cuda", 0)
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 640
eng = Engine(dev, prec)
eng.load_state_dict({k: v for k, v in syn.synthetic_state_dict(0).items() if k.startswith("spectrogram.")})
wave = syn.audio_noise(n, 3).to(dev)
for _ in range(3):
    out = eng.logmel(wave)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = eng.logmel(wave)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
frames = n * 188
print(f"{prec}: {n} segments, {frames} frames: {ms:.3f} ms -> {frames / ms * 1e3 / 1e6:.2f} M frames/s, "
      f"{frames * 2.0 * 2048 * 2050 / ms / 1e9:.1f} TFLOP/s fp32-DFT-equivalent; mean {float(out.mean()):.4f}")
if "--profile" in sys.argv:
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.logmel(wave)
        torch.cuda.synchronize()
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
        if e.device_time_total:
            print(f"{e.device_time_total:10.1f} us  n={e.count:4d} avg={e.device_time_total / e.count:8.1f}  {e.key[:100]}")
