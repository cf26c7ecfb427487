"""Runs the log-mel frontend alone (BASELINE configs[1]: 640 segments) - used under ncu for the DFT-GEMM capture."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from music2midi_b200 import synthetic as syn
from music2midi_b200.engine import Engine

dev = torch.device("cuda", 0)
eng = Engine(dev, sys.argv[1] if len(sys.argv) > 1 else "bf16")
eng.load_state_dict({k: v for k, v in syn.synthetic_state_dict(0).items() if k.startswith("spectrogram.")})
wave = syn.audio_noise(640, 3).to(dev)
for _ in range(3):
    out = eng.logmel(wave)
torch.cuda.synchronize()
print(out.shape, float(out.mean()))
