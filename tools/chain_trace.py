"""Timeline of the decode GEMM-chain kernels (csrc/chain_tc.cuh) from in-kernel clock64 stamps: per phase, when the
A producer passed the cluster handshake, when the first A tile landed, when the last MMA was committed, when the
epilogue saw the accumulator and finished.  Usage (GPU box):  M2M_CHAIN_TRACE=1 python tools/chain_trace.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["M2M_CHAIN_TRACE"] = "1"
import numpy as np
import torch

from music2midi_b200 import synthetic as syn
from music2midi_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2560
dev = torch.device("cuda", 0)
eng = Engine(dev, "bf16")
eng.load_state_dict(syn.synthetic_state_dict(0))
emb = eng.condition(eng.logmel(syn.audio_noise(64, 1).to(dev)), torch.zeros(64, 2, dtype=torch.long, device=dev))
emb = emb.repeat((n + 63) // 64, 1, 1)[:n].contiguous()
eng.generate_from_embeds(emb, 6)
grid, slots = C.c_int(0), C.c_int(0)
buf = np.zeros(4 * 6 * ((n + 127) // 128) * 256, dtype=np.int64)
eng.lib.m2m_debug_chain_trace(eng._ctx, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(grid), C.byref(slots))
G, S = grid.value, slots.value
t = buf[: 4 * G * S].reshape(4, G, S).astype(np.float64)
clk = 1.9  # GHz, approximate (clock64 counts SM cycles)
names = ["K0 (QKV l0)", "KB[0] (o-proj | cross-q)", "KA[0] (co | Wi | ffo | QKV)", "KA[last] (co | Wi | ffo | lm_head)"]
ev = ["A handshake passed", "A last TMA issued", "MMA first A tile", "MMA last commit", "epi saw acc", "epi stores done",
      "handshake sent", "B last TMA issued"]
for k in range(4):
    x = t[k]
    t0 = x[:, 0:1]
    rel = np.where(x > 0, (x - t0) / clk / 1e3, np.nan)  # us since this CTA's entry
    print(f"== {names[k]}: setup done {np.nanmean(rel[:, 1]):.2f} us, exit {np.nanmean(rel[:, 2]):.2f} us (max {np.nanmax(rel[:, 2]):.2f})")
    for p in range(4):
        seg = rel[:, 8 + 8 * p: 16 + 8 * p]
        if np.all(np.isnan(seg)):
            continue
        print(f"   phase {p}: " + "  ".join(f"{ev[i]} {np.nanmean(seg[:, i]):.2f}" for i in range(8) if not np.all(np.isnan(seg[:, i]))))

# per-k-block stamps of the MMA issuer in the traced phase (M2M_CHAIN_TRACE_PHASE, default 2) of KA[0], CTA 0
D = 40
x = t[2][0]
print("KA[0] CTA 0, MMA issuer per k-block (us since kernel entry): A ready, B ready, MMAs issued, commits issued")
for kb in range(32):
    r = x[D + 4 * kb: D + 4 * kb + 4]
    if r[0] == 0:
        break
    print(f"  kb {kb:2d}: " + "  ".join(f"{(v - x[0]) / clk / 1e3:7.3f}" for v in r))
