#!/usr/bin/env python
"""Headline benchmark: audio-seconds transcribed per second (log-mel + encoder + KV-cached greedy
decode to 1024 tokens) on synthetic 30 s clips, BASELINE.json configs[3] per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the whole hot path over one batch of `--clips` (default 256) synthetic 30 s
clips = 10 x clips independent 3 s segments per GPU (weak scaling: per-GPU work fixed), followed for
N > 1 by the NCCL all-gather of the int16 token streams.  One JSON line is printed by rank 0.

  value     inputs already resident in HBM, device-timed with CUDA events, max over ranks
  e2e       the same work through the C-ABI host-buffer entry point (m2m_transcribe_host): pinned host
            waveforms -> H2D -> hot path -> D2H tokens, copies inside the timed region
  roofline  the dominant kernel (KV-cached decode self-attention, HBM-bound): algorithmic KV bytes of
            every launch / CUDA-event time of every launch, in a separate instrumented pass
  cpu_baseline  the reference's own library code path (torchaudio MelSpectrogram + HF T5 generate,
            oracle/hf_path.py) on this box's host cores, on a bounded sample of the same workload

`--impl reference` times only that CPU path (rank 0; other ranks exit).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEG_SAMPLES = 48000
SEG_SECONDS = 3.0
SEGS_PER_CLIP = 10
MAX_LENGTH = 1024
METRIC = "audio_seconds_transcribed_per_second"
UNIT = "audio-s/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("M2M_BENCH_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--clips", type=int, default=int(os.environ.get("M2M_BENCH_CLIPS", 256)), help="30 s clips per GPU")
    ap.add_argument("--ref-clips", type=int, default=1,
                    help="clips in the bounded CPU-reference sample (1 clip = 10 segments is the reference's own single-"
                         "recording case and its best CPU throughput: 40 segments per batch measured 0.60 vs 0.88 "
                         "audio-s/s on 8 cores); 0 = largest batch that fits the time budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--max-length", type=int, default=MAX_LENGTH,
                    help="decode length cap; anything but 1024 is a profiling aid, not the benchmark workload")
    return ap.parse_args()


def workload_name(clips, max_length=MAX_LENGTH):
    return (f"full inference: {clips} synthetic 30 s clips ({clips * SEGS_PER_CLIP} x 3 s segments) per GPU, "
            f"log-mel + encoder + KV-cached greedy decode to {max_length} tokens, random-init weights at config.yaml dims")


# ------------------------------------------------------------------------------------ CPU reference
def cpu_reference_run(n_clips: int, steps: int, warmup: int, budget_s: float = 150.0):
    """The reference's CPU path (same torchaudio / HF calls as music2midi/transformer.py:41-45),
    all host threads, bounded sample.  n_clips <= 0: pick the largest batch (up to 12 clips = 120 segments, the
    reference chunks by inference.batch_size = 128 segments) whose (steps + warmup) full-length runs fit the time
    budget, from a short calibration.  Returns (audio_s_per_s, ms_per_step, cores, sample string)."""
    import torch

    from music2midi_b200 import synthetic as syn
    from oracle import hf_path

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = hf_path.build(syn.synthetic_state_dict(0))
    with torch.no_grad():
        model.generate(syn.audio_noise(2, 1), torch.zeros(2, 2, dtype=torch.long), max_length=8)  # library warm-up
        if n_clips <= 0:
            n_clips = 1
            for cand in (12, 8, 4, 2):
                w = syn.audio_noise(cand * SEGS_PER_CLIP, seed=99)
                c = torch.zeros(cand * SEGS_PER_CLIP, 2, dtype=torch.long)
                t0 = time.perf_counter()
                model.generate(w, c, max_length=25)
                per_token = (time.perf_counter() - t0) / 24
                # late steps attend over a longer cache: measured ~1.7x the early-step cost at 1024 tokens
                if per_token * 1.4 * (MAX_LENGTH - 1) * (steps + warmup) <= budget_s:
                    n_clips = cand
                    break
    n_seg = n_clips * SEGS_PER_CLIP
    wave = syn.audio_noise(n_seg, seed=100)
    cond = torch.zeros(n_seg, 2, dtype=torch.long)
    with torch.no_grad():
        for _ in range(warmup):
            model.generate(wave, cond, max_length=MAX_LENGTH)
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            out = model.generate(wave, cond, max_length=MAX_LENGTH)
            times.append(time.perf_counter() - t0)
    assert out.shape[0] == n_seg
    total = sum(times)
    sample = (f"{n_clips} clip(s) = {n_seg} segments as one batch, mel + HF T5 greedy generate to "
              f"{out.shape[1]} tokens, fp32, {steps} timed run(s)")
    return n_seg * SEG_SECONDS * steps / total, 1e3 * total / steps, torch.get_num_threads(), sample


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    val, ms, cores, sample = cpu_reference_run(args.ref_clips, max(1, args.steps), args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.clips), "reference_sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.path = gpu_index, None, f"/tmp/m2m_clocks_{os.getpid()}.csv"

    def start(self):
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        self.f = open(self.path, "w")
        self.proc = subprocess.Popen([exe, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                      "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for ln in open(self.path):
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return None
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ ours
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from music2midi_b200 import synthetic as syn
    from music2midi_b200.distributed import gather_tokens
    from music2midi_b200.engine import Engine

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there when NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_seg = args.clips * SEGS_PER_CLIP
    MAXLEN = args.max_length
    eng = Engine(dev, args.precision)
    eng.load_state_dict(syn.synthetic_state_dict(0))

    # synthetic audio, generated in slabs (seed = global clip index -> every rank has different clips)
    host_wave = torch.empty(n_seg, SEG_SAMPLES, dtype=torch.float32, pin_memory=True)
    slab = 64
    for c0 in range(0, args.clips, slab):
        c1 = min(args.clips, c0 + slab)
        host_wave[c0 * SEGS_PER_CLIP: c1 * SEGS_PER_CLIP] = syn.audio_noise(
            (c1 - c0) * SEGS_PER_CLIP, seed=1000 + rank * 100000 + c0)
    wave = host_wave.to(dev)
    cond = torch.zeros(n_seg, 2, dtype=torch.int64, device=dev)
    host_cond = np.zeros((n_seg, 2), dtype=np.int64)

    def step_device():
        tok = eng.generate(wave, cond, MAXLEN)
        if world > 1:
            full = torch.zeros(n_seg, MAXLEN, dtype=torch.int16, device=dev)
            full[:, : tok.shape[1]] = tok.to(torch.int16)
            tok = gather_tokens(full, n_seg * world)
        return tok

    def step_host():
        toks, lens = eng.transcribe_host(host_wave.numpy(), host_cond, MAXLEN, device_batch=n_seg)
        if world > 1:
            full = torch.from_numpy(toks).to(dev).to(torch.int16)
            gather_tokens(full, n_seg * world)
        return toks

    for _ in range(args.warmup):
        tok = step_device()
    barrier()
    eng.stats(reset=True)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        tok = step_device()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = int(eng.stats()["kernel_launches"])
    out_len = int(tok.shape[1])

    # end-to-end through the host-buffer C-ABI entry point
    step_host()  # warm the staging buffers
    barrier()
    e0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])

    roofline = None
    if rank == 0 and not args.no_roofline:
        eng.set_flags(graph=False, time_attention=True)
        eng.generate(wave, cond, MAXLEN)
        st = eng.stats()
        eng.set_flags(graph=True, time_attention=False)
        peak, how = measured_peak()
        n_launch = max(int(st["last_attn_launches"]), 1)
        achieved = st["attn_bytes"] / 1e9 / (st["last_attn_ms"] / 1e3) if st["last_attn_ms"] > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:  # one ncu --set full capture of this kernel; DRAM bytes scale with the algorithmic bytes
                traffic = json.load(open(tp))["traffic_over_algorithmic"] * st["attn_bytes"] / n_launch
            except Exception:
                traffic = None
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": "decode_attn_kernel<SELF> (KV-cached decode self-attention)",
            "peak_source": how, "launches": n_launch,
            "algorithmic_bytes_per_launch_avg": st["attn_bytes"] / n_launch,
            "avg_launch_us": 1e3 * st["last_attn_ms"] / n_launch,
            "share_of_step": st["last_attn_ms"] / st["last_generate_ms"] if st["last_generate_ms"] else None,
            "how": "separate instrumented pass: CUDA events around every launch on the launching stream",
        }

    # BASELINE.json configs[1] (mel frontend only, 64 clips = 640 segments = 120,320 frames), reported as extra
    extra = None
    if rank == 0:
        n_mel = min(640, n_seg)
        for _ in range(3):
            eng.logmel(wave[:n_mel])
        torch.cuda.synchronize()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        m0.record()
        for _ in range(reps):
            eng.logmel(wave[:n_mel])
        m1.record()
        torch.cuda.synchronize()
        mel_ms = m0.elapsed_time(m1) / reps
        frames = n_mel * (1 + SEG_SAMPLES // 256)
        dft_flop = frames * 2.0 * 2048 * 2050
        extra = {"mel_only": {"segments": n_mel, "frames": frames, "ms": mel_ms, "frames_per_s": frames / (mel_ms / 1e3),
                              "dft_gemm_tflops_algorithmic": dft_flop / (mel_ms / 1e3) / 1e12,
                              "path": "tcgen05 split-bf16 (6 bf16 MMA products per fp32 product)" if args.precision == "bf16"
                              else "fp32 CUDA-core DFT"}}

        # BASELINE.json configs[2]: encoder forward + teacher-forced decoder forward, bf16, batch 32 (inference-shape
        # arithmetic of T5Transformer.forward, labels of length 1024, L_enc = 190)
        try:
            nb, ld = 32, 1024
            dec_in = torch.randint(5, 333, (nb, ld), device=dev)
            dec_in[:, 0] = 1
            cz = torch.zeros(nb, 2, dtype=torch.int64, device=dev)

            def fwd():
                enc = eng.encode(eng.condition(eng.logmel(wave[:nb]), cz))
                return eng.decoder_forward(enc, dec_in)

            for _ in range(2):
                fwd()
            torch.cuda.synchronize()
            m0.record()
            for _ in range(5):
                fwd()
            m1.record()
            torch.cuda.synchronize()
            f_ms = m0.elapsed_time(m1) / 5
            gflop = nb * (1.579 + 5.262 + 47.29)  # mel DFT + encoder + teacher-forced decoder, per segment
            extra["teacher_forced_forward"] = {"batch": nb, "label_len": ld, "enc_len": 190, "ms": f_ms,
                                               "tflops_algorithmic": gflop / f_ms, "dtype": args.precision}
        except Exception as e:  # reported, never fatal for the headline
            extra["teacher_forced_forward"] = {"error": str(e)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms, cores, sample = cpu_reference_run(args.ref_clips, 1, 0, budget_s=30.0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "ms_per_sample": ms}
        try:  # BASELINE.json configs[1]: the reference's CPU feature extraction (torchaudio MelSpectrogram + log)
            from oracle import hf_path

            spec = hf_path.build().spectrogram
            w = syn.audio_noise(640, seed=7)
            spec(w[:64])
            t0 = time.perf_counter()
            spec(w)
            dt = time.perf_counter() - t0
            cpu["mel_only"] = {"segments": 640, "frames": 640 * 188, "ms": 1e3 * dt, "frames_per_s": 640 * 188 / dt}
        except Exception as e:
            cpu["mel_only"] = {"error": str(e)[:200]}

    if world > 1:
        dist.barrier()
    if rank == 0:
        audio_s = n_seg * SEG_SECONDS * world * args.steps
        line = {
            "metric": METRIC, "value": audio_s / (ms_total / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision if args.precision != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.clips, MAXLEN), "segments_per_gpu": n_seg, "device_batch": n_seg,
                       "max_length": MAXLEN, "generated_length": out_len, "parallelism": f"clip-sharded x{world}",
                       "l2": "inputs larger than L2 (KV cache per GPU >> 126 MB)"},
            "e2e": {"value": audio_s / (ms_e2e / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": n_seg * SEG_SAMPLES * 4 + n_seg * 16,
                    "d2h_bytes_per_step": n_seg * MAXLEN * 8,
                    "api": "m2m_transcribe_host (C ABI, pinned host buffers)"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
