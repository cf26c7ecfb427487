#!/usr/bin/env python
"""Headline benchmark: audio-seconds transcribed per second (log-mel + encoder + KV-cached greedy
decode to 1024 tokens) on synthetic 30 s clips, BASELINE.json configs[3] per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the whole hot path over one batch of `--clips` (default 256) synthetic 30 s
clips = 10 x clips independent 3 s segments per GPU (weak scaling: per-GPU work fixed), followed for
N > 1 by the NCCL all-gather of the int16 token streams.  `--clips-total C` instead fixes the TOTAL number
of clips (strong scaling, BASELINE.json configs[4] as written: 2048 clips split over the ranks; a rank
whose share exceeds `--device-batch` segments runs several device batches).  One JSON line is printed by rank 0.

  value     inputs already resident in HBM, device-timed with CUDA events, max over ranks
  e2e       the same work through the C-ABI host-buffer entry point (m2m_transcribe_host): pinned host
            waveforms -> H2D -> hot path -> D2H tokens, copies inside the timed region
  golden    16 rows of every batch are the golden inputs of tests/golden/generate.npz (recorded from the live
            reference): `tokens_match_golden` compares the tokens the TIMED run produced for them
  roofline  the dominant kernel (KV-cached decode self-attention, HBM-bound): algorithmic KV bytes of
            every launch / CUDA-event time of every launch, in a separate instrumented pass;
            roofline_by_class = the same for every kernel class of the path
  extra.fp32_parity_mode   the same workload in the fp32 parity mode (token output identical to the reference)
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
GOLDEN_STRIDE = 160  # golden input i sits at row GOLDEN_OFFSET + i * GOLDEN_STRIDE of every rank's batch
GOLDEN_OFFSET = 7


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("M2M_BENCH_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--clips", type=int, default=int(os.environ.get("M2M_BENCH_CLIPS", 256)), help="30 s clips per GPU")
    ap.add_argument("--clips-total", type=int, default=0,
                    help="total clips, split clip-wise over the ranks (strong scaling; 0 = use --clips per GPU)")
    ap.add_argument("--device-batch", type=int, default=2560, help="segments per device batch")
    ap.add_argument("--ref-clips", type=int, default=1,
                    help="clips in the bounded CPU-reference sample (1 clip = 10 segments is the reference's own single-"
                         "recording case and its best CPU throughput: 40 segments per batch measured 0.60 vs 0.88 "
                         "audio-s/s on 8 cores); 0 = largest batch that fits the time budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32 parity-mode line")
    ap.add_argument("--no-extra", action="store_true", help="skip configs 2/3 and the API-level e2e figure")
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


# ------------------------------------------------------------------------------------ roofline arithmetic
def measured_peaks():
    """(hbm GB/s, bf16 TFLOP/s burst, bf16 TFLOP/s sustained, source string)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return (float(d["hbm_gbs"]), float(d["bf16_tflops"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    "measured (MEASURED_PEAKS.json)")
        except Exception:
            pass
    return 6650.0, 1590.0, 1400.0, "fallback (B200_PROFILING.md: 6.65 TB/s, 1.59 PFLOP/s burst / 1.4 sustained)"


# algorithmic work per 3 s segment (SURVEY.md 8d conventions), as functions of the shapes
D, I, F, V, NL, NFFT = 384, 512, 1152, 400, 6, 2048


def flops_mel(t_frames):
    return 2.0 * t_frames * NFFT * (NFFT + 2)


def flops_enc_gemm(l_enc):
    return NL * l_enc * (2.0 * D * I * 4 + 2.0 * D * F * 3)


def flops_enc_attn(l_enc):
    return NL * 4.0 * l_enc * l_enc * I


def flops_cross_kv(l_enc):
    return NL * 2.0 * D * 2 * I * l_enc


def flops_dec_step_weights():
    return NL * (2.0 * D * I * 4 + 2.0 * D * I * 2 + 2.0 * D * F * 3) + 2.0 * D * V


def flops_teacher_forced(ld, l_enc):
    """Decoder with full (un-halved) causal self-attention, the convention of SURVEY.md 8d (47.29 GFLOP at 1024/190)."""
    per_layer = (2.0 * D * I * 4 * ld + 4.0 * ld * ld * I + 2.0 * D * I * 2 * ld + 2.0 * D * 2 * I * l_enc +
                 4.0 * ld * l_enc * I + 2.0 * D * F * 3 * ld)
    return NL * per_layer + 2.0 * D * V * ld


def class_rooflines(st, n_seg, l_enc, t_frames, elt, peaks):
    """One entry per kernel class of the instrumented pass: CUDA-event time, launches, share of the pass and, where the
    class has a meaningful bound, algorithmic work / time against the measured peak."""
    hbm, _, tf_sus, _ = peaks
    ms, ln = st["class_ms"], st["class_launches"]
    total = sum(ms.values()) or 1.0
    steps = ln["dec_select"] or 1
    work = {
        "mel_dft": ("tensor", n_seg * flops_mel(t_frames)),
        "enc_gemm": ("tensor", n_seg * flops_enc_gemm(l_enc)),
        "enc_attn": ("tensor", n_seg * flops_enc_attn(l_enc)),
        "cross_kv": ("tensor", n_seg * flops_cross_kv(l_enc)),
        "dec_chain": ("tensor", n_seg * flops_dec_step_weights() * steps),
        "dec_self_attn": ("hbm", float(st["attn_bytes"])),
        "dec_cross_attn": ("hbm", float(st["cross_attn_bytes"])),
        "enc_norm": ("hbm", ln["enc_norm"] * n_seg * l_enc * D * (4.0 + elt)),
        "mel_frame": ("hbm", n_seg * (SEG_SAMPLES * 4.0 + t_frames * NFFT * 2.0 * 3)),
        "mel_band": ("hbm", n_seg * t_frames * ((NFFT // 2 + 4) * 4.0 + D * 4.0)),
    }
    notes = {
        "dec_chain": "latency-bound by construction: 30.4 MFLOP per segment-step between two HBM-bound attention kernels",
        "mel_dft": "algorithmic = fp32 DFT-as-GEMM flops; the tcgen05 path executes 3x (split-bf16 products)",
        "mel_frame": "bytes = waveform in + three bf16 frame terms out (L2-resident slab)",
        "mel_band": "bytes = power spectrum in (L2-resident slab) + log-mel out",
    }
    out = {}
    for k in ms:
        if ln[k] == 0:
            continue
        e = {"ms": ms[k], "launches": ln[k], "share": ms[k] / total}
        if k in work and ms[k] > 0:
            bound, w = work[k]
            if bound == "hbm":
                e.update(bound="hbm", achieved=w / 1e9 / (ms[k] / 1e3), peak=hbm, unit="GB/s")
            else:
                e.update(bound="tensor", achieved=w / 1e12 / (ms[k] / 1e3), peak=tf_sus, unit="TFLOP/s")
            e["frac"] = e["achieved"] / e["peak"]
        if k in notes:
            e["note"] = notes[k]
        out[k] = e
    return out


# ------------------------------------------------------------------------------------ ours
def golden_inputs():
    """The 16 golden inputs (tests/golden/make_golden.py recorded the live reference's tokens for them)."""
    import numpy as np
    import torch

    from music2midi_b200 import synthetic as syn

    wave = torch.cat([syn.audio_noise(8, 0), syn.audio_tones(8, 0)])
    cond = torch.stack([torch.arange(16) % 6, torch.arange(16) % 3], 1)
    g = np.load(os.path.join(ROOT, "tests", "golden", "generate.npz"))
    return wave, cond, torch.from_numpy(g["tokens"].astype(np.int64)), torch.from_numpy(g["gap"])


def golden_report(tok_rows, gold_tokens, gap, precision):
    """tok_rows: [16, L] tokens the run produced for the golden rows."""
    import torch

    L = tok_rows.shape[1]
    ref = gold_tokens[:, :L]
    res = {"rows": 16, "tokens_per_row": L, "exact_rows": 0, "first_divergence": [], "golden_gap_at_divergence": []}
    for r in range(16):
        if torch.equal(tok_rows[r], ref[r]):
            res["exact_rows"] += 1
            res["first_divergence"].append(None)
            res["golden_gap_at_divergence"].append(None)
        else:
            k = int((tok_rows[r] != ref[r]).nonzero()[0])
            res["first_divergence"].append(k)
            res["golden_gap_at_divergence"].append(float(gap[r, k - 1]))
    res["match"] = res["exact_rows"] == 16
    res["expectation"] = ("fp32 parity mode: 16/16 rows identical to the reference's tokens" if precision == "fp32" else
                          "bf16 throughput mode: logits within the stated bf16 tolerance; tokens may leave the reference "
                          "at near-ties (golden top-2 gap below 0.2)")
    return res


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from music2midi_b200 import synthetic as syn
    from music2midi_b200.distributed import gather_tokens, shard_range
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

    strong = args.clips_total > 0
    if strong:
        lo, hi = shard_range(args.clips_total, rank, world)
        my_clips, clip0, total_clips = hi - lo, lo, args.clips_total
    else:
        my_clips, clip0, total_clips = args.clips, rank * args.clips, args.clips * world
    n_seg = my_clips * SEGS_PER_CLIP
    n_seg_total = total_clips * SEGS_PER_CLIP
    if strong:
        rank_rows = [(shard_range(args.clips_total, r, world)[1] - shard_range(args.clips_total, r, world)[0]) * SEGS_PER_CLIP
                     for r in range(world)]
    else:
        rank_rows = [n_seg] * world
    dev_batch = min(args.device_batch, max(n_seg, 1))
    MAXLEN = args.max_length
    peaks = measured_peaks()

    # synthetic audio (seed = global clip index block -> every rank has different clips); 16 golden rows per rank
    host_wave = torch.empty(n_seg, SEG_SAMPLES, dtype=torch.float32, pin_memory=True)
    slab = 64
    for c0 in range(0, my_clips, slab):
        c1 = min(my_clips, c0 + slab)
        host_wave[c0 * SEGS_PER_CLIP: c1 * SEGS_PER_CLIP] = syn.audio_noise(
            (c1 - c0) * SEGS_PER_CLIP, seed=1000 + (clip0 + c0))
    host_cond_t = torch.zeros(n_seg, 2, dtype=torch.int64)
    gw, gc, gold_tokens, gold_gap = golden_inputs()
    grows = [GOLDEN_OFFSET + i * GOLDEN_STRIDE for i in range(16)]
    have_golden = n_seg > grows[-1]
    if have_golden:
        host_wave[grows] = gw
        host_cond_t[grows] = gc
    host_cond = host_cond_t.numpy()
    wave = host_wave.to(dev)
    cond = host_cond_t.to(dev)

    def make_steps(eng):
        def step_device():
            outs = []
            for i0 in range(0, n_seg, dev_batch):
                t = eng.generate(wave[i0:i0 + dev_batch], cond[i0:i0 + dev_batch], MAXLEN)
                if t.shape[1] < MAXLEN:  # dynamic HF length: pad to the cap like the final token matrix
                    t = torch.nn.functional.pad(t, (0, MAXLEN - t.shape[1]))
                outs.append(t)
            tok = outs[0] if len(outs) == 1 else torch.cat(outs)
            if world > 1:
                tok = gather_tokens(tok.to(torch.int16), n_seg_total, counts=rank_rows, out_dtype=torch.int16)
            return tok

        def step_host():
            toks, lens = eng.transcribe_host(host_wave.numpy(), host_cond, MAXLEN, device_batch=dev_batch)
            if world > 1:
                gather_tokens(torch.from_numpy(toks.astype(np.int16)).to(dev), n_seg_total, counts=rank_rows,
                              out_dtype=torch.int16)
            return toks

        return step_device, step_host

    def my_rows(tok):
        """Rows of this rank inside a (possibly gathered) token matrix."""
        if world == 1:
            return tok
        off = clip0 * SEGS_PER_CLIP
        return tok[off: off + n_seg]

    def timed(fn, steps, warm):
        for _ in range(warm):
            out = fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        wall = 1e3 * (time.perf_counter() - t0)
        ms = max(e0.elapsed_time(e1), 0.0)
        t = torch.tensor([ms, wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t[0]), float(t[1])

    audio_s_step = n_seg_total * SEG_SECONDS

    # ------------------------------------------------------------------ headline (args.precision)
    eng = Engine(dev, args.precision)
    eng.load_state_dict(syn.synthetic_state_dict(0))
    step_device, step_host = make_steps(eng)
    for _ in range(args.warmup):
        step_device()
    barrier()
    eng.stats(reset=True)
    sampler = ClockSampler(local)
    sampler.start()
    tok, ms_total, _ = timed(step_device, args.steps, 0)
    clocks = sampler.stop()
    launches = int(eng.stats()["kernel_launches"])
    golden = None
    if have_golden and rank == 0:
        rows = my_rows(tok)[torch.tensor(grows, device=tok.device)].to(torch.int64).cpu()
        golden = golden_report(rows, gold_tokens, gold_gap, args.precision)
    out_len = int(tok.shape[1])

    # end-to-end through the host-buffer C-ABI entry point (H2D + D2H inside the timed region)
    toks_h, ms_e2e_dev, ms_e2e_wall = timed(step_host, args.steps, 1)
    ms_e2e = max(ms_e2e_dev, ms_e2e_wall)
    e2e_golden = None
    if have_golden and rank == 0:
        e2e_golden = golden_report(torch.from_numpy(np.ascontiguousarray(toks_h[grows])).to(torch.int64), gold_tokens,
                                   gold_gap, args.precision)["exact_rows"]

    # ------------------------------------------------------------------ instrumented pass: roofline(s)
    roofline, by_class = None, None
    if rank == 0 and not args.no_roofline:
        nb = min(dev_batch, n_seg)
        eng.set_flags(graph=False, time_classes=True)
        eng.generate(wave[:nb], cond[:nb], MAXLEN)
        st = eng.stats()
        eng.set_flags()
        elt = 2 if args.precision == "bf16" else 4
        n_launch = max(int(st["last_attn_launches"]), 1)
        achieved = st["attn_bytes"] / 1e9 / (st["last_attn_ms"] / 1e3) if st["last_attn_ms"] > 0 else 0.0
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:  # NOT measured in this run: DRAM/algorithmic ratio of one ncu --set full capture of this kernel
                tj = json.load(open(tp))
                traffic = tj["traffic_over_algorithmic"] * st["attn_bytes"] / n_launch
                traffic_src = f"profiles/roofline_traffic.json ({tj.get('source', 'ncu --set full capture')}), scaled by the algorithmic bytes"
            except Exception:
                traffic = None
        pass_ms = sum(st["class_ms"].values())
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": peaks[0], "unit": "GB/s", "frac": achieved / peaks[0],
            "traffic": traffic, "traffic_source": traffic_src,
            "kernel": "decode_attn_kernel<SELF> (KV-cached decode self-attention)",
            "peak_source": peaks[3], "launches": n_launch,
            "algorithmic_bytes_per_launch_avg": st["attn_bytes"] / n_launch,
            "avg_launch_us": 1e3 * st["last_attn_ms"] / n_launch,
            "share_of_step": st["last_attn_ms"] / pass_ms if pass_ms else None,
            "how": "separate instrumented pass: one CUDA-event pair around every launch on the launching stream; "
                   "finished rows are not skipped in this pass",
        }
        by_class = class_rooflines(st, nb, 190, 188, elt, peaks)

    # ------------------------------------------------------------------ extras: configs 2 and 3, API-level e2e
    extra = {}
    if rank == 0 and not args.no_extra:
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def time_ms(fn, reps, warm=2):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            m0.record()
            for _ in range(reps):
                fn()
            m1.record()
            torch.cuda.synchronize()
            return m0.elapsed_time(m1) / reps

        # BASELINE.json configs[1]: mel frontend only, 64 clips = 640 segments = 120,320 frames
        n_mel = min(640, n_seg)
        mel_ms = time_ms(lambda: eng.logmel(wave[:n_mel]), 10, 3)
        frames = n_mel * (1 + SEG_SAMPLES // 256)
        mel_tf = n_mel * flops_mel(188) / (mel_ms / 1e3) / 1e12
        extra["mel_only"] = {
            "segments": n_mel, "frames": frames, "ms": mel_ms, "frames_per_s": frames / (mel_ms / 1e3),
            "roofline": {"bound": "tensor", "achieved": mel_tf, "peak": peaks[1], "unit": "TFLOP/s", "frac": mel_tf / peaks[1],
                         "algorithmic": "fp32 DFT-as-GEMM flops 2*T*2048*2050 per segment", "peak_kind": "burst"},
            "path": "tcgen05 split-bf16 DFT"}

        # BASELINE.json configs[2]: mel + encoder + teacher-forced decoder forward, batch 32, at the inference shape
        # (S = 48000, L_enc = 190) and the training shape (3 s @ 22050 Hz: S = 66150, L_enc = 261), labels 256 / 1024
        tf = []
        for S, ld in ((48000, 1024), (48000, 256), (66150, 1024), (66150, 256)):
            try:
                nb = 32
                w = syn.audio_noise(nb, seed=5, samples=S).to(dev)
                dec_in = torch.randint(5, 333, (nb, ld), device=dev)
                dec_in[:, 0] = 1
                cz = torch.zeros(nb, 2, dtype=torch.int64, device=dev)

                def fwd():
                    return eng.decoder_forward(eng.encode(eng.condition(eng.logmel(w), cz)), dec_in)

                f_ms = time_ms(fwd, 5, 2)
                t_frames = 1 + S // 256
                l_enc = t_frames + 2
                gflop = nb * (flops_mel(t_frames) + flops_enc_gemm(l_enc) + flops_enc_attn(l_enc) +
                              flops_teacher_forced(ld, l_enc)) / 1e9
                tf.append({"batch": nb, "samples": S, "enc_len": l_enc, "label_len": ld, "ms": f_ms,
                           "tflops_algorithmic": gflop / f_ms, "frac_of_burst_peak": gflop / f_ms / peaks[1],
                           "dtype": args.precision})
            except Exception as e:  # reported, never fatal for the headline
                tf.append({"samples": S, "label_len": ld, "error": str(e)[:200]})
        extra["teacher_forced_forward"] = tf

        # the reference's public API end to end: list of recordings -> MIDI objects (Music2MIDI.generate_many:
        # pinned upload, hot path, D2H tokens, token -> notes state machine, notes -> PrettyMIDI-compatible objects)
        try:
            from music2midi_b200.model import Music2MIDI

            m = Music2MIDI(os.path.join(ROOT, "music2midi_b200", "config.yaml"), precision=args.precision)
            m.model.load_state_dict(syn.synthetic_state_dict(0))
            m = m.to(dev)
            n_api = min(my_clips, dev_batch // SEGS_PER_CLIP)
            clips_np = [host_wave[i * SEGS_PER_CLIP:(i + 1) * SEGS_PER_CLIP].reshape(-1).numpy() for i in range(n_api)]
            eng.close()  # one 38 GB KV cache at a time
            eng = None
            m.generate_many(clips_np[:2])
            m.generate_many(clips_np)  # warm-up at full size (allocations, graph capture)
            t0 = time.perf_counter()
            midis = m.generate_many(clips_np)
            dt = time.perf_counter() - t0
            extra["e2e_api"] = {"api": "Music2MIDI.generate_many (host arrays -> MIDI objects)", "clips": n_api,
                                "value": n_api * SEGS_PER_CLIP * SEG_SECONDS / dt, "unit": UNIT, "ms": 1e3 * dt,
                                "notes": int(sum(len(x.instruments[0].notes) for x in midis))}
            del m
        except Exception as e:
            extra["e2e_api"] = {"error": str(e)[:300]}

    # ------------------------------------------------------------------ fp32 parity mode on the same workload
    if eng is not None:
        eng.close()
    del eng
    torch.cuda.empty_cache()
    if not args.no_fp32 and args.precision != "fp32":
        barrier()
        eng32 = Engine(dev, "fp32")
        eng32.load_state_dict(syn.synthetic_state_dict(0))
        sd32, sh32 = make_steps(eng32)
        k32 = max(1, min(args.steps, 2))
        tok32, ms32, _ = timed(sd32, k32, 1)
        _, ms32_e2e_dev, ms32_e2e_wall = timed(sh32, 1, 0)
        fp32 = {"value": audio_s_step * k32 / (ms32 / 1e3), "unit": UNIT, "steps": k32, "warmup": 1,
                "ms_per_step": ms32 / k32, "dtype": "f32",
                "e2e": {"value": audio_s_step / (max(ms32_e2e_dev, ms32_e2e_wall) / 1e3), "unit": UNIT, "steps": 1}}
        if have_golden and rank == 0:
            rows = my_rows(tok32)[torch.tensor(grows, device=tok32.device)].to(torch.int64).cpu()
            fp32["tokens_match_golden"] = golden_report(rows, gold_tokens, gold_gap, "fp32")
        if rank == 0 and not args.no_roofline:
            nb = min(dev_batch, n_seg)
            eng32.set_flags(graph=False, time_classes=True)
            eng32.generate(wave[:nb], cond[:nb], MAXLEN)
            st = eng32.stats()
            eng32.set_flags()
            if st["last_attn_ms"] > 0:
                a = st["attn_bytes"] / 1e9 / (st["last_attn_ms"] / 1e3)
                fp32["roofline"] = {"bound": "hbm", "achieved": a, "peak": peaks[0], "unit": "GB/s", "frac": a / peaks[0],
                                    "kernel": "decode_attn_kernel<float, SELF>", "launches": int(st["last_attn_launches"]),
                                    "share_of_step": st["last_attn_ms"] / (sum(st["class_ms"].values()) or 1.0),
                                    "class_ms": {k: v for k, v in st["class_ms"].items() if v > 0}}
        extra["fp32_parity_mode"] = fp32
        eng32.close()
        del eng32
        torch.cuda.empty_cache()

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
        audio_s = audio_s_step * args.steps
        line = {
            "metric": METRIC, "value": audio_s / (ms_total / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": args.precision if args.precision != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": workload_name(my_clips, MAXLEN), "segments_per_gpu": n_seg, "device_batch": dev_batch,
                       "device_batches_per_step": (n_seg + dev_batch - 1) // max(dev_batch, 1),
                       "total_clips": total_clips, "max_length": MAXLEN, "generated_length": out_len,
                       "parallelism": f"clip-sharded x{world}",
                       "l2": "inputs larger than L2 (KV cache per GPU >> 126 MB)"},
            "e2e": {"value": audio_s / (ms_e2e / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": n_seg * SEG_SAMPLES * 4 + n_seg * 16,
                    "d2h_bytes_per_step": n_seg * MAXLEN * 2,
                    "api": "m2m_transcribe_host (C ABI, pinned host buffers, int16 token read-back)",
                    "golden_rows_exact": e2e_golden},
            "tokens_match_golden": golden,
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_by_class": by_class,
            "cpu_baseline": cpu, "extra": extra,
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
