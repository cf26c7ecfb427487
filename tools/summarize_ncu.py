"""Condenses `ncu -i x.ncu-rep --page raw --csv` exports (tools/r2_profile.sh) into one small JSON per capture:
the metrics the roofline discussion needs (duration, DRAM / L2 bytes, pipe utilisation, occupancy, stall mix).
Usage: python tools/summarize_ncu.py gpurun_out/r2_ncu_*.csv -o profiles/"""
import csv
import json
import os
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct_of_peak",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read_bytes",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__cluster_size": "cluster_size",
    "launch__waves_per_multiprocessor": "waves_per_sm",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem_blocks",
    "launch__occupancy_limit_registers": "occupancy_limit_regs_blocks",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__cycles_elapsed.max": "sm_cycles",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "Kbyte/block": 1e3, "byte/block": 1.0,
              "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}


def summarize(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
    if not hdr:
        return None
    h, u, d = rows[hdr[0]], rows[hdr[0] + 1], rows[hdr[0] + 2]
    out = {"kernel": d[h.index("Kernel Name")][:160], "source": os.path.basename(path)}
    stalls = {}
    for i, name in enumerate(h):
        if name in KEYS and i < len(d) and d[i]:
            try:
                v = float(d[i].replace(",", ""))
            except ValueError:
                continue
            out[KEYS[name]] = v * UNIT_SCALE.get(u[i], 1.0) if u[i] in UNIT_SCALE else v
        if name.startswith("smsp__average_warps_issue_stalled_") and name.endswith("_per_issue_active.ratio") and d[i]:
            try:
                stalls[name[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(d[i].replace(",", ""))
            except ValueError:
                pass
    if stalls:
        out["top_stalls_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
    if "dram_read_bytes" in out and "duration_us" in out:
        out["dram_gbs"] = (out["dram_read_bytes"] + out.get("dram_write_bytes", 0.0)) / out["duration_us"] / 1e3
    if "l2_to_sm_read_bytes" in out and "duration_us" in out:
        out["l2_to_sm_gbs"] = out["l2_to_sm_read_bytes"] / out["duration_us"] / 1e3
    return out


if __name__ == "__main__":
    args = sys.argv[1:]
    odir = "profiles"
    if "-o" in args:
        odir = args[args.index("-o") + 1]
        args = [a for a in args if a not in ("-o", odir)]
    for p in args:
        s = summarize(p)
        if s is None:
            print("skip", p)
            continue
        name = os.path.splitext(os.path.basename(p))[0] + "_summary.json"
        json.dump(s, open(os.path.join(odir, name), "w"), indent=1)
        print(name, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in s.items() if k not in ("kernel", "source", "top_stalls_warps_per_issue")})
