"""Summarise ncu captures (run where the .ncu-rep files are readable): python tools/ncu_r2_summary.py rep1 [rep2 ...] -> JSON lines.
Every field carries the unit ncu reports for it."""
import csv
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct_of_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active": "dmma_pipe_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
}
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        rec = {"capture": rep, "kernel": r[hdr.index("Kernel Name")]}
        for m, name in WANT.items():
            if m in hdr:
                i = hdr.index(m)
                rec[name] = f"{r[i]} {units[i]}".strip()
        print(json.dumps(rec))
