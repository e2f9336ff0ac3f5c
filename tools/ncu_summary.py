"""Summarise an .ncu-rep (read here with `ncu -i`) into the JSON kept under profiles/: one record per captured launch."""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "lts_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__cycles_elapsed.avg": "sm_cycles",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefront_pct",
}


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        rec = {"kernel": r[idx["Kernel Name"]][:90]}
        for metric, name in KEEP.items():
            if metric in idx:
                rec[name] = f"{r[idx[metric]]} {units[idx[metric]]}".strip()
        out.append(rec)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
