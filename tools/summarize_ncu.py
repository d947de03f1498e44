#!/usr/bin/env python
"""Reduce an .ncu-rep (ncu --set full) to the handful of metrics DESIGN.md / bench.py quote. Runs on the CPU box:
    python tools/summarize_ncu.py gpurun_out/prof_gemm_r01.ncu-rep > profiles/r01_conv_gemm_ncu.csv"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [i for k in KEYS for i, h in enumerate(hdr) if h == k]
    w = csv.writer(sys.stdout)
    w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in cols])
    for r in rows[2:]:
        w.writerow([r[i][:90] for i in cols])


if __name__ == "__main__":
    main(sys.argv[1])
