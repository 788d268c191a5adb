"""Markdown tables of the metrics the roofline contract needs from an .ncu-rep (read here, no GPU needed).

    python scripts/ncu_summarize.py gpurun_out/prof_tc.ncu-rep > section.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__cluster_size", "sm__cycles_elapsed.avg.per_second",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("\n| metric | value |\n|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                v = r[i]
                if k == "Kernel Name":
                    v = v.replace("tnb::", "")[:110]
                print(f"| {k} | {v} {units[i]} |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
