"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline argument uses.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx.txt   (runs here, no GPU needed)"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {path}")
    for r in rows[2:]:
        print(f"\n## {r[hdr.index('Kernel Name')][:110]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:75s} {r[i]:>16s} {units[i]}")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
            print(f"{'traffic = dram read + write':75s} {rd + wr:16.3f} (read unit: {units[hdr.index('dram__bytes_read.sum')]}, "
                  f"write unit: {units[hdr.index('dram__bytes_write.sum')]})")
        except Exception:
            pass


if __name__ == "__main__":
    main(sys.argv[1])
