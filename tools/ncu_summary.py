"""Condense an `ncu --page raw --csv` export into one line per launch (duration, tensor-pipe / DRAM / L2 utilisation,
DRAM bytes).  Usage: python tools/ncu_summary.py raw.csv > summary.csv"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "us"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor_inst_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    cols = [(k, n) for k, n in KEYS if k in hdr]
    extra = [h for h in hdr if "tensor" in h and "pct_of_peak_sustained" in h and h not in [c[0] for c in cols]][:3]
    w = csv.writer(sys.stdout)
    w.writerow(["idx", "kernel"] + [n + ("[" + units[hdr.index(k)] + "]" if units[hdr.index(k)] else "") for k, n in cols] + extra)
    for i, r in enumerate(rows[2:]):
        if len(r) != len(hdr):
            continue
        name = r[hdr.index("Kernel Name")]
        name = name[name.find("conv_gemm_kernel"):][:28] if "conv_gemm_kernel" in name else name[:40]
        w.writerow([i, name] + [r[hdr.index(k)] for k, _ in cols] + [r[hdr.index(h)] for h in extra])


if __name__ == "__main__":
    main(sys.argv[1])
