#!/usr/bin/env python
"""Markdown summary of an `ncu --set full` capture (the table of profiles/r01_summary_v10.md).

    ncu -i gpurun_out/full.ncu-rep --page raw --csv > full_raw.csv      # here, no GPU needed
    python tools/ncu_summary.py full_raw.csv [--selected out.csv]

Prints one row per captured launch: duration, DRAM read / write, tensor-pipe and issue-slot utilisation, active
warps, registers and the three largest warp-stall reasons; `--selected` also writes the columns worth keeping
under profiles/ (the raw export has ~900 columns).
"""
import csv
import sys

KEEP = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "derived__lts__lts2xbar_bytes.sum.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def to_mb(value, unit):
    v = float(value or 0)
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]

    def g(r, name, default=""):
        return r[col[name]] if name in col else default

    print("| kernel | us | DRAM read / write MB | tensor pipe % | issue slots % | warps active % | regs | top stalls |")
    print("|---|---:|---:|---:|---:|---:|---:|---|")
    for r in data:
        name = g(r, "Kernel Name").split("(")[0].replace("void ", "")[:40]
        rd = to_mb(g(r, "dram__bytes_read.sum"), units[col["dram__bytes_read.sum"]])
        wr = to_mb(g(r, "dram__bytes_write.sum"), units[col["dram__bytes_write.sum"]])
        st = sorted(((float(r[col[s]] or 0), s.replace("smsp__pcsamp_warps_issue_stalled_", "")) for s in stalls), reverse=True)
        tot = sum(v for v, _ in st) or 1.0
        top = ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in st[:3])
        tens = float(g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "0") or 0)
        print(f"| {name} | {float(g(r, 'gpu__time_duration.sum')):.1f} | {rd:.2f} / {wr:.2f} | "
              f"{tens:.1f} | {float(g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f} | "
              f"{float(g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} | "
              f"{g(r, 'launch__registers_per_thread')} | {top} |")
    if "--selected" in sys.argv:
        out = sys.argv[sys.argv.index("--selected") + 1]
        keep = [k for k in KEEP if k in col] + stalls
        with open(out, "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow(keep)
            w.writerow([units[col[k]] for k in keep])
            for r in data:
                w.writerow([r[col[k]] for k in keep])


if __name__ == "__main__":
    main()
