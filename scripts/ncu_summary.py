#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` (and optionally `--page source --csv`) exports into the
compact text kept under profiles/.

    python scripts/ncu_summary.py raw.csv [src.csv] > profiles/<name>.txt
"""
import csv
import sys

csv.field_size_limit(10**9)
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        print("kernel:", r[hdr.index("Kernel Name")][:120])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:72s} {r[i]:>18s} {units[i]}")


def src(path, top=25):
    rows = list(csv.reader(open(path)))
    kernel, hdr, body = None, None, []

    def flush():
        if not body:
            return
        cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[2] or 0) for r in body)
        print("=" * 100)
        print("kernel:", kernel[:120], "| samples", tot, "| SASS instrs", len(body))
        print("  stall totals:", ", ".join(f"{hdr[i][6:]}={sum(int(r[i] or 0) for r in body)}" for i in cols
                                           if sum(int(r[i] or 0) for r in body) > tot * 0.01))
        for r in sorted(body, key=lambda r: -int(r[2] or 0))[:top]:
            why = {hdr[i][6:]: int(r[i] or 0) for i in cols if int(r[i] or 0) > 0.2 * int(r[2] or 1)}
            print(f"  {int(r[2]):7d} {100.0 * int(r[2]) / max(tot, 1):5.1f}%  x{r[5]:>9s}  {r[1].strip()[:70]:70s} {why}")

    for r in rows:
        if r and r[0] == "Kernel Name":
            flush()
            kernel, hdr, body = r[1], None, []
        elif r and r[0] == "Address":
            hdr = r
        elif hdr and len(r) >= len(hdr) - 2 and r[0].startswith("0x"):
            body.append(r)
    flush()


if __name__ == "__main__":
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        src(sys.argv[2])
