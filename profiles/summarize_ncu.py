"""Summarise `ncu --page raw --csv` exports (profiles/ncu_all.sh) into a per-kernel table.
usage: python profiles/summarize_ncu.py <raw.csv> [more.csv ...]   -> prints a table, returns rows for ncu_kernels.json"""
import csv
import json
import re
import sys

WANT = {
    "gpu__time_duration.sum": "dur_ns",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct_active",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "smsp__cycles_active.avg": "cycles",
    "sm__cycles_elapsed.max": "cycles_elapsed",
    "gpc__cycles_elapsed.avg.per_second": "clk_hz",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tmem_pct",
}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "second": 1e9,
        "hz": 1, "Khz": 1e3, "Mhz": 1e6, "Ghz": 1e9}


def load(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": re.sub(r"^void |ddpm::", "", r[hdr.index("Kernel Name")])[:110]}
        for i, h in enumerate(hdr):
            if h in WANT and r[i] not in ("", "n/a"):
                try:
                    d[WANT[h]] = float(r[i].replace(",", "")) * UNIT.get(units[i], 1)
                except ValueError:
                    pass
        out.append(d)
    return out


if __name__ == "__main__":
    allrows = []
    for p in sys.argv[1:]:
        allrows += load(p)
    print(f"{'us':>8} {'DRAM rd MB':>10} {'wr MB':>8} {'GB/s':>7} {'dram%':>6} {'tensor%':>7} {'tmem%':>6} {'sm%':>5} {'regs':>4} {'GHz':>5}  kernel")
    for d in allrows:
        us = d.get("dur_ns", 0) / 1e3
        rd, wr = d.get("dram_rd", 0) / 1e6, d.get("dram_wr", 0) / 1e6
        gbs = (rd + wr) / us * 1e3 if us else 0
        print(f"{us:8.1f} {rd:10.1f} {wr:8.1f} {gbs:7.0f} {d.get('dram_pct', 0):6.1f} {d.get('tensor_pct_active', 0):7.1f} "
              f"{d.get('tmem_pct', 0):6.1f} {d.get('sm_pct', 0):5.1f} {int(d.get('regs', 0)):4d} {d.get('clk_hz', 0) / 1e9:5.2f}  {d['kernel']}")
    json.dump(allrows, open("/tmp/ncu_rows.json", "w"), indent=1)
