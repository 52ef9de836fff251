"""Turn `ncu -i X.ncu-rep --page raw --csv` output into the small JSON summary committed under profiles/.

    python tools/ncu_summary.py raw.csv out.json "<command that was profiled>"
"""
import csv
import json
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_bytes.sum", "smsp__cycles_active.avg")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(raw, out, command):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    launches = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEEP or h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("_not_issued"):
                d[h] = {"value": r[i], "unit": units[i]}
        launches.append(d)

    def nbytes(d, k):
        return float(d[k]["value"].replace(",", "")) * UNIT.get(d[k]["unit"], 1)
    per = [nbytes(d, "dram__bytes_read.sum") + nbytes(d, "dram__bytes_write.sum") for d in launches]
    json.dump({"command": command, "launches": launches, "dram_bytes_per_launch": sum(per) / len(per)}, open(out, "w"), indent=1)
    print(out, "dram_bytes_per_launch", sum(per) / len(per))


if __name__ == "__main__":
    main(*sys.argv[1:4])
