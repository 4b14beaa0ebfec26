"""ncu -i <rep> --page raw --csv  ->  a compact JSON summary (one entry per captured launch) for profiles/."""
import csv
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def main(rep, out):
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    head, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[head.index("Kernel Name")][:120]}
        for k in KEEP:
            if k in head:
                i = head.index(k)
                d[k] = "%s %s" % (r[i], units[i])
        launches.append(d)
    with open(out, "w") as f:
        json.dump({"source": "ncu --set full --clock-control none --import-source on (cold caches, serialised launches)",
                   "launches": launches}, f, indent=1)
    print("%d launches -> %s" % (len(launches), out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
