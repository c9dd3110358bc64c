"""Summarise an `ncu -i X.ncu-rep --page raw --csv` dump: per captured launch, the metrics the profiles/ summaries quote.
   ncu -i gpurun_out/X.ncu-rep --page raw --csv | python tools/ncu_summary.py [label ...]"""
import csv
import sys

METRICS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
           "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__block_size", "sm__cycles_elapsed.max"]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
labels = sys.argv[1:]
for n, r in enumerate(rows[2:]):
    name = r[hdr.index("Kernel Name")]
    print(f"== {labels[n] if n < len(labels) else n}   [{name[:110]}]")
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print(f"   {m:75s} {r[i]} {units[i]}")
