"""Summary of an `ncu --set full` report as a small CSV for profiles/: one line per captured launch with the metrics the
rooflines in bench.py / DESIGN.md quote (duration, DRAM bytes read / written, L2 hit rate, SM and tensor-pipe
utilisation, achieved occupancy, registers). Reads the report with `ncu -i REPORT --page raw --csv`.

usage: ncu_subset.py REPORT.ncu-rep > profiles/<name>_raw_subset.csv"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
           "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__cluster_size", "sm__cycles_elapsed.avg.per_second"]

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], check=True, capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units, body = rows[0], rows[1], rows[2:]
keep = [head.index(c) for c in ("Kernel Name", "Grid Size", "Block Size") if c in head]
keep += [head.index(m) for m in METRICS if m in head]
out = csv.writer(sys.stdout)
out.writerow([head[i] for i in keep])
out.writerow([units[i] for i in keep])
for r in body:
    out.writerow([r[i] for i in keep])
