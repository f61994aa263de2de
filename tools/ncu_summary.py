import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','l1tex__t_bytes.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','launch__grid_size','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','launch__occupancy_limit_registers','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__inst_executed_pipe_fp64.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']
idx={h:i for i,h in enumerate(hdr)}
for r in rows[2:]:
    print('---')
    for w in want:
        if w in idx: print(w, '=', r[idx[w]][:120], units[idx[w]])
