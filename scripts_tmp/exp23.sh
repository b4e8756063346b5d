#!/bin/bash
python profiles/prof_driver.py short > gpurun_out/plain_short.log 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:"check_culled" -c 4 -f -o /tmp/prof_short python profiles/prof_driver.py short > gpurun_out/ncu_short.log 2>&1
ncu -i /tmp/prof_short.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
want=['Kernel Name','launch__block_size','gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sass__inst_executed_register_spilling']
for r in rows[2:]:
    print({w.split('.')[0][-28:]: r[h.index(w)][:60] for w in want if w in h})
"
