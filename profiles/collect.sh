#!/bin/bash
# Round-end evidence, run on the GPU box through gpurun (one ncu tool per call):
#   bash profiles/collect.sh launches   -> plain bench run, then the ncu launch list of the same command
#   bash profiles/collect.sh full       -> plain driver run, then ncu --set full of every kernel of the path
set -e
mkdir -p gpurun_out
case "$1" in
launches)
  python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_launches.log 2>&1
  ;;
full)
  # every kernel of the path; the report (> 64 MiB) stays on the box, its summaries come back.
  # The two descent kernels are captured again on their own, with sources, into small reports that do.
  python profiles/prof_driver.py all > gpurun_out/plain.log 2>&1 &&
  ncu --set full --clock-control none \
      -k regex:"classify|check_culled|check_brute|check_spheres|transform_soa|transform_aos|aabb_|narrow_soa|broad_kernel|gather_|compact_|conf_generate" -c 60 \
      -f -o /tmp/prof_full python profiles/prof_driver.py all > gpurun_out/ncu_full.log 2>&1 &&
  python profiles/summarize.py /tmp/prof_full.ncu-rep "ncu summary, round ${ROUND:-2} (B200, sm_100a)" gpurun_out/ncu_metrics.json > gpurun_out/ncu_summary.md
  for m in culled gobs; do
    ncu --set full --clock-control none --import-source on -k regex:"check_culled" -c 2 \
        -f -o gpurun_out/prof_$m python profiles/prof_driver.py $m > gpurun_out/ncu_$m.log 2>&1
  done
  ;;
one)  # collect.sh one <driver mode> <kernel regex> <count>: one kernel family, e.g. `one gobs check_culled 3`
  python profiles/prof_driver.py $2 > gpurun_out/plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"$3" -c ${4:-3} \
      -o gpurun_out/prof_$2 python profiles/prof_driver.py $2 > gpurun_out/ncu_$2.log 2>&1
  ;;
esac
tail -n 3 gpurun_out/*.log
