#!/bin/bash
bash profiles/collect.sh full
ls -la gpurun_out /tmp/prof_full.ncu-rep
bash scripts_tmp/small_sweep.sh > gpurun_out/small_sweep.jsonl 2> gpurun_out/small_sweep.err
cat gpurun_out/small_sweep.jsonl
