python scripts_tmp/tiny.py > gpurun_out/tiny_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tiny_launches.csv python scripts_tmp/tiny.py > gpurun_out/tiny_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/tiny_launches.csv')) if len(r)>10]
hdr=rows[0]
ki,vi=hdr.index('Kernel Name'),hdr.index('Metric Value')
for r in rows[1:]:
    print(r[ki][:60], r[vi])
PY
for p in 1 0; do RC_PDL=$p RC_GMAIL=-1 python profiles/tools/quick.py pdl$p 2>&1 | grep -E "easy2k|hard64|hard2k|hard10k"; done
