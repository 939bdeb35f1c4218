#!/bin/bash
# Round-2 ncu evidence for profiles/ (run on the GPU box, one GPU):
#   1. launch list of the headline bench command, kernels launched one by one (no graph replay, optimizer at the end of its
#      own step so that launch order = program order):  gpurun_out/${TAG}_launches.csv
#   2. `ncu --set full` over three consecutive steps of the same command (a TV iteration is every third): gpurun_out/${TAG}_full.ncu-rep
# Numbers printed by bench.py under ncu are not bench values.
TAG=${1:-r02}
ARGS="bench.py --no-graph --no-defer --steps 6 --warmup 3 --no-cpu-baseline --no-parity-check --sustain 0"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python $ARGS > gpurun_out/${TAG}_launches.log 2>&1
read SKIP COUNT < <(python - <<EOF
import csv, re
rows = list(csv.reader(open('gpurun_out/${TAG}_launches.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; ki = h.index('Kernel Name')
names = [r[ki] for r in rows[hdr + 1:] if len(r) > ki]
idx = [i for i, n in enumerate(names) if 'k_march_flags' in n]
# steps 4..6 of the timed region: calibrate + 12 pre-steps + 3 warm-up come first; take the three steps before the last 30 % of the list
k = int(len(idx) * 0.6)
print(idx[k], idx[k + 3] - idx[k])
EOF
)
echo "full capture: skip $SKIP count $COUNT"
timeout 1500 ncu --set full --clock-control none -s $SKIP -c $COUNT -f -o gpurun_out/${TAG}_full python $ARGS > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv
ls -la gpurun_out/${TAG}_full.ncu-rep gpurun_out/${TAG}_full_raw.csv
# the report itself (> 100 MB for ~80 launches) does not fit the 64 MiB return channel: the raw CSV page carries every metric
rm -f gpurun_out/${TAG}_full.ncu-rep
