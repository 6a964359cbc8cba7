#!/bin/bash
# Times bench.py under different tuning knobs (PVW_OPTS).  Usage (on a GPU box): bash tools/sweep_opts.sh "opt=a,opt2=b" ...
mkdir -p gpurun_out
for opts in "$@"; do
  PVW_OPTS=$opts timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/sw.json 2> gpurun_out/sw.err || tail -3 gpurun_out/sw.err
  python -c "
import json; d=json.load(open('gpurun_out/sw.json')); print('$opts', round(d['value']/1e6,2), 'M/s  e2e', round(d['e2e']['value']/1e6,2), ' step', round(d['ms_per_step'],2), 'ms ', d['roofline']['kernel_ms_per_step'], ' MAC/s', round(d['roofline']['modmuladds_per_s']/1e12,3))"
done
