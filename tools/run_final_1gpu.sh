#!/bin/bash
# One bench line per BASELINE.json configuration on one B200 (run on a GPU box: gpurun -- bash tools/run_final_1gpu.sh <tag>)
TAG=${1:-r02_final}
OUT=gpurun_out/${TAG}_cfg_1gpu.jsonl
: > $OUT
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err; echo "C3 rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
for args in "--config C1" "--config C2" "--config C3 --parties 8192" "--config C4 --steps 3" "--config C5 --parties 1024" "--config C5 --parties 4096" "--config C5 --parties 16384"; do
  python bench.py --steps 5 --warmup 3 $args >> $OUT 2> gpurun_out/${TAG}_cfg.err; echo "$args rc=$?"
done
