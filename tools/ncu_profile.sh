# ncu --set full of the kernels of one device-resident C3 step (run on a GPU box: gpurun -- bash tools/ncu_profile.sh [tag] [skip] [count])
# Matching launches per device-resident step: r -> planes, c1 product, c1 finisher, c2 product, c2 finisher, then per decrypt chunk
# narrow, sk -> planes, product, decode phase 1, decode phase 2 (15 with two chunks).  bench.py --steps 2 --warmup 1 runs the guard
# (one device step, one host step ~ 25 matches), two warm-up and two timed device steps before the host-buffer legs.
TAG=${1:-r02b}; SKIP=${2:-55}; COUNT=${3:-15}
set -x
if [ -z "$NCU_SKIP_PLAIN" ]; then python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_plain.log 2>&1 || exit 1; fi
ncu --set full --clock-control none --import-source on -k regex:"decode_fused_claim|ntt_small_kernel|ntt_planes4|ntt_c1_finish|imma_gemm|narrow_i64" -s $SKIP -c $COUNT -o /tmp/${TAG}_prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
if [ -z "$NCU_SKIP_SOURCE" ]; then
  for k in decode_fused_claim ntt_small_kernel ntt_planes4 imma_gemm; do
    ncu -i /tmp/${TAG}_prof.ncu-rep --page source --csv --kernel-name regex:$k > gpurun_out/${TAG}_src_$k.csv 2>/dev/null
  done
  gzip -f gpurun_out/${TAG}_src_*.csv
fi
ls -la gpurun_out/ | tail -8
