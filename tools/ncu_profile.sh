# ncu --set full of the kernels of one device-resident C3 step (run on a GPU box: gpurun -- bash tools/ncu_profile.sh [tag])
TAG=${1:-r02b}
set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"decode_fused_claim|ntt_small_kernel|ntt_planes4|ntt_c1_finish|imma_gemm" -s 18 -c 9 -o /tmp/${TAG}_prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
for k in decode_fused_claim ntt_small_kernel ntt_planes4 imma_gemm; do
  ncu -i /tmp/${TAG}_prof.ncu-rep --page source --csv --kernel-name regex:$k > gpurun_out/${TAG}_src_$k.csv 2>/dev/null
done
gzip -f gpurun_out/${TAG}_src_*.csv
ls -la gpurun_out/ | tail -12
