set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_plain2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"decode_fused_share|ntt_small_kernel|ntt_planes4|imma_gemm" -s 16 -c 8 -o /tmp/r02_prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu2.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02_prof.ncu-rep --page raw --csv > gpurun_out/r02_prof_raw.csv 2>/dev/null
ncu -i /tmp/r02_prof.ncu-rep --page source --csv --kernel-name regex:decode_fused_share > gpurun_out/r02_src_fused.csv 2>/dev/null
ncu -i /tmp/r02_prof.ncu-rep --page source --csv --kernel-name regex:ntt_small_kernel > gpurun_out/r02_src_ntt.csv 2>/dev/null
ncu -i /tmp/r02_prof.ncu-rep --page source --csv --kernel-name regex:imma_gemm > gpurun_out/r02_src_imma.csv 2>/dev/null
gzip -f gpurun_out/r02_src_*.csv
ls -la gpurun_out/
