mkdir -p gpurun_out
for k in k_geo_tc k_tex_tc k_bwd_geo_tc k_bwd_tex_tc; do
  n=1; [ $k = k_geo_tc ] && n=2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^${k}" -s $n -c $n -f -o gpurun_out/r02_base_cfg3_$k python bench.py --workload config3 --profile --steps 1 > gpurun_out/r02_base_cfg3_$k.log 2>&1
  echo "$k rc=$?"; tail -3 gpurun_out/r02_base_cfg3_$k.log
done
ls -la gpurun_out/*.ncu-rep
