#!/usr/bin/env bash
# GPU box: sanitizer logs and ncu captures of the heavy kernels (text exports only).  usage: tools/evidence.sh <part>
part=${1:-all}
mkdir -p gpurun_out
K='geometry_plugin_matches_reference or forward_field_grid or renderer_training_matches_reference and c8 or renderer_eval'
if [ "$part" = sanitizer ] || [ "$part" = all ]; then
  for tool in racecheck synccheck memcheck; do
    for fam in tcgen05-ws tcgen05-r1; do
      timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_options.py -m gpu -q -x \
          -k "$fam and ($K) or tiny and $fam or grid_line_gather and 8-16-16" > gpurun_out/r02_sanitizer_${tool}_${fam}.log 2>&1
      echo "$tool $fam rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/r02_sanitizer_${tool}_${fam}.log | tr '\n' ' ')"
    done
  done
  # the round-1 kernels built without the memory-phase lock (TT_LIBNAME=libtriplane_b200_nolock.so bash triplaneturbo_b200/csrc/build.sh -DTT_MEMLOCK=0)
  TT_B200_LIB=$PWD/triplaneturbo_b200/lib/libtriplane_b200_nolock.so timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tcgen05-r1 and ($K)" > gpurun_out/r02_sanitizer_racecheck_tcgen05-r1_nolock.log 2>&1
  echo "racecheck r1 nolock rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/r02_sanitizer_racecheck_tcgen05-r1_nolock.log | tr '\n' ' ')"
fi
if [ "$part" = ncu ] || [ "$part" = all ]; then
  for wl in config3q config2q; do
    C=32; [ $wl = config2q ] && C=40
    bash tools/ncu_capture.sh r02_ncu_${wl}_geo_proposal $wl "^k_geo_ws" 2 1 "k_geo_ws<(int)$C, (bool)0" > /dev/null 2>&1
    bash tools/ncu_capture.sh r02_ncu_${wl}_geo_fine $wl "^k_geo_ws" 3 1 "k_geo_ws<(int)$C, (bool)1" > /dev/null 2>&1
    bash tools/ncu_capture.sh r02_ncu_${wl}_tex $wl "^k_tex_tc" 1 1 "k_tex_tc" > /dev/null 2>&1
    bash tools/ncu_capture.sh r02_ncu_${wl}_bwd_geo $wl "^k_bwd_geo_tc" 1 1 "k_bwd_geo_tc" > /dev/null 2>&1
    bash tools/ncu_capture.sh r02_ncu_${wl}_bwd_tex $wl "^k_bwd_tex_tc" 1 1 "k_bwd_tex_tc" > /dev/null 2>&1
    echo "ncu $wl done"
  done
  # launch list of one full default step (shares, not absolutes)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 16 --csv --log-file gpurun_out/r02_launches_config3.csv \
      python bench.py --workload config3 --profile --steps 1 > /dev/null 2>&1
  rm -f gpurun_out/*_raw.csv.tmp
  grep -h "gpu__time_duration\|dram__bytes\|sm__pipe_tensor\|## " gpurun_out/r02_ncu_*_summary.txt | head -80
fi
ls -la gpurun_out | tail -40
du -sh gpurun_out
