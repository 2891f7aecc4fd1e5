#!/usr/bin/env bash
# GPU box helper: ncu --set full capture of ONE kernel of a bench workload, exported as text (reports stay on the box:
# gpurun_out/ may not exceed 64 MiB).   usage: tools/ncu_capture.sh <tag> <workload> <kernel regex> <skip> <count> [pattern for ncu_lines]
tag=$1; wl=$2; kre=$3; skip=$4; cnt=$5; pat=${6:-$3}
mkdir -p gpurun_out
rep=/tmp/${tag}.ncu-rep
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$kre" -s $skip -c $cnt -f -o /tmp/${tag} \
    python bench.py --workload $wl --profile --steps 1 > gpurun_out/${tag}.log 2>&1
echo "ncu rc=$?"
ncu -i $rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
python tools/ncu_summary.py $rep 25 > gpurun_out/${tag}_summary.txt 2>&1
python tools/ncu_lines.py $rep "$pat" 70 > gpurun_out/${tag}_lines.txt 2>&1
cat gpurun_out/${tag}_summary.txt; head -75 gpurun_out/${tag}_lines.txt
