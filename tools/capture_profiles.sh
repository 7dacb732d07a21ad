#!/bin/bash
# ncu evidence of one build (run under gpurun, one GPU): the launch list of one eager config-2 step, `--set full`
# captures of the persistent kernels + the dP kernel of that step, and of the per-timestep path's tensor-core energy
# backward (12 free-running decoder steps). Each ncu run follows a plain run of the same command that exited 0.
# Usage: tools/capture_profiles.sh <tag>   ->  gpurun_out/<tag>_{launches.csv,launches_summary.txt,ncu_full_summary.txt,ncu_ebwd_summary.txt}
set -u
tag=${1:-rXX}
out=gpurun_out
python bench.py --profile-step > $out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/${tag}_launches.csv python bench.py --profile-step > $out/${tag}_ncu_launches.log 2>&1
python tools/summarize_launches.py $out/${tag}_launches.csv 45 > $out/${tag}_launches_summary.txt 2>&1
ncu --set full --clock-control none --profile-from-start off -k regex:"lstm_persist|dec_persist|att_dp_mma" \
    -o $out/${tag}_full python bench.py --profile-step > $out/${tag}_ncu_full.log 2>&1
python tools/ncu_summary.py $out/${tag}_full.ncu-rep > $out/${tag}_ncu_full_summary.txt 2>&1
python tools/ssl_one_pass.py > $out/${tag}_plain_ssl.log 2>&1 || { echo "plain ssl run failed"; exit 1; }
ncu --set full --clock-control none --profile-from-start off -k regex:"att_energy_bwd_mma" -c 2 \
    -o $out/${tag}_ebwd python tools/ssl_one_pass.py > $out/${tag}_ncu_ebwd.log 2>&1
python tools/ncu_summary.py $out/${tag}_ebwd.ncu-rep > $out/${tag}_ncu_ebwd_summary.txt 2>&1
tail -3 $out/${tag}_launches_summary.txt
