#!/bin/bash
# Run on the GPU box (gpurun).  Produces the per-launch list and full ncu captures of the hot kernels.
set -x
B=${1:-256}
K='regex:k_(lm_gather|schur|chol|step|misc_lin|proj_lin|control)'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 40 -c 60 --csv --log-file gpurun_out/launches_B${B}.csv python tools/profile_target.py $B 3 > gpurun_out/prof_target.log 2>&1
for kn in proj_lin chol lm_gather schur step misc_lin; do
  ncu --set full --clock-control none --import-source on -k "regex:k_${kn}" -s 4 -c 1 -f -o gpurun_out/prof_${kn}_B${B} python tools/profile_target.py $B 3 >> gpurun_out/prof_target.log 2>&1
done
ls -la gpurun_out/
