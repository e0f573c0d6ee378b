set -x
python bench.py 2>gpurun_out/r2_bench_final_n1.err > gpurun_out/r2_bench_final_n1.json; tail -c 1200 gpurun_out/r2_bench_final_n1.json; tail -3 gpurun_out/r2_bench_final_n1.err
rm -f gpurun_out/*.ncu-rep
bash tools/profile.sh 592 > gpurun_out/profile_sh.log 2>&1
for kn in leaf_elim leaf_back; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_$kn -s 1 -c 1 -f -o gpurun_out/prof_${kn}_A8 python tools/profile_target_swarm.py 8 74 2 > gpurun_out/prof_swarm.log 2>&1
done
python tools/summarize_profiles.py r02 592 > gpurun_out/summarize.log 2>&1
mkdir -p gpurun_out/profiles_r02; cp profiles/r02_ncu_* profiles/r02_launches_B592.* profiles/traffic.json gpurun_out/profiles_r02/
ls gpurun_out/*.ncu-rep | grep -v "proj_lin_pp_B592\|lm_gather16_B592\|leaf_elim_A8" | xargs rm -f
du -sh gpurun_out; ls gpurun_out/profiles_r02 | head -30
