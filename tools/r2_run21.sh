set -x
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
python bench.py 2>gpurun_out/r2_bench_final_n1.err > gpurun_out/r2_bench_final_n1.json; tail -c 600 gpurun_out/r2_bench_final_n1.json; tail -3 gpurun_out/r2_bench_final_n1.err
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
bash tools/profile.sh 592 > gpurun_out/profile_sh.log 2>&1
for kn in leaf_elim leaf_back; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_$kn -s 1 -c 1 -f -o gpurun_out/prof_${kn}_A8 python tools/profile_target_swarm.py 8 74 2 > gpurun_out/prof_swarm.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -14
