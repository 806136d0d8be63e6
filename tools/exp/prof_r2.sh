set -x
# launch list of a short bench (per-launch gpu time)
python bench.py --steps 1 --warmup 3 --no-cpu --no-mc --no-fp32 --no-e2e --no-configs > gpurun_out/r2_b_plain.json 2> gpurun_out/r2_b_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-mc --no-fp32 --no-e2e --no-configs > gpurun_out/r2_ncu_launch.log 2>&1
# full-set capture of one launch of each pass kernel at the bench's batch (16 realizations in two groups: 8 per launch)
ncu --set full --clock-control none --import-source on -k regex:pmx_k_pass -s 120 -c 3 -o gpurun_out/prof_r2_final python tools/prof_one.py 16 20 > gpurun_out/ncu_r2_final.log 2>&1
# the on-chip kernel on a batch of small fields
ncu --set full --clock-control none --import-source on -k regex:onchip -c 1 -o gpurun_out/prof_r2_onchip python tools/prof_one.py 148 12 > gpurun_out/ncu_r2_onchip.log 2>&1
tail -2 gpurun_out/ncu_r2_final.log gpurun_out/ncu_r2_onchip.log
