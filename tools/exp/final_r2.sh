set -x
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2f_gpu_tests.txt 2>&1
tail -14 gpurun_out/r2f_gpu_tests.txt
python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
tail -c 600 gpurun_out/r2f_bench_n1.err
python bench.py --steps 1 --warmup 3 --no-cpu --no-mc --no-fp32 --no-e2e --no-configs > gpurun_out/r2f_b_plain.json 2> gpurun_out/r2f_b_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-mc --no-fp32 --no-e2e --no-configs > gpurun_out/r2f_ncu_launch.log 2>&1
python tools/prof_one.py 16 20 > gpurun_out/r2f_prof_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pmx_k_pass -s 120 -c 3 -o gpurun_out/prof_r2f python tools/prof_one.py 16 20 > gpurun_out/ncu_r2f.log 2>&1
tail -2 gpurun_out/ncu_r2f.log
