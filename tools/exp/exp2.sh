set -x
V=polmux_b200/lib/variants
python tools/pass_breakdown.py 8 20 1
for n in nophys nofft none; do POLMUX_SSFM_LIB=$V/libpolmux_ssfm_$n.so python tools/pass_breakdown.py 8 20 1; done
POLMUX_SSFM_LIB=$V/libpolmux_ssfm_timing.so python tools/phase_timing.py
tools/ubench/fp64_rate
ncu --set full --clock-control none --import-source on -k regex:passB -s 40 -c 2 -o gpurun_out/prof_r2_1 python tools/prof_one.py 8 20 > gpurun_out/ncu_r2_1.log 2>&1
tail -3 gpurun_out/ncu_r2_1.log
