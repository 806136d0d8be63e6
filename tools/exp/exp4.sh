for b in 2 3 4 6; do PMX_GROUPS=1 python tools/span_time.py $b 20 4; done
for b in 3 4; do PMX_GROUPS=2 python tools/span_time.py $b 20 4; done
PMX_GROUPS=1 python tools/pass_breakdown.py 3 20 1
python tools/pass_breakdown.py 8 20 1
