for g in 1 2 3 4; do PMX_GROUPS=$g python tools/span_time.py 16 20 3; done
for b in 12 24 32; do python tools/span_time.py $b 20 3; done
PMX_GROUPS=3 python tools/span_time.py 24 20 3
PMX_GROUPS=4 python tools/span_time.py 32 20 3
