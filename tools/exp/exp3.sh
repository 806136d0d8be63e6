V=polmux_b200/lib/variants
for n in gb2 gb2none gb4 gb4none; do echo "== $n"; PMX_VERBOSE=1 POLMUX_SSFM_LIB=$V/libpolmux_ssfm_$n.so python tools/pass_breakdown.py 8 20 1 2>&1 | tail -3; done
echo "== gb2 span"; POLMUX_SSFM_LIB=$V/libpolmux_ssfm_gb2.so python tools/span_time.py 16 20 4
