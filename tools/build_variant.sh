#!/bin/bash
# Build a tuning variant of the CUDA library next to the product one:
#   tools/build_variant.sh <name> "<extra nvcc flags>"   ->  polmux_b200/lib/variants/libpolmux_ssfm_<name>.so
# Run with POLMUX_SSFM_LIB=polmux_b200/lib/variants/libpolmux_ssfm_<name>.so python tools/pass_breakdown.py
set -e
cd "$(dirname "$0")/../polmux_b200/csrc"
mkdir -p ../lib/variants
make -j"$(nproc)" BUILD=build/var_$1 LIB=../lib/variants/libpolmux_ssfm_$1.so EXTRA="$2" 2>&1 | grep -E "error|spill stores, [1-9]" || true
ls -la ../lib/variants/libpolmux_ssfm_$1.so
