#!/bin/bash
# symmetric site-specific preorder without the 1/Pi factors: parity tests + A/B against the previous build on the same box
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "site_specific or cfg5 or ss or float_message or mask or sitespec" 2>&1 | tail -4
for rep in 1 2; do
TTB_LIB=$PWD/treetime_b200/libttb_prev.so python tools/ab_pass.py cfg5 20 2>/dev/null | tail -1
python tools/ab_pass.py cfg5 20 2>/dev/null | tail -1
done
