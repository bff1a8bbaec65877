#!/bin/bash
# Sample-level sweep for the float configs C1 / C2 ($@ = level strings)
cd "$(dirname "$0")/../.."
for lv in "$@"; do
  envs="GPANN_X=1"; [ $lv != default ] && envs="GPANN_TC_LEVELS=$lv"
  echo "== levels $lv"; env $envs bash profiles/tools/r2_c12.sh
done
