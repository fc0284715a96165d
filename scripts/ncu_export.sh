#!/bin/bash
# usage: scripts/ncu_export.sh <report.ncu-rep> <out_prefix>   -- raw metrics CSV + per-launch source (SASS) CSVs
set -e
rep=$1; out=$2
ncu -i "$rep" --page raw --csv > "${out}_raw.csv"
n=$(ncu -i "$rep" --page raw --csv | tail -n +3 | wc -l)
for i in $(seq 0 $((n-1))); do
  ncu -i "$rep" --page source --csv --print-source sass --launch-skip $i --launch-count 1 > "${out}_src${i}.csv" 2>/dev/null || true
done
