#!/bin/bash
# usage: scripts/ncu_export.sh <report.ncu-rep> <out_prefix> [launch indices...]  -- raw metrics CSV + per-launch SASS CSVs
set -e
rep=$1; out=$2; shift 2
ncu -i "$rep" --page raw --csv > "${out}_raw.csv"
for i in "$@"; do
  ncu -i "$rep" --page source --csv --print-source sass --launch-skip $i --launch-count 1 > "${out}_src${i}.csv" 2>/dev/null || true
done
