#!/bin/bash
# Round evidence on one B200: tests, smoke, bench (all configs, both arms), launch list with DRAM bytes, per-launch GEMM table,
# ncu --set full of a few launches of the dominant kernels (exported as CSV: the .ncu-rep files stay on the box).
# usage: scripts/final_capture.sh <tag>
tag=$1
mkdir -p gpurun_out /tmp/n
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/${tag}_tests.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -8) > gpurun_out/${tag}_smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
timeout 300 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_config3.json 2> gpurun_out/${tag}_bench_config3.err
timeout 400 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/${tag}_bench_train.json 2> gpurun_out/${tag}_bench_train.err
timeout 300 python scripts/config5_sweep.py > gpurun_out/${tag}_config5.md 2>&1
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --profile-step > /tmp/n/prof.log 2>&1
timeout 300 python scripts/h3_launch_table.py > gpurun_out/${tag}_h3_table.txt 2>&1
# --set full: U-Net 3x3 convolutions (cta_group::2 form), the candidate chain kernel, attention + fp16 gather
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:linear_h3 --launch-skip 57 --launch-count 2 -o /tmp/n/h3a python bench.py --profile-step > /tmp/n/a.log 2>&1
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:sdf_chain --launch-count 1 -o /tmp/n/chain python bench.py --profile-step > /tmp/n/b.log 2>&1
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"attention_tc128|gather_sum_h16" --launch-count 2 -o /tmp/n/att python bench.py --profile-step > /tmp/n/c.log 2>&1
for f in h3a chain att; do ncu -i /tmp/n/$f.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_${f}_raw.csv 2>/dev/null; done
tail -3 gpurun_out/${tag}_tests.log; tail -3 gpurun_out/${tag}_smoke.log; cut -c1-300 gpurun_out/${tag}_bench.json; cut -c1-300 gpurun_out/${tag}_bench_reference.json; ls -la gpurun_out | grep ${tag}
