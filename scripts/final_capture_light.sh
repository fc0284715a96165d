#!/bin/bash
# Final evidence of the build as shipped (no ncu): tests, smoke, bench of every config and the reference arm.
# usage: scripts/final_capture_light.sh <tag>
tag=$1
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/${tag}_tests.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -8) > gpurun_out/${tag}_smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 2> gpurun_out/${tag}_bench.err | grep "^{" > gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2> /dev/null | grep "^{" > gpurun_out/${tag}_bench_reference.json
timeout 300 python bench.py --config 3 --steps 5 --warmup 3 2> /dev/null | grep "^{" > gpurun_out/${tag}_bench_config3.json
timeout 400 python bench.py --config 4 --steps 5 --warmup 3 2> /dev/null | grep "^{" > gpurun_out/${tag}_bench_train.json
tail -3 gpurun_out/${tag}_tests.log; tail -4 gpurun_out/${tag}_smoke.log
for f in bench bench_reference bench_config3 bench_train; do cut -c1-260 gpurun_out/${tag}_$f.json; done
