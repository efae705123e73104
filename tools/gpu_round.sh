#!/bin/bash
# One GPU session: parity tests, the bench lines and the ncu launch list of the eager step.
# usage: tools/gpu_round.sh <tag> [tests|notests]
tag=${1:-x}; mode=${2:-tests}
mkdir -p gpurun_out
if [ "$mode" = tests ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
  tail -5 gpurun_out/${tag}_pytest.log
fi
timeout 600 python bench.py > gpurun_out/${tag}_bench_tsm.json 2> gpurun_out/${tag}_bench_tsm.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${tag}_bench_tsm.json
timeout 600 python bench.py --temporal action --classes 25 --no-cpu-baseline > gpurun_out/${tag}_bench_action.json 2> gpurun_out/${tag}_bench_action.err; echo "bench action rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
