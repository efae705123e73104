#!/bin/bash
# N3 on the GPU box: whole GPU suite, the headline bench line (regression check), then TSM-ResNet50 MTMM steps (configs[4] shape).
tag=${1:-n3b}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1; echo "full pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 120 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_tsm.json 2> gpurun_out/${tag}_bench_tsm.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/${tag}_bench_tsm.json
for b in ${N3_BATCHES:-32 64}; do
  timeout 100 python bench.py --backbone resnet50 --segments 16 --batch $b --steps 5 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_bench_resnet50_b${b}.json 2> gpurun_out/${tag}_bench_resnet50_b${b}.err; echo "resnet b=$b rc=$?"
  cut -c1-600 gpurun_out/${tag}_bench_resnet50_b${b}.json; tail -3 gpurun_out/${tag}_bench_resnet50_b${b}.err
done
