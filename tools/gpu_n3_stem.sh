#!/bin/bash
# N3 follow-up on the GPU box: the ResNet tests, then the TSM-ResNet50 MTMM step at the configs[4] shape.
tag=${1:-n3c}
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_resnet_gpu.py -q --tb=short -p no:cacheprovider > gpurun_out/${tag}_resnet_pytest.log 2>&1
echo "resnet pytest rc=$?" | tee -a gpurun_out/${tag}_resnet_pytest.log
grep -E "passed|failed|error" gpurun_out/${tag}_resnet_pytest.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/${tag}_resnet_pytest.log | cut -c1-300 | head -20
timeout 100 python bench.py --backbone resnet50 --segments 16 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline \
  > gpurun_out/${tag}_bench_resnet50_b64.json 2> gpurun_out/${tag}_bench_resnet50_b64.err; echo "resnet b=64 rc=$?"
cut -c1-330 gpurun_out/${tag}_bench_resnet50_b64.json; tail -2 gpurun_out/${tag}_bench_resnet50_b64.err
