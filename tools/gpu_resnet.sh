#!/bin/bash
# N3 check on the GPU box: the ResNet tests alone (all failures shown, not -x).   usage: tools/gpu_resnet.sh <tag> [full]
tag=${1:-n3}
mkdir -p gpurun_out
timeout ${N3_TIMEOUT:-150} python -m pytest tests/test_resnet_gpu.py -q --tb=short -p no:cacheprovider > gpurun_out/${tag}_resnet_pytest.log 2>&1
echo "resnet pytest rc=$?" | tee -a gpurun_out/${tag}_resnet_pytest.log
grep -E "passed|failed|error" gpurun_out/${tag}_resnet_pytest.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/${tag}_resnet_pytest.log | cut -c1-300 | head -40
if [ "$2" = full ]; then
  timeout 200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1; echo "full pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
  tail -3 gpurun_out/${tag}_pytest.log
fi
