#!/bin/bash
# final check of a tree on the GPU box: whole GPU suite, smoke(), one TSM-ResNet50 MTMM bench line (configs[4] shape)
tag=${1:-fin}
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1; echo "full pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -2 gpurun_out/${tag}_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
timeout 80 python bench.py --backbone resnet50 --segments 16 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline \
  > gpurun_out/${tag}_bench_resnet50_b64.json 2> gpurun_out/${tag}_bench_resnet50_b64.err; echo "resnet b=64 rc=$?"
cut -c1-330 gpurun_out/${tag}_bench_resnet50_b64.json
