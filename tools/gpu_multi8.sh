#!/bin/bash
# configs #3 / #4 of BASELINE.json on N GPUs: SD stage-2 and ACTION-MobileNetV2 (25 classes); usage: tools/gpu_multi8.sh <tag> <n>
tag=${1:-x}; n=${2:-8}
mkdir -p gpurun_out
run() {
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $n --steps 20 --warmup 5 $2 > gpurun_out/${tag}_${1}_${n}gpu.json 2> gpurun_out/${tag}_${1}_${n}gpu.err
  echo "$1 rc=$?"; tail -c 300 gpurun_out/${tag}_${1}_${n}gpu.json
}
run sd_tsm "--workload sd"
run mtmm_action25 "--temporal action --classes 25"
run mtmm_tsm ""
