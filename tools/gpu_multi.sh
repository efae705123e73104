#!/bin/bash
# Multi-GPU bench lines (one rank per GPU, NCCL): usage tools/gpu_multi.sh <tag> <ngpus>
tag=${1:-x}; n=${2:-2}
mkdir -p gpurun_out
run() {  # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 20 --warmup 5 $2 > gpurun_out/${tag}_${1}_${n}gpu.json 2> gpurun_out/${tag}_${1}_${n}gpu.err
  echo "$1 rc=$?"; tail -c 600 gpurun_out/${tag}_${1}_${n}gpu.json | cut -c1-400
}
run mtmm_tsm ""
run mtmm_action25 "--temporal action --classes 25"
run sd_tsm "--workload sd"
