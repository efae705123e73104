#!/bin/bash
# configs #2 / #3 / #4 of BASELINE.json on N GPUs; usage: tools/gpu_multi8.sh <tag> <n> [workloads: tsm sd action]
tag=${1:-x}; n=${2:-8}; shift 2; what=${@:-"sd action tsm"}
mkdir -p gpurun_out
run() {
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $n --steps 20 --warmup 5 $2 > gpurun_out/${tag}_${1}_${n}gpu.json 2> gpurun_out/${tag}_${1}_${n}gpu.err
  echo "$1 rc=$?"; tail -c 300 gpurun_out/${tag}_${1}_${n}gpu.json
}
for w in $what; do
  case $w in
    sd) run sd_tsm "--workload sd" ;;
    action) run mtmm_action25 "--temporal action --classes 25" ;;
    tsm) run mtmm_tsm "" ;;
  esac
done
