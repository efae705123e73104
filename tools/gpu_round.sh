#!/bin/bash
# One GPU session: parity tests, the bench lines, the ncu launch list of the eager step and --set full captures of the
# top kernels.   usage: tools/gpu_round.sh <tag> [tests|notests] [full|nofull]
tag=${1:-x}; mode=${2:-tests}; full=${3:-full}
mkdir -p gpurun_out
if [ "$mode" = tests ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
  tail -5 gpurun_out/${tag}_pytest.log
fi
timeout 600 python bench.py > gpurun_out/${tag}_bench_tsm.json 2> gpurun_out/${tag}_bench_tsm.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/${tag}_bench_tsm.json
timeout 600 python bench.py --temporal action --classes 25 > gpurun_out/${tag}_bench_action.json 2> gpurun_out/${tag}_bench_action.err; echo "bench action rc=$?"
timeout 600 python bench.py --workload sd > gpurun_out/${tag}_bench_sd.json 2> gpurun_out/${tag}_bench_sd.err; echo "bench sd rc=$?"
timeout 600 python bench.py --workload mtmm_sd --no-cpu-baseline > gpurun_out/${tag}_bench_mtmm_sd.json 2> gpurun_out/${tag}_bench_mtmm_sd.err; echo "bench mtmm_sd rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
if [ "$full" = full ]; then
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:dw_bwd_sw_kernel -s 1 -c 1 -o gpurun_out/${tag}_full_dw_bwd \
    python tools/bench_kernels.py --only dw_bwd --reps 1 > gpurun_out/${tag}_full_dw_bwd.log 2>&1; echo "full dw_bwd rc=$?"
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:pw_gemm_tc_kernel -s 1 -c 3 -o gpurun_out/${tag}_full_pw_gemm \
    python tools/bench_kernels.py --only pw_fwd,pw_proj,pw_dgrad1 --reps 1 > gpurun_out/${tag}_full_pw_gemm.log 2>&1; echo "full gemm rc=$?"
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:pw_wgrad_tc_kernel -s 1 -c 1 -o gpurun_out/${tag}_full_pw_wgrad \
    python tools/bench_kernels.py --only pw_wgrad1 --reps 1 > gpurun_out/${tag}_full_pw_wgrad.log 2>&1; echo "full wgrad rc=$?"
fi
