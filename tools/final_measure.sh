#!/bin/bash
# One GPU box, one call: GPU test suite, headline bench (C3) with every leg, the secondary configs, an ncu launch list of the C3
# step and `--set full` captures of the dominant kernels.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout -k 5 400 python -m pytest tests -m gpu -q --timeout 200 --timeout-method=thread > gpurun_out/pytest_final.txt 2>&1; echo "rc=$?" >> gpurun_out/pytest_final.txt
timeout -k 5 240 python bench.py --steps 20 --warmup 5 --trace gpurun_out/trace_n1_r02.txt > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_final.err; echo "C3 rc=$?" >> gpurun_out/bench_final.err
NCU=/usr/local/cuda/bin/ncu
timeout -k 5 240 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02.csv \
  python bench.py --steps 1 --warmup 1 --graph 0 --no-e2e --no-cpu-baseline --no-gpu-ref > gpurun_out/ncu_bench.log 2>&1
cap() {  # name, kernel regex, command...
  local name=$1 rx=$2; shift 2
  timeout -k 5 90 $NCU --set full --clock-control none --import-source on -k regex:$rx -s 6 -c 2 -f -o gpurun_out/ncu_$name "$@" > gpurun_out/ncu_$name.log 2>&1
}
cap conv_fprop_d2_512 conv_gemm_tc2 python tools/conv_micro.py --shape 16,33,33,512,512,3,1,2 --kind fwd --iters 4
cap conv_dgrad_aspp_d12 conv_gemm_tc2 python tools/conv_micro.py --shape 16,33,33,2048,256,3,1,12 --kind dgrad --iters 4 --beta 1
cap conv_wgrad_aspp_d12 conv_gemm_tc python tools/conv_micro.py --shape 16,33,33,2048,256,3,1,12 --kind wgrad --iters 4
cap conv_fprop_1x1_256_1024 conv_gemm_tc2 python tools/conv_micro.py --shape 16,33,33,256,1024,1,1,1 --kind fwd --iters 4
cap bn_bwd_fused bn_bwd_fused python tools/bn_micro.py --shape 17424,256 --res 1 --iters 4
for c in C2 C4 C5; do
  timeout -k 5 120 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --gpu-ref-steps 10 > gpurun_out/bench_r02_$c.json 2>> gpurun_out/bench_final.err; echo "$c rc=$?" >> gpurun_out/bench_final.err
done
ls -la gpurun_out | tail -30
