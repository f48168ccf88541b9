#!/bin/bash
# C4 / C5 strong scaling at N = 1 and 2 (run under `gpurun --gpus 2`).
mkdir -p gpurun_out
run() {
  local n=$1 out=$2; shift 2
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
    bench.py --gpus $n --no-cpu-baseline --steps 3 --warmup 3 "$@" 2>>gpurun_out/scale_err.txt | grep '^{' | tail -1 > gpurun_out/$out
  cut -c1-260 gpurun_out/$out
}
run 1 r2_scale_c4_strong_1.json --config c4 --scaling strong
run 2 r2_scale_c4_strong_2.json --config c4 --scaling strong
run 1 r2_scale_c5_strong_1.json --config c5 --scaling strong
run 2 r2_scale_c5_strong_2.json --config c5 --scaling strong
tail -5 gpurun_out/scale_err.txt
