#!/bin/bash
# Multi-GPU evidence pass (run under `gpurun --gpus 8` from the repo root): per-N bench lines through trieste_b200.parallel.
mkdir -p gpurun_out
run() {  # run <nproc> <outfile> <bench args...>
  local n=$1 out=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
    bench.py --gpus $n --no-cpu-baseline --steps 4 --warmup 3 "$@" 2>>gpurun_out/scale_err.txt | grep '^{' | tail -1 > gpurun_out/$out
  cat gpurun_out/$out
}
run 8 r2_scale_weak_8.json
run 8 r2_scale_strong_8.json --scaling strong
run 4 r2_scale_strong_4.json --scaling strong
run 4 r2_scale_weak_4.json
run 8 r2_scale_c4_strong_8.json --config c4 --scaling strong
run 8 r2_scale_c5_strong_8.json --config c5 --scaling strong
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
