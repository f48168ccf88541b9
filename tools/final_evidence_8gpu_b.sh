#!/bin/bash
# Second multi-GPU pass (run under `gpurun --gpus 8`): NCCL world-2 parity tests, then C4 / C5 strong scaling at 8 GPUs
# after the small-batch optimiser rounds and the batched host-side gathers.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu --timeout 500 2>&1 | tail -3 > gpurun_out/r2_gputest_multi.txt
cat gpurun_out/r2_gputest_multi.txt
run() {
  local n=$1 out=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
    bench.py --gpus $n --no-cpu-baseline --steps 4 --warmup 3 "$@" 2>>gpurun_out/scale_err.txt | grep '^{' | tail -1 > gpurun_out/$out
  cut -c1-330 gpurun_out/$out
}
run 8 r2_scale_c4_strong_8b.json --config c4 --scaling strong
run 8 r2_scale_c5_strong_8b.json --config c5 --scaling strong
