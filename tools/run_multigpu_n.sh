#!/bin/bash
# N-GPU run (gpurun --gpus N -- bash tools/run_multigpu_n.sh N): multigpu_check + bench with / without the peer window.
# Outputs: gpurun_out/mgpuN/.
mkdir -p gpurun_out/mgpuN
N=${1:-8}
( time TNALG_EXPECT_PEER=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py ) > gpurun_out/mgpuN/multigpu_check_$N.log 2>&1
tail -4 gpurun_out/mgpuN/multigpu_check_$N.log
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 2 --no-cpu --no-other --no-cfg5 ) > gpurun_out/mgpuN/bench_${N}gpu.json 2> gpurun_out/mgpuN/bench_${N}gpu.err
tail -c 400 gpurun_out/mgpuN/bench_${N}gpu.err
( time TNALG_NO_PEER=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 2 --no-cpu --no-other --no-cfg5 --no-e2e ) > gpurun_out/mgpuN/bench_${N}gpu_nopeer.json 2> gpurun_out/mgpuN/bench_${N}gpu_nopeer.err
tail -c 300 gpurun_out/mgpuN/bench_${N}gpu_nopeer.err
