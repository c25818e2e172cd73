#!/bin/bash
# Two-GPU run (gpurun --gpus 2): tests/multigpu_check.py (library collectives, peer window, sharded runs vs goldens) and the
# bench at N=2 with and without the NVLink peer window (TNALG_NO_PEER).  Outputs: gpurun_out/mgpu2/.
mkdir -p gpurun_out/mgpu2
( time TNALG_EXPECT_PEER=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py ) > gpurun_out/mgpu2/multigpu_check.log 2>&1
tail -8 gpurun_out/mgpu2/multigpu_check.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu --no-other ) > gpurun_out/mgpu2/bench_2gpu.json 2> gpurun_out/mgpu2/bench_2gpu.err
tail -c 1500 gpurun_out/mgpu2/bench_2gpu.json
( time TNALG_NO_PEER=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu --no-other --no-e2e ) > gpurun_out/mgpu2/bench_2gpu_nopeer.json 2> gpurun_out/mgpu2/bench_2gpu_nopeer.err
tail -c 300 gpurun_out/mgpu2/bench_2gpu_nopeer.err
