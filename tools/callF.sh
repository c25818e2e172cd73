mkdir -p gpurun_out/F
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/F/pytest_gpu.log 2>&1
tail -4 gpurun_out/F/pytest_gpu.log
for chi in 256 512; do
echo "fast ritz"; python tools/profile_lanczos.py --chi $chi --solves 10 2>&1 | tail -1
echo "QL ritz"; TNALG_RITZ_QL=1 python tools/profile_lanczos.py --chi $chi --solves 10 2>&1 | tail -1
done
python bench.py --workload heis_chain100_chi256 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/F/bench_chi256.json 2> gpurun_out/F/bench_chi256.err
tail -c 900 gpurun_out/F/bench_chi256.json
