O=gpurun_out; T=r2y; N=$1
python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/${T}_pytest_mgpu_${N}.log 2>&1; tail -3 $O/${T}_pytest_mgpu_${N}.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 5 --warmup 3 > $O/${T}_bench_${N}gpu.json 2> $O/${T}_bench_${N}gpu.err; tail -c 700 $O/${T}_bench_${N}gpu.json; tail -2 $O/${T}_bench_${N}gpu.err
