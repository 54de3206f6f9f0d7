O=gpurun_out; T=r2j
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 3 > $O/${T}_bench_2gpu.json 2> $O/${T}_bench_2gpu.err; tail -c 1500 $O/${T}_bench_2gpu.json; tail -5 $O/${T}_bench_2gpu.err
python bench.py --steps 3 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; tail -c 3000 $O/${T}_bench.json; tail -5 $O/${T}_bench.err
