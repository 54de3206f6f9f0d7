O=gpurun_out; T=r2z; N=$1
if [ "$N" = "1" ]; then python bench.py --config5 > $O/${T}_config5_1gpu.jsonl 2> $O/${T}_config5_1gpu.err; else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --config5 > $O/${T}_config5_${N}gpu.jsonl 2> $O/${T}_config5_${N}gpu.err; fi
tail -3 $O/${T}_config5_${N}gpu.jsonl | cut -c1-250; tail -2 $O/${T}_config5_${N}gpu.err
