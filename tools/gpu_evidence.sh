#!/bin/bash
# Round evidence on one B200 (run under gpurun): tests, both bench arms, the ncu launch list of the bench command,
# full ncu captures of the top kernels (raw pages exported here: gpurun merges at most 64 MiB back).
# Outputs: gpurun_out/${TAG}_*   (copy the ones to be judged into profiles/)
TAG=${1:-r2z}
O=gpurun_out
python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub-metrics > $O/${TAG}_ncu_bench.log 2>&1
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $O/${TAG}_prof_$name -f "$@" > $O/${TAG}_ncu_$name.log 2>&1
  ncu -i $O/${TAG}_prof_$name.ncu-rep --page raw --csv > $O/${TAG}_ncu_${name}_raw.csv 2>/dev/null
  rm -f $O/${TAG}_prof_$name.ncu-rep
}
cap miller6 k_miller6 0 1 python tools/time_pairing.py 16
cap msm_acc k_msm_accumulate 0 1 python tools/time_msm.py 18
cap fold k_fold4_xp 6 2 python tools/time_tipp.py 12 1
cap fexp k_final_exp18 10 1 python tools/time_tipp.py 12 1
cap miller18 k_miller18 10 1 python tools/time_tipp.py 12 1
cap scale k_scale_parts 0 2 python tools/time_tipp.py 12 1
tail -3 $O/${TAG}_pytest_gpu.log
cut -c1-400 $O/${TAG}_bench.json
