#!/bin/bash
# Round evidence on one B200 (run under gpurun): tests, both bench arms, the ncu launch list of the bench
# command, full ncu captures of the top kernels, the config-5 sweep.  Outputs: gpurun_out/${TAG}_*
TAG=${1:-r1c}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_miller6 -c 1 -o $O/${TAG}_prof_miller6 -f \
    python tools/time_pairing.py 16 > $O/${TAG}_ncu_miller6.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -c 1 -o $O/${TAG}_prof_msm_acc -f \
    python tools/time_msm.py 18 > $O/${TAG}_ncu_msm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fold -s 30 -c 2 -o $O/${TAG}_prof_fold -f \
    python tools/time_tipp.py 12 1 > $O/${TAG}_ncu_fold.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_final_exp6 -s 10 -c 1 -o $O/${TAG}_prof_fexp -f \
    python tools/time_tipp.py 12 1 > $O/${TAG}_ncu_fexp.log 2>&1
# gpurun merges at most 64 MiB back: export the raw pages here and drop the reports
for k in miller6 msm_acc fold fexp; do
  ncu -i $O/${TAG}_prof_$k.ncu-rep --page raw --csv > $O/${TAG}_ncu_${k}_raw.csv 2>/dev/null
  rm -f $O/${TAG}_prof_$k.ncu-rep
done
timeout 400 python tools/sweep.py 20 22 > $O/${TAG}_sweep.jsonl 2> $O/${TAG}_sweep.err
tail -3 $O/${TAG}_pytest.log
cat $O/${TAG}_bench.json | cut -c1-600
