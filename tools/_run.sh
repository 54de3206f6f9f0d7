O=gpurun_out; T=r2m
python -m pytest tests/test_gpu_msm.py tests/test_gpu_fullsize.py tests/test_gpu_protocols.py tests/test_gpu_verify.py -m gpu -q -x > $O/${T}_pytest.log 2>&1; tail -5 $O/${T}_pytest.log
for lg in 13 18 22; do python tools/time_msm.py $lg 2>&1 | tail -1; done
python tools/time_tipp.py 12 5 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_msm18.csv python tools/time_msm.py 18 > $O/${T}_ncu_msm18.log 2>&1
