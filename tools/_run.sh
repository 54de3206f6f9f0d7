O=gpurun_out; T=r2r
python -m pytest tests/test_gpu_pairing.py tests/test_gpu_fullsize.py tests/test_gpu_protocols.py -m gpu -q -x > $O/${T}_pytest.log 2>&1; tail -3 $O/${T}_pytest.log
python tools/time_round.py 2>&1 | tail -6
python tools/time_tipp.py 12 6 2>&1 | tail -2
for lg in 14 15 16 18; do python tools/time_pairing.py $lg 2>&1 | tail -1; done
