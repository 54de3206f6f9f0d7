O=gpurun_out; T=r2w
python -m pytest tests/test_gpu_protocols.py tests/test_gpu_verify.py tests/test_gpu_fullsize.py tests/test_gpu_sharded.py tests/test_gpu_polycommit.py -m gpu -q -x > $O/${T}_pytest.log 2>&1; tail -3 $O/${T}_pytest.log
python tools/time_tipp.py 12 6 2>&1 | tail -2
