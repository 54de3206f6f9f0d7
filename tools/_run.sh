O=gpurun_out; T=r2l
python -m pytest tests/test_gpu_abi_host.py tests/test_gpu_variants.py -m gpu -q > $O/${T}_pytest.log 2>&1; tail -15 $O/${T}_pytest.log
