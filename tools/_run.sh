O=gpurun_out; T=r2n
python -m pytest tests/test_gpu_msm.py tests/test_gpu_pairing.py tests/test_gpu_field.py tests/test_gpu_protocols.py -m gpu -q -x > $O/${T}_pytest.log 2>&1; tail -3 $O/${T}_pytest.log
python tools/time_pairing.py 16 2>&1 | tail -2
python tools/time_tipp.py 12 5 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_bench.json 2> $O/${T}_bench.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'miller',d['roofline']['miller_2^16']['kernel_ms'],d['roofline']['miller_2^16']['frac'],'msm',d['roofline']['msm_2^18']['kernel_ms'],d['roofline']['msm_2^18']['frac'])
print(d['sub_metrics']['strong_scaling_one_instance'])
P
