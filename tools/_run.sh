O=gpurun_out; T=r2h
python -m pytest tests -m gpu -q -x > $O/${T}_pytest.log 2>&1; tail -5 $O/${T}_pytest.log
python tools/time_tipp.py 12 6 > $O/${T}_tipp.log 2>&1; tail -3 $O/${T}_tipp.log
