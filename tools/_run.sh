O=gpurun_out; T=r2v
for kp in 4 6 8; do for lg in 16 18; do echo "KP=$kp"; RIPP_B200_M6_KP=$kp python tools/time_pairing.py $lg 2>&1 | tail -1; done; done
echo model; for lg in 14 15 16 17 18; do python tools/time_pairing.py $lg 2>&1 | tail -1; done
python -m pytest tests/test_gpu_pairing.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -2
python tools/time_tipp.py 12 6 2>&1 | tail -2
