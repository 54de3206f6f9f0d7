O=gpurun_out; T=r2o
for cfg in "4 1184" "1 1184" "1 1776" "1 2368" "1 3552"; do set -- $cfg
echo "== L18_KP=$1 L18_WARPS=$2"
RIPP_B200_L18_KP=$1 RIPP_B200_L18_WARPS=$2 python tools/time_round.py 2>&1 | sed -n 6,13p
RIPP_B200_L18_KP=$1 RIPP_B200_L18_WARPS=$2 python tools/time_tipp.py 12 5 2>&1 | tail -2 | head -1
done
