#!/bin/bash
# parity of the tma family + sweep + one ncu capture of it
LBM_B200_ENGINE=tma timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strips.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/tma_sweep.py 4096 4096 2>&1 | tail -14
export LBM_B200_ENGINE=tma LBM_B200_TMA_VARIANT=${1:-0} LBM_B200_TMA_CTAS=${2:-1}
python tools/quick_perf.py 4096 4096 float64 > gpurun_out/tma_plain64.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lbm_step_tma -s 5 -c 1 -o gpurun_out/tma_f64 -f \
    python tools/quick_perf.py 4096 4096 float64 > gpurun_out/tma_ncu64.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/tma_plain64.log
