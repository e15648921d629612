#!/bin/bash
# round-2 check: template tests, multiply timing, then the whole GPU suite with per-test timeouts
timeout 300 python -m pytest tests/test_gpu_templates.py -x -q --timeout 60 -p no:cacheprovider > gpurun_out/r2c_tpl.log 2>&1; echo rc_tpl=$?
tail -15 gpurun_out/r2c_tpl.log
timeout 200 python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --mul-paths auto,no_templates > gpurun_out/r2c_mm.log 2>&1; echo rc_mm=$?
cat gpurun_out/r2c_mm.log | tail -5
timeout 600 python -m pytest tests -x -m gpu --timeout 90 -v -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1; echo rc_all=$?
tail -8 gpurun_out/r2c_pytest.log
