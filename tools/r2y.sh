#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next_rows.py tests/test_dropin_reference.py -x -q --timeout 200 -p no:cacheprovider 2>&1 | tail -2
python tools/small_probe.py 2>&1 | tail -2
