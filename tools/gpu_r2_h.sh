#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_kzg.py -x -q --durations=5 > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
grep -v "site-packages" gpurun_out/r2h_pytest.log | tail -40
