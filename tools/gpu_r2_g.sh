#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -X faulthandler -m pytest tests/test_gpu_multi.py -x -q -s -k single_process > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
grep -v "^E  \|site-packages" gpurun_out/r2g_pytest.log | tail -40
