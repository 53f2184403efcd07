cd "$GRAFT_REPO_ROOT"
ZKSC_CHECK_MAX_N=22 timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
grep -v "^E  \|site-packages" gpurun_out/m_pytest.log | grep "batch\|PARITY\|passed\|failed\|rc=\|Error\|error" | tail -12
