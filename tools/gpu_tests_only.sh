#!/bin/bash
# one B200: the -m gpu suite, smoke(), and the c4 line
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_only_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_only_pytest.log
grep -v "site-packages" gpurun_out/tests_only_pytest.log | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/tests_only_c4.json 2>/dev/null; python -c "
import json; d=json.loads([l for l in open('gpurun_out/tests_only_c4.json') if l.startswith('{')][-1]); print('c4 ms/step %.4f value %.3e' % (d['ms_per_step'], d['value']))"
