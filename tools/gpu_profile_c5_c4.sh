cd "$GRAFT_REPO_ROOT"
ZKSC_PROFILE=1 timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/c5_profile.err > gpurun_out/c5_profile.json
grep "zksc profile" gpurun_out/c5_profile.err | tail -22
ZKSC_PROFILE=1 timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/c4_profile.err > gpurun_out/c4_profile.json
grep "zksc profile" gpurun_out/c4_profile.err | tail -45
