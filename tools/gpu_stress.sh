cd "$GRAFT_REPO_ROOT"
timeout 300 python tools/stress.py --seconds 60 2>&1 | tail -3
timeout 300 python tools/stress.py --seconds 40 --devices 0,1 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/stress.py --seconds 40 2>&1 | tail -3
