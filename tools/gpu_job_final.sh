#!/bin/bash
# full GPU test suite + the driver's bench command at N=1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_final_pytest.log
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
echo "bench rc=$?" >> gpurun_out/r2_final_pytest.log
