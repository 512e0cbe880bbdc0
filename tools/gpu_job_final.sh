#!/bin/bash
# what the driver runs at round end: GPU tests, smoke(), the reference arm, the bench at N=1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/r2_final_pytest.log 2>&1
timeout 1700 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
echo "bench rc=$?" >> gpurun_out/r2_final_pytest.log
