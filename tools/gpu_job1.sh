#!/bin/bash
# GPU job: new tests + small-lat bench flow check + full bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python -m pytest tests/test_gpu_streaming.py tests/test_gpu_fullsize.py tests/test_zz_gpu_addenda.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_job1_pytest.log
timeout 600 python bench.py --lat 48 --steps 3 --cpu-lat 1 > gpurun_out/r2_bench_small.json 2> gpurun_out/r2_bench_small.err
echo "small rc=$?" >> gpurun_out/r2_job1_pytest.log
if [ -s gpurun_out/r2_bench_small.json ]; then
  timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err
  echo "full rc=$?" >> gpurun_out/r2_job1_pytest.log
fi
