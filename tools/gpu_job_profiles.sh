#!/bin/bash
# ncu captures of the round-2 kernels (full set, one launch each) + the launch list of the bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N="ncu --set full --clock-control none -f"   # (the 64 MiB return limit: source import only for the headline kernel)
$N --import-source on -k regex:percentile_doy_w5t -s 1 -c 1 -o gpurun_out/prof_r2_w5t python tools/time_pctl.py 721 2 > gpurun_out/prof_r2_w5t.log 2>&1
$N -k regex:bootstrap5 -c 1 -o gpurun_out/prof_r2_boot5c python bench_extra.py --lat 180 --steps 1 --warmup 0 --only bootstrap > gpurun_out/prof_r2_boot5c.log 2>&1
$N -k regex:eqm_train8 -c 1 -o gpurun_out/prof_r2_train16 python tools/time_eqm.py 721 > gpurun_out/prof_r2_train16.log 2>&1
$N -k regex:eqm_adjust -c 1 -o gpurun_out/prof_r2_adjust python tools/time_eqm.py 721 > gpurun_out/prof_r2_adjust.log 2>&1
$N -k regex:period_multi_kernel -c 2 -o gpurun_out/prof_r2_multi2 python tools/batch_once.py 180 > gpurun_out/prof_r2_multi2.log 2>&1
$N -k regex:period_runstat_kernel -s 2 -c 1 -o gpurun_out/prof_r2_cdd python bench.py --steps 2 --warmup 1 --sections none > gpurun_out/prof_r2_cdd.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --sections tx90p,bootstrap,eqm,batch50 > gpurun_out/launches_r2.log 2>&1
# summarise on the box (the return channel is capped at 64 MiB; full reports are ~10 MB each)
python tools/ncu_summarise.py gpurun_out/prof_r2_w5t.ncu-rep gpurun_out/prof_r2_cdd.ncu-rep gpurun_out/prof_r2_boot5c.ncu-rep \
  gpurun_out/prof_r2_train16.ncu-rep gpurun_out/prof_r2_adjust.ncu-rep gpurun_out/prof_r2_multi2.ncu-rep > gpurun_out/ncu_r2_summary.md 2>&1
python tools/ncu_traffic.py gpurun_out/prof_r2_w5t.ncu-rep:10950x721x1440 gpurun_out/prof_r2_cdd.ncu-rep:10950x721x1440 \
  gpurun_out/prof_r2_boot5c.ncu-rep:10950x180x1440 gpurun_out/prof_r2_train16.ncu-rep:10950x721x1440 \
  gpurun_out/prof_r2_adjust.ncu-rep:10950x721x1440 gpurun_out/prof_r2_multi2.ncu-rep:10950x180x1440 > gpurun_out/ncu_traffic.log 2>&1
cp profiles/ncu_traffic.csv gpurun_out/ncu_traffic_r2.csv
rm -f gpurun_out/prof_r2_train16.ncu-rep gpurun_out/prof_r2_adjust.ncu-rep gpurun_out/prof_r2_multi2.ncu-rep
du -sh gpurun_out > gpurun_out/profiles_done.txt
