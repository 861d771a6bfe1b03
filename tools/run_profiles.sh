#!/bin/bash
# Round-2 final evidence (builder run): `gpurun -- bash tools/run_profiles.sh` -> gpurun_out/
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
for wl in resnet56_admm mobilenetv2 densenet40 resnet50_dann; do
  timeout 300 python bench.py --workload $wl --steps 30 --warmup 5 > gpurun_out/r02_bench_cfg_$wl.json 2> gpurun_out/r02_bench_cfg_$wl.err
done
python bench.py --timeline --no-cpu-baseline > gpurun_out/timeline.log 2>&1
python tools/conv_bench.py > gpurun_out/conv_bench.txt 2>&1
ALIGNQ_CUDNN_BENCHMARK=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python tools/launch_shares.py gpurun_out/launches.csv gpurun_out/r02_launches_one_step.csv gpurun_out/r02_step_kernel_shares_final.txt > gpurun_out/shares.log 2>&1
rm -f gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3 -c 9 -o gpurun_out/conv_full python tools/conv_one.py > gpurun_out/ncu_conv.log 2>&1
python tools/ncu_summary.py gpurun_out/conv_full.ncu-rep gpurun_out/r02_ncu_full_conv_kernels.json > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:act_ -c 4 -o gpurun_out/act_full python tools/act_one.py > gpurun_out/ncu_act.log 2>&1
python tools/ncu_summary.py gpurun_out/act_full.ncu-rep gpurun_out/r02_ncu_full_act_kernels.json > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:bnq_ -c 8 -o gpurun_out/bn_full python tools/bn_one.py 128 16 32 32 > gpurun_out/ncu_bn.log 2>&1
python tools/ncu_summary.py gpurun_out/bn_full.ncu-rep gpurun_out/r02_ncu_full_bn_kernels.json > /dev/null 2>&1
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -30
