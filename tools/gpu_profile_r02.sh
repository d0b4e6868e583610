#!/bin/bash
# r02 profile round: launch list of the train step + ncu --set full of the dominant kernels + the transform kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_launches.csv python tools/train_step_once.py 3 > gpurun_out/r02_launches.log 2>&1
python tools/launch_summary2.py gpurun_out/r02_train_launches.csv 40 > gpurun_out/r02_train_step_launch_summary.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_wgrad_umma" -s 112 -c 56 -f -o gpurun_out/r02_wgrad python tools/train_step_once.py 3 > gpurun_out/r02_ncu_wgrad.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv2d_umma_kernel" -s 206 -c 103 -f -o gpurun_out/r02_conv python tools/train_step_once.py 3 > gpurun_out/r02_ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"instnorm_bwd_fused|adam_flat|pack_weights_tiled" -s 140 -c 20 -f -o gpurun_out/r02_misc python tools/train_step_once.py 3 > gpurun_out/r02_ncu_misc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mdct4_fwd|imdct4_inv" --launch-skip 5 --launch-count 13 -f -o gpurun_out/r02_mdct python tools/mdct_bench.py --flavours mixed --reps 1 --out gpurun_out/r02_mdct_under_ncu.json > gpurun_out/r02_ncu_mdct.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mdct4_fwd|imdct4_inv" --launch-skip 5 --launch-count 13 -f -o gpurun_out/r02_mdct_fp32 python tools/mdct_bench.py --flavours fp32 --reps 1 --out gpurun_out/r02_mdct32_under_ncu.json > gpurun_out/r02_ncu_mdct32.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r02_ncu_conv.log; head -12 gpurun_out/r02_train_step_launch_summary.txt
