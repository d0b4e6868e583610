#!/bin/bash
# Final r02 GPU round: parity tests, smoke, both bench arms, recipe table, launch lists and ncu --set full summaries of the final kernels.
tag=${1:-r02z}
O=gpurun_out
mkdir -p $O /tmp/ncu
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_ddp_gpu.py > $O/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${tag}_pytest.log
tail -3 $O/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py --steps 30 --warmup 5 > $O/${tag}_bench.json 2> $O/${tag}_bench.err; echo "bench rc=$?" >> $O/${tag}_bench.err
tail -2 $O/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${tag}_bench_reference.json 2> $O/${tag}_bench_reference.err; echo "ref rc=$?"; head -c 600 $O/${tag}_bench_reference.json; echo
timeout 600 python tools/recipe_profile.py 20 > $O/${tag}_recipe_profile.txt 2>&1; head -3 $O/${tag}_recipe_profile.txt
MDCTGAN_CONV_ENGINE=tf32 timeout 600 python tools/recipe_profile.py 20 > $O/${tag}_recipe_profile_tf32.txt 2>&1; head -2 $O/${tag}_recipe_profile_tf32.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/ncu/bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-strong --no-longform --no-extras --cpu-seconds 0.5 > $O/${tag}_bench_under_ncu.log 2>&1
python tools/ncu_launch_list.py /tmp/ncu/bench_launches.csv $O/${tag}_bench_launches.csv.gz "python bench.py --steps 2 --warmup 3 --no-strong --no-longform --no-extras --cpu-seconds 0.5" > $O/${tag}_bench_launch_summary.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/ncu/train_launches.csv python tools/train_step_once.py 3 > $O/${tag}_launches.log 2>&1
python tools/ncu_launch_list.py /tmp/ncu/train_launches.csv $O/${tag}_train_step_launches.csv.gz "python tools/train_step_once.py 3 (last step)" last_step > $O/${tag}_train_step_launch_summary.txt 2>&1
head -16 $O/${tag}_train_step_launch_summary.txt
full() {   # name regex skip count cmd...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o /tmp/ncu/$name "$@" > $O/${tag}_ncu_${name}.log 2>&1
  python tools/ncu_report.py /tmp/ncu/$name.ncu-rep > $O/${tag}_ncu_${name}.txt 2>&1
  tail -1 $O/${tag}_ncu_${name}.log
}
full conv "conv2d_umma_kernel" 206 28 python tools/train_step_once.py 3
full wgrad "conv_wgrad_umma" 92 12 python tools/train_step_once.py 3
full misc "instnorm_bwd_fused|adam_flat|pack_weights_tiled|cout1|multi_loss|conv_wgrad_small|attention" 60 24 python tools/train_step_once.py 3
full mdct "mdct4_fwd|imdct4_inv" 5 8 python tools/mdct_bench.py --flavours mixed --reps 1 --out /tmp/ncu/mdct_under_ncu.json
du -sh $O; ls $O | wc -l
