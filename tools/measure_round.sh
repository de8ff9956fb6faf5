#!/bin/bash
# Round measurements on one B200 (run through gpurun from the repo root); everything lands in gpurun_out/final/.
# usage: bash tools/measure_round.sh
set -u
O=gpurun_out/final; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
# the bench line (default flags) + per-launch table, and the reference arm
python bench.py --profile-out $O/per_launch_s_bs32.json > $O/bench_s_bs32.json 2> $O/bench_s_bs32.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
# other BASELINE configs
python bench.py --model pidnet_m --classes 11 --batch 32 --height 720 --width 960 --steps 10 --warmup 3 --skip-cpu-baseline > $O/bench_pidnet_m_720.json 2>/dev/null
python bench.py --model pidnet_m --batch 16 --steps 10 --warmup 3 --skip-cpu-baseline > $O/bench_pidnet_m_1024.json 2>/dev/null
python bench.py --model pidnet_l --batch 16 --steps 10 --warmup 3 --skip-cpu-baseline --profile-out $O/per_launch_l_bs16.json > $O/bench_pidnet_l_1024.json 2>/dev/null
python bench.py --model pidnet_s --classes 11 --batch 32 --height 720 --width 960 --steps 10 --warmup 3 --skip-cpu-baseline > $O/bench_pidnet_s_720.json 2>/dev/null
# training step
STEPS=10 WARMUP=3 PROFILE=$O/train_per_launch_s_bs12.txt python tools/bench_train.py > $O/train_s_bs12.json 2> $O/train_s_bs12.err
# ncu: launch list of the bench command, then one full capture of the weight-stationary conv kernels of the first forward
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launch_list_s_bs32.csv \
    python bench.py --steps 2 --warmup 3 --skip-extras --skip-cpu-baseline > $O/ncu_list.log 2>&1
if [ "${NCU_FULL:-1}" = "1" ]; then
  # (raw pages are exported on the box and the .ncu-rep files deleted: gpurun_out/ may not exceed 64 MiB)
  ncu --set full --clock-control none -k regex:"conv3_ws_kernel" -c 24 -o $O/ncu_full_conv3_ws \
      python bench.py --steps 1 --warmup 1 --skip-extras --skip-cpu-baseline > $O/ncu_full.log 2>&1
  ncu -i $O/ncu_full_conv3_ws.ncu-rep --page raw --csv > $O/ncu_full_conv3_ws_raw.csv 2>/dev/null
  ncu --set full --clock-control none -k regex:"stem2_tc_kernel|conv_tc_kernel" -c 4 -o $O/ncu_full_other \
      python bench.py --steps 1 --warmup 1 --skip-extras --skip-cpu-baseline >> $O/ncu_full.log 2>&1
  ncu -i $O/ncu_full_other.ncu-rep --page raw --csv > $O/ncu_full_other_raw.csv 2>/dev/null
  rm -f $O/*.ncu-rep
fi
ls -la $O
