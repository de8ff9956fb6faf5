#!/bin/bash
# Round measurements on one B200 (run through gpurun from the repo root); everything lands in gpurun_out/final/.
# usage: bash tools/measure_round.sh [part ...]    parts: tests bench configs train ncu_list ncu_full   (default: all but ncu_full)
# Every step runs under its own `timeout`; .ncu-rep files never stay in gpurun_out/ (64 MiB limit).
set -u
O=gpurun_out/final; mkdir -p $O
PARTS="${*:-tests bench configs train ncu_list}"
has() { case " $PARTS " in *" $1 "*) return 0;; esac; return 1; }
if has tests; then
  ( time timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/pytest_gpu.log 2>&1
  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
fi
if has bench; then   # the bench line (default flags) + per-launch table, and the reference arm
  timeout 420 python bench.py --profile-out $O/per_launch_s_bs32.json > $O/bench_s_bs32.json 2> $O/bench_s_bs32.err
  timeout 120 python bench.py --impl reference --steps 5 --warmup 5 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
fi
X="--steps 10 --warmup 3 --skip-cpu-baseline --skip-train --skip-ref-gpu"
if has configs; then   # BASELINE configs 3 (M CamVid bs32) and 4 (L 1024x2048 bs16)
  timeout 180 python bench.py --model pidnet_m --classes 11 --batch 32 --height 720 --width 960 $X > $O/bench_pidnet_m_720x960.json 2>/dev/null
  timeout 180 python bench.py --model pidnet_l --classes 19 --batch 16 --height 1024 --width 2048 $X --profile-out $O/per_launch_l_bs16.json > $O/bench_pidnet_l_1024x2048.json 2>/dev/null
fi
if has configs_extra; then
  timeout 180 python bench.py --model pidnet_m --classes 19 --batch 16 --height 1024 --width 2048 $X > $O/bench_pidnet_m_1024x2048.json 2>/dev/null
  timeout 180 python bench.py --model pidnet_s --classes 11 --batch 32 --height 720 --width 960 $X > $O/bench_pidnet_s_720x960.json 2>/dev/null
fi
if has train; then   # training step with the per-launch table; the criterion alone
  STEPS=20 WARMUP=5 PROFILE=$O/train_per_launch_s_bs12.txt timeout 180 python tools/bench_train.py > $O/train_s_bs12.json 2> $O/train_s_bs12.err
  timeout 120 python tools/bench_criterion.py > $O/criterion_12x1024x1024.json 2> $O/criterion.err
fi
if has ncu_list; then   # ncu launch list of the inference bench command (cold-cache, serialised: shares, not absolutes)
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launch_list_s_bs32.csv \
      python bench.py --steps 2 --warmup 3 --skip-extras --skip-cpu-baseline --skip-train --skip-ref-gpu > $O/ncu_list.log 2>&1
fi
if has ncu_full; then   # `ncu --set full` of selected kernels (NCU_K = kernel-name regex, NCU_C = launches, NCU_CMD = infer | train)
  K="${NCU_K:-conv3_ws_kernel}"; C="${NCU_C:-12}"
  if [ "${NCU_CMD:-infer}" = "train" ]; then CMD="python tools/bench_train.py"; export QUICK=1 STEPS=1 WARMUP=1 GRAPH=0; T=train
  else CMD="python bench.py --steps 1 --warmup 1 --skip-extras --skip-cpu-baseline --skip-train --skip-ref-gpu"; T=infer; fi
  timeout 400 ncu --set full --clock-control none -k regex:"$K" -c $C -o /tmp/ncu_full_$T $CMD > $O/ncu_full_$T.log 2>&1
  timeout 120 ncu -i /tmp/ncu_full_$T.ncu-rep --page raw --csv > $O/ncu_full_${T}_raw.csv 2>/dev/null
  rm -f /tmp/ncu_full_$T.ncu-rep
fi
ls -la $O
