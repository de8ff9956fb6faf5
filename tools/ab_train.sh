#!/bin/bash
# A/B of training-step switches (config 5); run under gpurun.  Output: gpurun_out/ab/*.json
O=gpurun_out/ab; mkdir -p $O
run() { name=$1; shift; env "$@" STEPS=30 WARMUP=5 timeout 300 python tools/bench_train.py > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/$name.json'));print(d['ms_per_step'], d['fwd_plus_criterion_ms'], d['bwd_ms'], d['loss'])")"; tail -2 $O/$name.err | cut -c1-300; }
run batch1 PIDNET_WG_BATCH=1
run batch0 PIDNET_WG_BATCH=0
run batch1b PIDNET_WG_BATCH=1
