set -u
O=gpurun_out/final; mkdir -p $O
timeout 300 python -m pytest tests/test_net_gpu.py -q -k "options_agree" 2>&1 | tail -2 > $O/pytest_options.log
ncu --set full --clock-control none -k regex:"conv3_ws_kernel" -c 24 -o $O/ncu_full_conv3_ws python bench.py --steps 1 --warmup 1 --skip-extras --skip-cpu-baseline > $O/ncu_full.log 2>&1
ncu -i $O/ncu_full_conv3_ws.ncu-rep --page raw --csv > $O/ncu_full_conv3_ws_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:"stem2_tc_kernel|conv_tc_kernel" -c 4 -o $O/ncu_full_other python bench.py --steps 1 --warmup 1 --skip-extras --skip-cpu-baseline >> $O/ncu_full.log 2>&1
ncu -i $O/ncu_full_other.ncu-rep --page raw --csv > $O/ncu_full_other_raw.csv 2>/dev/null
rm -f $O/*.ncu-rep
ls -la $O | tail -8; cat $O/pytest_options.log
