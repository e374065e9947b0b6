out=gpurun_out/r02_wg2; mkdir -p $out
(time timeout 120 python -m pytest tests/test_gpu_render.py -m gpu -q -x -k "wgrad_building or all_parameter or cfg1_full") > $out/pytest_wgrad.log 2>&1
tail -4 $out/pytest_wgrad.log
timeout 120 python tools/time_wgrad.py > $out/time_wgrad.txt 2>&1; cat $out/time_wgrad.txt | tail -8
timeout 150 python bench.py --workload train --all-params --no-cpu-baseline --steps 6 --warmup 3 > $out/bench_train_allparams.json 2> $out/bench.err
cut -c1-330 $out/bench_train_allparams.json; tail -2 $out/bench.err
