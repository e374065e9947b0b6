out=gpurun_out/r02_last; mkdir -p $out
(time timeout 170 python -m pytest tests/test_gpu_render.py -m gpu -q -x -k "all_parameter or rowgemm or wgrad_building or cfg1_full") > $out/pytest_allparams.log 2>&1
tail -3 $out/pytest_allparams.log
timeout 120 python bench.py --workload train --all-params --no-cpu-baseline > $out/bench_train_allparams.json 2> $out/bench.err
cut -c1-300 $out/bench_train_allparams.json
(time timeout 420 python -m pytest tests -m gpu -q) > $out/pytest_full.log 2>&1
tail -3 $out/pytest_full.log
timeout 200 python bench.py > $out/bench_exact.json 2>> $out/bench.err
cut -c1-300 $out/bench_exact.json
