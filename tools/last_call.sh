# Final check of the tree: full GPU suite, all-parameter training bench, default bench.
out=gpurun_out/r02_final2; mkdir -p $out
(time timeout 300 python -m pytest tests -m gpu -q) > $out/pytest_full.log 2>&1
tail -3 $out/pytest_full.log | head -1
timeout 100 python bench.py --workload train --all-params --no-cpu-baseline --steps 6 --warmup 3 > $out/bench_train_allparams.json 2> $out/bench.err
cut -c1-300 $out/bench_train_allparams.json
timeout 150 python bench.py > $out/bench_exact.json 2>> $out/bench.err
cut -c1-300 $out/bench_exact.json
