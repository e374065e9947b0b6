# Round-end measurement set (1 GPU): full GPU test suite, benches (exact / fast / train / reference arm), launch list, ncu full captures.
out=gpurun_out/$1; mkdir -p $out
(time timeout 600 python -m pytest tests -m gpu -q) > $out/pytest.log 2>&1
python bench.py > $out/bench_exact.json 2> $out/bench.err
python bench.py --mode fast --no-cpu-baseline > $out/bench_fast.json 2>> $out/bench.err
python bench.py --workload train --no-cpu-baseline > $out/bench_train.json 2>> $out/bench.err
python bench.py --workload train --all-params --no-cpu-baseline > $out/bench_train_allparams.json 2>> $out/bench.err
python bench.py --workload image --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_image.json 2>> $out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_exact.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/launches_bench_train.csv python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_launch_train.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render_tc -s 3 -c 1 -f -o $out/ncu_exact python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_exact.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render_tc -s 3 -c 1 -f -o $out/ncu_fast python bench.py --mode fast --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_fast.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sem_wgrad -s 21 -c 1 -f -o $out/ncu_semwgrad python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_semwgrad.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
tail -3 $out/pytest.log; tail -2 $out/smoke.log; tail -3 $out/bench.err; for f in $out/bench_*.json; do echo $f; cut -c1-160 $f; done
