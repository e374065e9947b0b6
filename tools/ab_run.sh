# usage: bash tools/ab_run.sh OUTDIR VARIANT...   ("" = default build); benches + timeline trace per variant
out=gpurun_out/$1; shift; mkdir -p $out
for v in "$@"; do
  sfx=${v:+_$v}
  export NSOS_LIB=nerf-sos_b200/lib/libnerfsos$sfx.so
  python bench.py --no-cpu-baseline > $out/bench_exact$sfx.json 2>> $out/bench.err
  python bench.py --mode fast --no-cpu-baseline > $out/bench_fast$sfx.json 2>> $out/bench.err
  python tools/trace_report.py exact > $out/trace_exact$sfx.txt 2>&1; python tools/trace_report.py fast > $out/trace_fast$sfx.txt 2>&1
  (timeout 200 python -m pytest tests/test_gpu_render.py -m gpu -x -q 2>&1 | tail -2) > $out/pytest$sfx.log
done
tail -3 $out/bench.err; tail -n 2 $out/pytest*.log
for f in $out/bench_*.json; do echo $f; cut -c50-100 $f; done
