mkdir -p gpurun_out/s2i
for v in "" _semloop _regs _regsx32; do
  export NSOS_LIB=nerf-sos_b200/lib/libnerfsos$v.so
  python bench.py --no-cpu-baseline > gpurun_out/s2i/bench_exact$v.json 2>> gpurun_out/s2i/bench.err
  python bench.py --mode fast --no-cpu-baseline > gpurun_out/s2i/bench_fast$v.json 2>> gpurun_out/s2i/bench.err
done
(timeout 300 python -m pytest tests/test_gpu_render.py -m gpu -x -q) > gpurun_out/s2i/pytest_regsx32.log 2>&1
python tools/trace_report.py exact > gpurun_out/s2i/trace_exact_regsx32.txt 2>&1; python tools/trace_report.py fast > gpurun_out/s2i/trace_fast_regsx32.txt 2>&1
tail -3 gpurun_out/s2i/pytest_regsx32.log; tail -3 gpurun_out/s2i/bench.err
for f in gpurun_out/s2i/bench_*.json; do echo $f; cut -c50-140 $f; done
