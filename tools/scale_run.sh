# usage: bash tools/scale_run.sh OUTDIR N   -- eval / image / train benches on N GPUs of one node (torchrun, NCCL)
out=gpurun_out/$1; n=$2; mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n "${@:2}"; }
run 29521 --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_exact_n$n.json 2> $out/err.txt
run 29522 --workload image --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_image_n$n.json 2>> $out/err.txt
run 29523 --workload train --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_train_n$n.json 2>> $out/err.txt
grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version" $out/err.txt | tail -5
for f in $out/bench_*_n$n.json; do echo $f; cut -c1-240 $f; done
