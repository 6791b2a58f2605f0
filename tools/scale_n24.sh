mkdir -p gpurun_out
for n in 2 4; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 50 --warmup 3 --no-extras > gpurun_out/r02ap_bench_n$n.json 2> gpurun_out/r02ap_bench_n$n.err; echo "bench n=$n rc=$?"
done
