# N-GPU A/B of the STAGE kernel on the node-type-sharded 10 M-node table: tools/stage_n2.sh <gpus> <tag>
n=${1:-2}; tag=${2:-r02bo}
timeout -s KILL 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -2
for st in 1 0; do
  GQE_STAGE=$st timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$st bench.py --gpus $n --steps 50 --warmup 3 --no-extras > gpurun_out/${tag}_bench_n${n}_stage$st.json 2> gpurun_out/${tag}_bench_n${n}_stage$st.err
  echo "stage=$st rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${tag}_bench_n${n}_stage$st.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"])
for k,v in d.get("sharded",{}).items(): print(k, v["ms_per_step"], v["per_gpu"], v["parity"]["parity_max_abs_err"], v["nvlink"]["achieved_gbs_in"])
PY
done
