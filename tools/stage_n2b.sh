# N=2: STAGE kernel with different operand masks
n=2; tag=${1:-r02bw}
GQE_FORCE_STAGE=1 timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grouped or full or golden or mix" 2>&1 | tail -1
for mask in 0x1F 0x19; do
  GQE_STAGE_MASK=$mask timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $n --steps 50 --warmup 3 --no-extras > gpurun_out/${tag}_bench_n${n}_$mask.json 2> gpurun_out/${tag}_bench_n${n}_$mask.err
  echo "mask=$mask rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${tag}_bench_n${n}_$mask.json") if l.startswith("{")][-1])
for k,v in d.get("sharded",{}).items(): print(k, v["ms_per_step"], v["per_gpu"], v["parity"]["parity_max_abs_err"], v["nvlink"]["achieved_gbs_in"])
PY
done
