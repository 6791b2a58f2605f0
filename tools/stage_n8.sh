# 8-GPU A/B of the STAGE kernel (no tests, short runs): tools/stage_n8.sh <tag>
n=8; tag=${1:-r02bv}
for st in 1 0; do
  GQE_STAGE=$st timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$st bench.py --gpus $n --steps 30 --warmup 3 --no-extras > gpurun_out/${tag}_bench_n${n}_stage$st.json 2> gpurun_out/${tag}_bench_n${n}_stage$st.err
  echo "stage=$st rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${tag}_bench_n${n}_stage$st.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k,v in d.get("sharded",{}).items(): print(k, v["ms_per_step"], v["per_gpu"], v["parity"]["parity_max_abs_err"], v["nvlink"]["achieved_gbs_in"])
PY
done
