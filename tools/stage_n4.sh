n=4; tag=${1:-r02cc}
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 30 --warmup 3 --no-extras > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
echo "rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${tag}_bench_n${n}.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k,v in d.get("sharded",{}).items(): print(k, v["ms_per_step"], v["per_gpu"], v["parity"]["parity_max_abs_err"], v["nvlink"]["achieved_gbs_in"])
PY
