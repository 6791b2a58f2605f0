# Run on the GPU box (under gpurun): parity tests, the default bench line (both arms), the ncu launch list of
# the same command and full captures of the two dominant kernels.   tools/gpu_profile_r02.sh <tag>
set -u
tag=${1:-r02x}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout -s KILL 500 python bench.py --steps 100 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gqe|k_" -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eval-shape --no-train-step > gpurun_out/${tag}_launches.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gqe_fused_tc -s 4 -c 1 -f -o gpurun_out/${tag}_fused_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eval-shape --no-train-step --no-extras --workload bio-mix-d256-b65536 > gpurun_out/${tag}_full.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gqe_fused_vec -s 4 -c 1 -f -o gpurun_out/${tag}_vec_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eval-shape --no-train-step --no-extras --workload synth-10m-transe-minsimple-d256-b65536 > gpurun_out/${tag}_vec_full.log 2>&1
ls -la gpurun_out | grep ${tag}
