set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02ba_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02ba_pytest.log
timeout 420 python bench.py --steps 100 --warmup 5 > gpurun_out/r02ba_bench.json 2> gpurun_out/r02ba_bench.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gqe|k_" -c 400 --csv --log-file gpurun_out/r02ba_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eval-shape --no-train-step > gpurun_out/r02ba_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gqe_fused_tc -s 4 -c 1 -f -o gpurun_out/r02ba_fused_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eval-shape --no-train-step --workload bio-mix-d256-b65536 > gpurun_out/r02ba_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gqe_fused_vec -s 4 -c 1 -f -o gpurun_out/r02ba_vec_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eval-shape --no-train-step --workload synth-10m-transe-minsimple-d256-b65536 > gpurun_out/r02ba_vec_full.log 2>&1
ls -la gpurun_out | grep r02ba
