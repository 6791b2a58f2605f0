# single-GPU check of the STAGE kernel: every table treated as remote (GQE_FORCE_STAGE=1), or nothing staged (=2)
GQE_FORCE_STAGE=1 timeout -s KILL 120 python tools/ab_kernel.py bio-mix-d256-b65536 2>&1 | tail -1
GQE_FORCE_STAGE=2 timeout -s KILL 120 python tools/ab_kernel.py bio-mix-d256-b65536 2>&1 | tail -1
timeout -s KILL 100 python tools/ab_kernel.py bio-mix-d256-b65536 2>&1 | tail -1
GQE_FORCE_STAGE=1 timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_nodes.py -m gpu -x -q 2>&1 | tail -2
