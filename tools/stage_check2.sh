for m in 3 2 0; do GQE_FORCE_STAGE=$m timeout -s KILL 120 python tools/ab_kernel.py bio-mix-d256-b65536 2>&1 | tail -1; done
