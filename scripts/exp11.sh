echo "== small"; GMAPDP_TRACE=1 python bench.py --no-cpu-baseline --chain-problems 0 --steps 1 --warmup 1 --small 2>&1 | grep "gmapdp_run_batch" | tail -7
echo "== large"; GMAPDP_TRACE=1 python bench.py --no-cpu-baseline --chain-problems 0 --steps 1 --warmup 1 2>&1 | grep "gmapdp_run_batch" | tail -7
