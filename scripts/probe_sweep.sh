python scripts/perf_probe.py --mbp 3000 --pairs 1 --jit 2 --iters 3 2>&1 | grep -E "iter 2|filter:" | tail -2
for cfg in "IPCR_X=1" "IPCR_PLAN_VERIFY_COST=3000" "IPCR_JIT_GROUP=5" "IPCR_JIT_GROUP=3"; do
echo "== $cfg"
env $cfg python scripts/perf_probe.py --mbp 3000 --pairs 100 --jit 2 --iters 2 2>&1 | grep -E "iter 1|filter:" | tail -2
done
echo "== mm3"
python scripts/perf_probe.py --mbp 3000 --pairs 100 --jit 2 --iters 2 --mm 3 2>&1 | grep -E "iter 1|filter:" | tail -2
