python scripts/perf_probe.py --mbp 3000 --pairs 1 --jit 2 --iters 3 2>&1 | grep -E "iter 2|filter:" | tail -2
for cfg in "IPCR_JIT_GROUP=8" "IPCR_JIT_GROUP=4" "IPCR_JIT_GROUP=5" ; do
echo "== $cfg"; env $cfg python scripts/perf_probe.py --mbp 3000 --pairs 100 --jit 2 --iters 2 2>&1 | grep -E "panel|iter 1|filter:" | tail -3
done
