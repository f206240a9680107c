for mb in 3000; do
for cfg in "IPCR_X=0" "IPCR_JIT_STAGES=3" "IPCR_JIT_STAGES=2" "IPCR_JIT_STAGE_ROWS=4 IPCR_JIT_STAGES=4" "IPCR_JIT_WARPS=8 IPCR_JIT_STAGES=3" "IPCR_JIT_WARPS=2" ; do
echo "== $mb Mbp $cfg"; env $cfg python scripts/perf_probe.py --mbp $mb --pairs 1 --jit 2 --iters 3 2>&1 | grep -E "iter 2|filter:" | tail -2
done; done
