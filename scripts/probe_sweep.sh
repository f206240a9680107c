python scripts/perf_probe.py --mbp 3000 --pairs 1 --jit 2 --iters 3 2>&1 | grep -E "iter 2|filter:" | tail -2
python scripts/perf_probe.py --mbp 500 --pairs 1 --jit 2 --iters 3 2>&1 | grep -E "iter 2|filter:" | tail -2
for sl in 0 4 6 8; do
echo "== slabs $sl"
IPCR_STREAM_SLABS=$sl python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], d['stage_ms']); print(d['e2e']['ms_per_step'], d['e2e_stage_ms'])"
done
