for st in 1 2 3; do
echo "== streams $st"
IPCR_JIT_STREAMS=$st python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stage_ms'], d['products']); print(d['e2e']['ms_per_step'], d['e2e_stage_ms'])"
IPCR_JIT_STREAMS=$st python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --genome-mbp 375 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('375Mbp', d['ms_per_step'], d['stage_ms'], d['products'])"
done
