B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras"
run() { name=$1; shift; e=(); a=(); for x in "$@"; do if [[ "$x" == *=* ]]; then e+=("$x"); else a+=("$x"); fi; done; env "${e[@]}" $B "${a[@]}" 2>gpurun_out/r2y_err_$name.log | python -c "
import json,sys
p=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); s=p['stage_ms']
print('$name', 'step %.3f e2e %.3f table %.3f frac %.3f | pack %.3f est %.3f cins %.3f sk %.3f exp %.3f init %.3f fin %.3f lists %.3f walk %.3f passes %d launches %d' % (p['ms_per_step'], p['e2e']['ms_per_step'], p['roofline']['kernel_ms'], p['roofline']['frac'], s['ms_pack'], s['ms_estimate'], s['ms_chunk_insert'], s['ms_expand_sketch'], s['ms_expand'], s['ms_table_init'], s['ms_finalize'], s['ms_lists'], s['ms_walk_total'], p['graph']['insert_passes'], p['gpu_launches_per_step']), 'cold', p['table_sizing']['cold'] and round(p['table_sizing']['cold']['ms_per_step'],3))
"; }
{
run c2 A=1
run c3 A=1 --config c3
run c5s A=1 --config c5s
run c4 A=1 --config c4
} 2>&1 | tee gpurun_out/r2y_runs.log
