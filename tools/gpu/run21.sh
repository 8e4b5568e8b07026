B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras"
run() { name=$1; shift; e=(); a=(); for x in "$@"; do if [[ "$x" == *=* ]]; then e+=("$x"); else a+=("$x"); fi; done; env "${e[@]}" $B "${a[@]}" 2>gpurun_out/r2z_err_$name.log | python -c "
import json,sys
p=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); s=p['stage_ms']
print('$name', 'step %.3f e2e %.3f table %.3f frac %.3f | pack %.3f cins %.3f exp %.3f fin %.3f lists %.3f walk %.3f' % (p['ms_per_step'], p['e2e']['ms_per_step'], p['roofline']['kernel_ms'], p['roofline']['frac'], s['ms_pack'], s['ms_chunk_insert'], s['ms_expand'], s['ms_finalize'], s['ms_lists'], s['ms_walk_total']), 'cold', p['table_sizing']['cold'] and round(p['table_sizing']['cold']['ms_per_step'],3))
"; }
{
for bgb in 2 1 8 64; do
run c2_bg$bgb DBG_B200_BG_BLOCKS=$bgb
run c3_bg$bgb DBG_B200_BG_BLOCKS=$bgb --config c3
done
run c5s_bg1 DBG_B200_BG_BLOCKS=1 --config c5s
run c5s_bg8 DBG_B200_BG_BLOCKS=8 --config c5s
} 2>&1 | tee gpurun_out/r2z_runs.log
