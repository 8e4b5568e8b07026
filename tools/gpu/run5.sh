B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras"
run() { name=$1; shift; env "$@" $B 2>/dev/null | python -c "
import json,sys
p=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); s=p['stage_ms']
print('$name', 'step %.3f table %.3f frac %.3f | pack %.3f est %.3f cins %.3f sk %.3f exp %.3f init %.3f fin %.3f walk %.3f passes %d cslots %d' % (p['ms_per_step'], p['roofline']['kernel_ms'], p['roofline']['frac'], s['ms_pack'], s['ms_estimate'], s['ms_chunk_insert'], s['ms_expand_sketch'], s['ms_expand'], s['ms_table_init'], s['ms_finalize'], s['ms_walk_total'], p['graph']['insert_passes'], p['graph']['chunk_slots_rank0']))
"; }
run base A=1
run cslots2.2M DBG_B200_CHUNK_SLOTS=2200000
run cslots4M DBG_B200_CHUNK_SLOTS=4000000
run load35 DBG_B200_TABLE_LOAD=35
run load40_cs DBG_B200_TABLE_LOAD=40 DBG_B200_CHUNK_SLOTS=2200000
run onepass DBG_B200_L2_REGION_BYTES=4000000000
run threepass DBG_B200_L2_REGION_BYTES=110000000
