fmt() { python -c "
import sys,json
for l in sys.stdin:
    if '|' not in l: print(l.strip()); continue
    tag,res=l.split('|',1)
    try:
        d=json.loads(res)
        print('%-28s pack %.3f cins %.3f exp %.3f order %.3f fin %.3f build %.3f wall %.3f nodes %d' % (tag.strip(), d['ms_pack'], d['ms_chunk_insert'], d['ms_expand'], d.get('ms_region_order',0), d['ms_finalize'], d['ms_build_total'], d['wall_ms'], d['nodes']))
    except Exception as e: print(l.strip()[:300])
"; }
L=""
for v in merge0 merge1 merge1_p2 merge1_c256 merge0 merge1; do L="$L --lib py-debruijn_b200/lib/var/lib_$v.so"; done
{
echo c2; python tools/tune.py c2 --steps 20 $L 2>&1 | fmt
echo c3; python tools/tune.py c3 --steps 10 --lib py-debruijn_b200/lib/var/lib_merge0.so --lib py-debruijn_b200/lib/var/lib_merge1.so 2>&1 | fmt
echo c5s; python tools/tune.py c5s --steps 10 --lib py-debruijn_b200/lib/var/lib_merge0.so --lib py-debruijn_b200/lib/var/lib_merge1.so 2>&1 | fmt
} | tee gpurun_out/r2ab_merge.log
