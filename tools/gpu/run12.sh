L=""
for v in base np2 sort merge merge_nosort all_ilp1 merge_ilp1 base_ilp4; do L="$L --lib py-debruijn_b200/lib/var/lib_$v.so"; done
for c in c2 c3; do
python tools/tune.py $c --steps 10 $L 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if '|' not in l: print(l.strip()); continue
    tag,res=l.split('|',1)
    try:
        d=json.loads(res)
        print('%-28s exp %.3f order %s cins %.3f fin %.3f pack %.3f build %.3f wall %.3f' % (tag.strip(), d['ms_expand'], d.get('ms_region_order'), d['ms_chunk_insert'], d['ms_finalize'], d['ms_pack'], d['ms_build_total'], d['wall_ms']))
    except Exception as e: print(l.strip()[:300])
"
done 2>&1 | tee gpurun_out/r2l_tune.log
