L=""
for v in all1 novh nomod nomemset none all1 none; do L="$L --lib py-debruijn_b200/lib/var/lib_$v.so"; done
python tools/tune.py c2 --steps 20 $L 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if '|' not in l: print(l.strip()); continue
    tag,res=l.split('|',1)
    try:
        d=json.loads(res)
        print('%-22s pack %.3f cins %.3f exp %.3f order %.3f fin %.3f build %.3f wall %.3f' % (tag.strip(), d['ms_pack'], d['ms_chunk_insert'], d['ms_expand'], d.get('ms_region_order',0), d['ms_finalize'], d['ms_build_total'], d['wall_ms']))
    except Exception as e: print(l.strip()[:300])
" | tee gpurun_out/r2x_tune.log
