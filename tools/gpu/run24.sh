python -m pytest tests/test_contention.py tests/test_degrees.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2ac_contention.log
python tools/tune.py c2 --steps 20 --env DBG_B200_FAR_GRID=4,8,16,32,4 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if '|' not in l: print(l.strip()); continue
    tag,res=l.split('|',1)
    try:
        d=json.loads(res)
        print('%-28s pack %.3f cins %.3f exp %.3f fin %.3f lists %.3f build %.3f walk %.3f wall %.3f' % (tag.strip(), d['ms_pack'], d['ms_chunk_insert'], d['ms_expand'], d['ms_finalize'], d['ms_lists'], d['ms_build_total'], d['ms_walk_total'], d['wall_ms']))
    except Exception as e: print(l.strip()[:300])
" | tee gpurun_out/r2ac_far.log
