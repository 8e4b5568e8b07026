set -x
python -m pytest tests -m gpu -x -q -k "not c5_full and not c5s_full" --durations=8 > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/r2d_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2d_smoke.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_c2.json 2> gpurun_out/r2d_bench_c2.err; echo "bench rc=$?"; tail -3 gpurun_out/r2d_bench_c2.err
python - <<'PY'
import sys, time, contextlib, io
sys.path[:0]=['.','py-debruijn_b200']
import bench, debruijn as db
cfg=bench.CONFIGS['c1']; bases,off=bench.make_reads(cfg); L=cfg['L']; raw=bases.tobytes().decode()
reads=[raw[i*L:(i+1)*L] for i in range(len(off)-1)]
import cProfile, pstats
def call():
    with contextlib.redirect_stdout(io.StringIO()):
        g,p,b,pl,e=db.construct_graph(reads,31,threshold=2); return db.output_contigs(g,b,pl)
call(); call()
pr=cProfile.Profile(); pr.enable(); 
for _ in range(5): c=call()
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
PY
