"""Differential fuzzing on the CPU emulator (test infrastructure): random small read sets through the drop-in module against
the Python restatement, bit-exact incl. list order, under a table layout / build mode forced by environment variables.

    python tools/fuzz_campaign.py SEED N_CASES plain|regions16|small|slices|passes[+both]

Round 2: 54,000 cases (seeds 1000/2000/3000 x 3000 cases x plain, regions16, small, slices, plain+both, regions16+both; 1 in 5 with
k >= 15, min_count 1..3, final True/False): 0 mismatches."""
import sys, os, random, time, traceback
sys.path[:0] = ['/root/repo', '/root/repo/py-debruijn_b200', '/root/repo/tests']
from emu import build_emu
from dbg_b200 import _native
_native._lib = _native.Library(build_emu.build())
import util
from oracle import dbg_oracle
seed0 = int(sys.argv[1]); n_cases = int(sys.argv[2]); mode = sys.argv[3]
envs = {'plain': {}, 'regions16': {'DBG_B200_REGIONS': '1', 'DBG_B200_REGION_SLOTS': '16', 'DBG_B200_VB_BITS': '3'},
        'small': {'DBG_B200_CHUNK_SLOTS': '64', 'DBG_B200_NODE_SLOTS': '128'}, 'slices': {'DBG_B200_SLICE_BYTES': '64'},
        'passes': {'DBG_B200_L2_REGION_BYTES': '2048', 'DBG_B200_REGIONS': '0'}}
os.environ.update(envs[mode.split('+')[0]])
r = random.Random(seed0)
bad = 0; n = 0; t0 = time.time()
def rc(s): return s.translate(str.maketrans('ACGT', 'TGCA'))[::-1]
while n < n_cases:
    reads, k, thr, final = util.random_dna_case(r, big=(n % 5 == 4))
    mc = r.choice([1, 1, 1, 2, 3])
    both = 'both' in mode and set(''.join(reads)) <= set('ACGT')
    try:
        src = reads + [rc(x) for x in reads] if both else reads
        exp = dbg_oracle.snapshot(src, k, thr, final, max_paths=200, **({'min_count': mc} if mc != 1 else {}))
    except RuntimeError:
        continue
    except TypeError:
        exp = dbg_oracle.snapshot(src, k, thr, final, max_paths=200); mc = 1
    n += 1
    try:
        if both:
            import debruijn as db, io, contextlib
            with contextlib.redirect_stdout(io.StringIO()):
                g, pull, branch, pulled, ect = db.construct_graph(list(reads), k, threshold=thr, final=final, min_count=mc, both_strands=True)
                contigs = db.output_contigs(g, branch, pulled)
            got = {'nodes': [[v, nn.indegree, nn.outdegree] for v, nn in g[0].items()], 'edges': [[v, list(s)] for v, s in g[1].items()],
                   'pull_out_read': list(pull), 'branch': list(branch), 'pulled': list(pulled), 'ect': [[e, c] for e, c in ect.items()], 'contigs': list(contigs)}
        else:
            got = util.product_snapshot(reads, k, thr, final, min_count=mc)
        util.assert_same(got, exp, 'case')
    except ZeroDivisionError:
        pass
    except Exception as e:
        bad += 1
        print('MISMATCH', mode, repr((reads, k, thr, final, mc)), str(e)[:300], flush=True)
        if bad > 3: break
print(mode, 'seed', seed0, 'cases', n, 'bad', bad, 'time', round(time.time() - t0, 1), flush=True)
