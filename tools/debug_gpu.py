"""Scratch GPU diagnostic: build-only checks of the node table against the C oracle."""
import os, sys, random, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'py-debruijn_b200'), os.path.join(ROOT, 'tests')]
import numpy as np
import util
from dbg_b200 import Graph
from oracle import c_oracle

def check(reads, k, thr, final, name):
    try:
        g = Graph(k)
        g.build(reads, threshold=thr, final=final)
    except Exception as e:
        print(name, 'BUILD FAIL', k, thr, final, len(reads), repr(e)[:300]); return False
    st = g.stats()
    res = c_oracle.run(reads, k, thr, final, strings=False)
    nd = g.nodes()
    keys = (nd['key_hi'].astype(object) << 64) | nd['key_lo'].astype(object) if k > 31 else nd['key_lo']
    ndup = len(keys) - len(set(keys.tolist()))
    ok = st.n_nodes == res.n_nodes and ndup == 0 and np.array_equal(nd['cnt'], res.cnt) and np.array_equal(nd['tag'], res.tag)
    if not ok:
        print(name, 'NODE MISMATCH k=%d nodes %d vs %d dups %d' % (k, st.n_nodes, res.n_nodes, ndup))
    try:
        g.walk(stop_at_branch=not final)
    except Exception as e:
        print(name, 'WALK FAIL', k, thr, final, len(reads), 'nodes', st.n_nodes, 'multi', st.n_multi, 'branch', st.n_branch, repr(e)[:300]); return False
    return ok

r = random.Random(3)
genome = ''.join(r.choice('ACGT') for _ in range(3000))
reads = []
for _ in range(600):
    p = r.randrange(0, 3000 - 80 + 1)
    s = genome[p:p + 80]
    reads.append(''.join(r.choice('ACGT') if r.random() < 0.01 else c for c in s))
reads = list(dict.fromkeys(reads))
for k in (12, 16, 17, 21, 31, 32, 40):
    print('smoke-like k', k, check(reads, k, 2, False, 'smoke'))
r = random.Random(1003)
bad = 0
for i in range(60):
    rd, k, thr, final = util.random_dna_case(r, big=True)
    if not check(rd, k, thr, False, 'big%d' % i):
        bad += 1
print('bad', bad)
