import os, sys, random
os.environ['DBG_B200_IGNORE_FAULT'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'py-debruijn_b200'), os.path.join(ROOT, 'tests')]
import numpy as np
from dbg_b200 import Graph, decode_keys
from oracle import c_oracle
r = random.Random(5)
read = ''.join(r.choice('ACGT') for _ in range(100))
for k in (32, 33, 40):
    g = Graph(k); g.build([read], threshold=2, final=False)
    nd = g.nodes(); st = g.stats()
    labs = decode_keys(nd['key_lo'], nd['key_hi'], k)
    exp = [read[i:i+k] for i in range(len(read)-k+1)]
    print('k', k, 'nodes', st.n_nodes, 'expected', len(exp), 'slots', st.table_slots)
    bad = [(i, int(nd['tag'][i]) & 0xffffffff, l) for i, l in enumerate(labs) if l not in exp]
    print(' not-expected labels:', len(bad))
    for i, pos, l in bad[:6]:
        print('  pos', pos, l, '\n  exp  ', exp[pos] if pos < len(exp) else None, hex(int(nd['key_hi'][i])), hex(int(nd['key_lo'][i])))
    missing = [i for i, e in enumerate(exp) if e not in labs]
    print(' missing positions:', missing[:40])
    ids = nd['id']; print(' dup ids', len(ids) - len(set(ids.tolist())), 'tags', sorted((nd['tag'] & 0xffffffff).tolist())[:80])
