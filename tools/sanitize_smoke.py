"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck): every kernel of the chunk build, the
graph stages with branches and tips, both key widths, the batched mode and the streamed host upload."""
import contextlib
import io
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'py-debruijn_b200')]
os.environ.setdefault('DBG_B200_SLICE_BYTES', '16384')       # several slices even for a small input
import debruijn as db                                           # noqa: E402
from oracle import dbg_oracle                                   # noqa: E402

r = random.Random(3)
genome = ''.join(r.choice('ACGT') for _ in range(6000))
reads = []
for _ in range(1500):
    p = r.randrange(0, 6000 - 80 + 1)
    s = genome[p:p + 80]
    reads.append(''.join(r.choice('ACGT') if r.random() < 0.01 else c for c in s))
reads = list(dict.fromkeys(reads))
for k, final in ((21, False), (21, True), (40, False)):
    with contextlib.redirect_stdout(io.StringIO()):
        g, pull, branch, pulled, ect = db.construct_graph(reads, k, threshold=2, final=final)
        contigs = db.output_contigs(g, branch, pulled)
    exp = dbg_oracle.snapshot(reads, k, 2, final)
    assert contigs == exp['contigs'] and list(branch) == exp['branch'] and list(pulled) == exp['pulled']
batches = [reads[i::4] for i in range(4)]
with contextlib.redirect_stdout(io.StringIO()):
    res = db.construct_graph_batch(batches, 25, threshold=2)
    for b, (g, pull, branch, pulled, ect) in enumerate(res):
        assert db.output_contigs(g, branch, pulled) == dbg_oracle.snapshot(batches[b], 25, 2, False)['contigs']
print('sanitize smoke ok')
