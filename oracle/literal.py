"""Harness around the LITERAL reference (``/root/reference/debruijn.py``)  --  TEST INFRASTRUCTURE ONLY.

Used only in the build container (the reference does not travel to the GPU box) by
``oracle/make_golden.py`` to generate the committed fixtures in ``tests/golden/`` and by
``tests/test_oracle_golden.py::test_literal_live`` (skipped when the reference is absent).

The reference is loaded by path and run unmodified.  Two accommodations, both outside
the reference's code: the recursion limit / thread stack are raised because
``DFS`` (debruijn.py:288-316) recurses once per path node, and stdout is captured
because ``construct_graph`` / ``output_contigs`` print three status lines
(debruijn.py:224,236,336).
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import threading

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the build container's reference tree, or an unmodified copy that travelled with the repo (git-ignored baseline/_ref)
REFERENCE_PATHS = ('/root/reference/debruijn.py', os.path.join(_ROOT, 'baseline', '_ref', 'debruijn.py'))


def available():
    return any(os.path.exists(p) for p in REFERENCE_PATHS)


def load():
    for p in REFERENCE_PATHS:
        if os.path.exists(p):
            spec = importlib.util.spec_from_file_location('_literal_debruijn', p)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    raise FileNotFoundError('literal reference not present: %s' % (REFERENCE_PATHS,))


def _run_big_stack(fn):
    """Run fn() in a thread with a 1 GiB stack and a huge recursion limit."""
    box = {}

    def tgt():
        try:
            box['v'] = fn()
        except BaseException as e:  # noqa: BLE001 - re-raised in caller
            box['e'] = e

    old = sys.getrecursionlimit()
    sys.setrecursionlimit(10_000_000)
    threading.stack_size(1 << 30)
    try:
        t = threading.Thread(target=tgt)
        t.start()
        t.join()
    finally:
        threading.stack_size(0)
        sys.setrecursionlimit(old)
    if 'e' in box:
        raise box['e']
    return box['v']


def snapshot(reads, k, threshold, final, mod=None):
    """Run the literal reference; same JSON-able structure as dbg_oracle.snapshot plus
    the captured stdout lines."""
    mod = mod or load()

    def go():
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            g, pull, branch, pulled, ect = mod.construct_graph(list(reads), k, threshold=threshold,
                                                               final=final)
            V, E = g
            nodes = [[v, n.indegree, n.outdegree] for v, n in V.items()]
            edges = [[v, list(s)] for v, s in E.items()]
            contigs = mod.output_contigs(g, branch, pulled)
        return {
            'nodes': nodes,
            'edges': edges,
            'pull_out_read': list(pull),
            'branch': list(branch),
            'pulled': list(pulled),
            'ect': [[e, c] for e, c in ect.items()],
            'contigs': list(contigs),
            'stdout': buf.getvalue().splitlines(),
        }

    return _run_big_stack(go)


def timed(reads, k, threshold, final, mod=None):
    """construct_graph + output_contigs of the literal reference, timed separately (seconds)."""
    import time
    mod = mod or load()

    def go():
        with contextlib.redirect_stdout(io.StringIO()):
            t0 = time.perf_counter()
            g, pull, branch, pulled, ect = mod.construct_graph(list(reads), k, threshold=threshold, final=final)
            t1 = time.perf_counter()
            contigs = mod.output_contigs(g, branch, pulled)
            t2 = time.perf_counter()
        return t1 - t0, t2 - t1, len(contigs)

    return _run_big_stack(go)
