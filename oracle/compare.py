"""Descriptor-level comparison of a built + walked graph handle with a run of the C oracle
(oracle/dbg_oracle.c), used at sizes where strings cannot be materialised:
node table in the reference's dict order (keys, first-occurrence tags, edge counts, indegree
flag, outdegree, kept successors in E order), branch / pulled lists, pull-out read flags and the
contig descriptors (start, n_nodes, terminal, kind) in the reference's order.

TEST INFRASTRUCTURE (like everything under oracle/): imported by tests/, tools/check_*.py and the parity
block of bench.py (the checker, outside every timed region) -- never by the product.
"""
import numpy as np


def compare_with_oracle(g, res, final, report=None):
    """Returns the list of rows that differ (empty = bit-exact).  `report(name, ok)` is called per row."""
    st = g.stats()
    bad = []

    def check(name, cond):
        cond = bool(cond)
        if report:
            report(name, cond)
        if not cond:
            bad.append(name)

    nd = g.nodes()
    check('n_nodes', st.n_nodes == res.n_nodes)
    check('n_windows', st.n_windows == res.n_windows)
    if st.n_nodes != res.n_nodes:
        return bad
    check('node keys (dict order)', np.array_equal(nd['key_lo'], res.key_lo) and np.array_equal(nd['key_hi'], res.key_hi))
    check('first-occurrence tags', np.array_equal(nd['tag'], res.tag))
    check('edge counts', np.array_equal(nd['cnt'], res.cnt))
    indeg = ((nd['flags'] & 2) == 0).astype(np.uint8)
    check('indegree flags', np.array_equal(indeg, res.indeg))
    check('outdegree', np.array_equal(((nd['kept'] >> 12) & 7).astype(np.uint8), res.outdeg))
    nk = ((nd['kept'] >> 8) & 7).astype(np.uint8)
    check('kept successor counts', np.array_equal(nk, res.n_kept))
    kb = np.stack([(nd['kept'] >> (2 * j)) & 3 for j in range(4)], axis=1).astype(np.uint8)
    msk = np.arange(4)[None, :] < nk[:, None]
    check('kept successors (E order)', np.array_equal(kb * msk, res.kept_base * msk))
    slot_to_idx = np.full(int(st.table_slots), -1, np.int64)
    slot_to_idx[nd['id']] = np.arange(len(nd['id']))
    if not final:
        check('branch list', np.array_equal(slot_to_idx[g.branch_ids()], res.branch.astype(np.int64)))
    check('pulled list', np.array_equal(slot_to_idx[g.pulled_ids()], res.pulled.astype(np.int64)))
    if not final:
        check('pull-out read flags', np.array_equal(g.pull_read_flags(), res.read_flags))
    d = g.contig_descriptors()
    check('n_contigs', len(d) == len(res.contigs))
    if len(d) == len(res.contigs):
        check('contig starts', np.array_equal(slot_to_idx[d['start']], res.contigs[:, 0].astype(np.int64)))
        check('contig n_nodes', np.array_equal(d['n_nodes'], res.contigs[:, 1]))
        check('contig terminals', np.array_equal(slot_to_idx[d['terminal']], res.contigs[:, 2].astype(np.int64)))
        check('contig kinds', np.array_equal(d['kind'], res.contigs[:, 3]))
    return bad


def compare_batches_with_oracle(g, desc, bases, off, batch_off, k, thr, picks, c_oracle):
    """Batched build (BASELINE configs[3]) against the C oracle run on single templates `picks`."""
    st = g.stats()
    nd = g.nodes()
    T = len(batch_off) - 1
    nb_lo = np.searchsorted(nd['batch'], np.arange(T + 1))
    lo = np.zeros(len(desc), np.uint64)
    hi = np.zeros(len(desc), np.uint64)
    ids = np.ascontiguousarray(desc['start'], np.uint32)
    if len(ids):
        g.lib.check(g.h, g.lib.dbg_node_keys(g.h, ids.ctypes.data, len(ids), lo.ctypes.data, hi.ctypes.data))
    cb = g.batch_of_keys(lo, hi)
    c_lo = np.searchsorted(cb, np.arange(T + 1))
    slot_to_idx = np.full(int(st.table_slots), -1, np.int64)
    slot_to_idx[nd['id']] = np.arange(len(nd['id']))
    kmask = np.uint64((1 << (2 * k)) - 1) if 2 * k < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    bad = []
    for b in picks:
        r0, r1 = int(batch_off[b]), int(batch_off[b + 1])
        sub_off = off[r0:r1 + 1] - off[r0]
        sub = bases[int(off[r0]):int(off[r1])]
        res = c_oracle.run(None, k, thr, False, strings=False, packed=(sub.tobytes(), sub_off))
        n0, n1 = int(nb_lo[b]), int(nb_lo[b + 1])
        good = (n1 - n0) == res.n_nodes
        if good:
            good &= np.array_equal(nd['key_lo'][n0:n1] & kmask, res.key_lo)
            good &= np.array_equal(nd['tag'][n0:n1] - (np.uint64(r0) << np.uint64(32)), res.tag)
            good &= np.array_equal(nd['cnt'][n0:n1], res.cnt)
            good &= np.array_equal(((nd['flags'][n0:n1] & 2) == 0).astype(np.uint8), res.indeg)
            a, z = int(c_lo[b]), int(c_lo[b + 1])
            good &= (z - a) == len(res.contigs)
            if good and z > a:
                d = desc[a:z]
                good &= np.array_equal(slot_to_idx[d['start']] - n0, res.contigs[:, 0].astype(np.int64))
                good &= np.array_equal(d['n_nodes'], res.contigs[:, 1])
                good &= np.array_equal(slot_to_idx[d['terminal']] - n0, res.contigs[:, 2].astype(np.int64))
                good &= np.array_equal(d['kind'], res.contigs[:, 3])
        if not good:
            bad.append(b)
    return bad
