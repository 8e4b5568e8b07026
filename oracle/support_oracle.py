"""CPU restatement of IV_sortOutputs.py:10-15 and :54  --  TEST INFRASTRUCTURE ONLY (see oracle/dbg_oracle.py
for the rules: only tests/, smoke() and bench.py's cpu_baseline leg may import this).

Pinned against the literal reference function, imported from /root/reference/IV_sortOutputs.py when that tree is
present (tests/test_support.py::test_literal_reference)."""


def find_support_read_score(contig, score_table):
    """IV_sortOutputs.py:10-15, line by line."""
    score = 0
    for read in score_table.keys():
        if read in contig:
            score += score_table[read]
    return score


def sort_outputs(contigs, score_table):
    """IV_sortOutputs.py:54 (stable, descending) on a copy; returns (sorted contigs, scores)."""
    out = list(contigs)
    out.sort(key=lambda x: find_support_read_score(x, score_table=score_table), reverse=True)
    return out, [find_support_read_score(c, score_table) for c in out]
