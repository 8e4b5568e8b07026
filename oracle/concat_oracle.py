"""CPU restatement of III_concatenateOutpus.py:57-142  --  TEST INFRASTRUCTURE ONLY
(see oracle/dbg_oracle.py for the rules: only tests/, smoke() and bench.py's cpu_baseline leg may import this).

The reference keeps this algorithm inline under `if __name__ == '__main__'` and draws `current` with the module
level `random.choice` (III_concatenateOutpus.py:65, unseeded).  The "seeded variant" SURVEY asks for is the SAME
code after `random.seed(s)`: the draws then are a function of s and of the list lengths the algorithm itself
produces.  Here the loop is a function over an explicit `random.Random`; it is pinned against outputs of the
UNMODIFIED script run under `random.seed(s)` (oracle/make_golden_concat.py -> tests/golden/concat.json.gz,
tests/test_concat.py::test_oracle_vs_golden)."""
from oracle.support_oracle import find_support_read_score


def check_sub_sequence(contig, output):
    """III_concatenateOutpus.py:13-17."""
    for o in output:
        if contig in o:
            return False
    return True


def concatenate(contigs, sequences_scores, rng, ks=None, out=print, max_iter=100000):
    """III_concatenateOutpus.py:57-137, line by line.  Mutates and returns `contigs`; `out` receives what the
    script prints (the arguments of each print call joined the way print joins them)."""
    def emit(*a):
        out(' '.join(str(x) for x in a))
    if ks is None:
        ks = [i for i in range(20, 5, -1)]                                           # :57
    for k in ks:
        nonconcatenatable = [item for item in contigs if len(item) <= k]            # :59
        for _ in range(max_iter):                                                    # :60
            concat = False
            concatednatable = [item for item in contigs if item not in nonconcatenatable]   # :62
            if len(concatednatable) == 0:
                break
            current = rng.choice(concatednatable)                                    # :65
            for i in range(len(contigs)):
                hit_ht = False
                hit_th = False
                contig = contigs[i]
                if contig == current or len(contig) <= k:                            # :70
                    continue
                current_head = current[:k]
                contig_tail = contig[len(contig) - k:]
                current_tail = current[len(current) - k:]
                contig_head = contig[:k]
                if current_head == contig_tail:                                      # :76
                    contigs[i] = contig + current[k:]
                    concat = True
                    continue
                if current_tail == contig_head:                                      # :80
                    contigs[i] = current + contig[k:]
                    concat = True
                    continue
                for j in range(k - 1):                                               # :85
                    sh = list(current_head)
                    sh[j], sh[j + 1] = sh[j + 1], sh[j]
                    if ''.join(sh) == contig_tail:
                        result_1 = contig + current[k:]
                        result_2 = contig[:len(contig) - k] + current
                        score_1 = find_support_read_score(result_1, sequences_scores)
                        score_2 = find_support_read_score(result_2, sequences_scores)
                        result = result_1 if score_1 > score_2 else result_2
                        hit_ht = True
                        break
                if hit_ht:                                                           # :101
                    emit(current, ' ', contig)
                    emit(result)
                    contigs[i] = result
                    concat = True
                    continue
                for j in range(k - 1):                                               # :108
                    st = list(current_tail)
                    st[j], st[j + 1] = st[j + 1], st[j]
                    if ''.join(st) == contig_head:
                        result_1 = current + contig[k:]
                        result_2 = current[:len(current) - k] + contig
                        score_1 = find_support_read_score(result_1, sequences_scores)
                        score_2 = find_support_read_score(result_2, sequences_scores)
                        result = result_1 if score_1 > score_2 else result_2
                        hit_th = True
                        break
                if hit_th:                                                           # :125
                    emit(current, ' ', contig)
                    emit(result)
                    contigs[i] = result
                    concat = True
                    continue
            if concat:                                                               # :133
                contigs.remove(current)
            else:
                nonconcatenatable.append(current)
        emit('max length after concat(k = {}): '.format(k), len(max(contigs, key=lambda x: len(x))))   # :137
        emit(len(contigs))
    return contigs


def dedup_substrings(contigs):
    """III_concatenateOutpus.py:139-142."""
    outputs = []
    for contig in contigs:
        if contig not in outputs and check_sub_sequence(contig, outputs):
            outputs.append(contig)
    return outputs


def concatenate_outputs(contigs, sequences_scores, rng, out=print):
    """III_concatenateOutpus.py:53-146: the prints around the loop, the loop, the substring dedup."""
    contigs = list(contigs)
    out('max length before concat:  %d' % len(max(contigs, key=lambda x: len(x))))    # :53
    out(str(len(contigs)))                                                             # :55
    concatenate(contigs, sequences_scores, rng, out=out)
    outputs = dedup_substrings(contigs)
    out(str(len(outputs)))                                                             # :144
    out('max length after concat:  %d' % len(max(outputs, key=lambda x: len(x))))     # :145
    return outputs
