"""Differential fuzzing of the generic-alphabet engine on the CPU emulator (test infrastructure): random read sets over
alphabets of 2 .. 255 symbols (latin-1 and beyond), k up to the key width's limit, against the Python restatement.

    [DBG_EMU_CHAOS=seed:permille] python tools/fuzz_alphabets.py SEED N_CASES"""
import random
import sys
import time
sys.path[:0] = ['/root/repo', '/root/repo/py-debruijn_b200', '/root/repo/tests']
from emu import build_emu
from dbg_b200 import _native
_native._lib = _native.Library(build_emu.build())
import util
from oracle import dbg_oracle

r = random.Random(int(sys.argv[1]))
n_cases = int(sys.argv[2])
bad = n = 0
t0 = time.time()
while n < n_cases:
    ns = r.choice([2, 3, 5, 8, 9, 16, 17, 20, 25, 32, 33, 64, 65, 100, 128, 129, 200, 255])
    bits = max(1, (ns - 1).bit_length())
    kmax = min(63, 125 // bits - 1)
    k = r.choice([1, 2, 3, 5, 8, kmax - 1, kmax]) if r.random() < 0.7 else r.randrange(1, kmax + 1)
    k = max(1, min(k, kmax))
    base = r.choice([1, 33, 0x80, 0x3B1, 0x4E00]) if ns < 120 else r.choice([0, 1, 0x4E00])
    if base + ns > 256 and base < 0x100:
        base = 0
    al = [chr(base + c) for c in r.sample(range(ns + (0 if base + ns >= 256 and base < 0x100 else r.randrange(0, 20))), ns)] if base >= 0x100 else [chr(base + c) for c in range(ns)]
    G = r.choice([20, 60, 150, 400])
    genome = ''.join(r.choice(al) for _ in range(G))
    if r.random() < 0.5 and G > 2 * k + 10:
        a = r.randrange(0, G // 2); ln = r.randrange(k, k + 6); b = r.randrange(G // 2, max(G // 2 + 1, G - ln))
        genome = (genome[:b] + genome[a:a + ln] + genome[b + ln:])[:G]
    L = max(1, min(G, r.choice([k, k + 1, k + 3, 12, 20, 35, 70])))
    reads = []
    for _ in range(r.choice([1, 3, 8, 20, 60])):
        ln = max(0, min(G, L + r.choice([0, 0, -1, 2, -3])))
        p = r.randrange(0, G - ln + 1)
        s = list(genome[p:p + ln])
        for j in range(len(s)):
            if r.random() < r.choice([0.0, 0.03, 0.08]):
                s[j] = r.choice(al)
        reads.append(''.join(s))
    if r.random() < 0.5:
        reads = list(dict.fromkeys(reads))
    thr, final = r.choice([1, 2, 2, 3, 5, 1.5, 0.5]), r.random() < 0.4
    try:
        exp = dbg_oracle.snapshot(reads, k, thr, final, max_paths=300)
    except RuntimeError:
        continue
    n += 1
    try:
        got = util.product_snapshot(reads, k, thr, final)
        util.assert_same(got, exp, 'case')
    except ZeroDivisionError:
        pass
    except Exception as e:
        if 'distinct characters' in str(e) or 'out of range' in str(e):      # the used characters fit fewer bits / more than 255
            continue
        bad += 1
        print('MISMATCH', repr((reads, k, thr, final)), str(e)[:300], flush=True)
        if bad > 3:
            break
print('alphabets seed', sys.argv[1], 'cases', n, 'bad', bad, 'time', round(time.time() - t0, 1), flush=True)
