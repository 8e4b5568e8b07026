"""The lock-free paths under an interleaved schedule (CPU, emulator chaos mode -- test infrastructure).

The serial emulator runs one thread to its next barrier before the next one starts, so the places where the product's
kernels meet each other half-way never happen on CPU: the chunk table's publish protocol (CAS on the first key half,
plain store of the second, "not published yet" retries, `dbg_chunks.cuh: chunk_probe`), the node table's two-slot
probing with batched CAS (`k_expand`), the retry rings, min / add on a slot another lane is still creating, table
growth while reads arrive in slices.  With ``DBG_EMU_CHAOS=seed:permille`` (tests/emu/cuda_emu.h) a fiber gives way in
front of every atomic / ``__ldcg`` / ``__stcg`` / 128-bit CAS with that probability and the scheduler starts each pass
at a random thread.  Results must not depend on the schedule: every case is compared bit for bit (list order
included) with the literal reference's golden output or with the Python restatement.

A mutation shows the mode finds what it is for: with ``atomicAdd(&s->cnt, 1u)`` in ``chunk_probe`` replaced by a
load + store, the serial emulator still passes 300/300 random cases and chaos mode fails 4 of the first 11.
The whole CPU suite also passes with the variable set (``DBG_EMU_CHAOS=5:400 python -m pytest tests -m "not gpu"``).
"""
import ctypes
import random

import pytest

import util
from oracle import dbg_oracle

LAYOUTS = {
    'plain': {},
    'small_tables': {'DBG_B200_CHUNK_SLOTS': '64', 'DBG_B200_NODE_SLOTS': '128'},           # growth / overflow / rebuild
    'regions16': {'DBG_B200_REGIONS': '1', 'DBG_B200_REGION_SLOTS': '16', 'DBG_B200_VB_BITS': '3'},
    'slices': {'DBG_B200_SLICE_BYTES': '64'},                                                # chunk table grows mid-stream
}


def _yields():
    from emu import build_emu
    lib = ctypes.CDLL(build_emu.build())
    lib.emu_chaos_yields.restype = ctypes.c_ulonglong
    return lib.emu_chaos_yields()


@pytest.fixture
def emu(monkeypatch, emu_library):
    from dbg_b200 import _native
    monkeypatch.setattr(_native, '_lib', emu_library)


@pytest.mark.parametrize('layout', sorted(LAYOUTS))
@pytest.mark.parametrize('seed', [1, 2])
def test_random_cases_do_not_depend_on_the_schedule(emu, monkeypatch, layout, seed):
    for name, value in LAYOUTS[layout].items():
        monkeypatch.setenv(name, value)
    monkeypatch.setenv('DBG_EMU_CHAOS', '%d:500' % seed)
    before = _yields()
    r = random.Random(77 * seed + len(layout))
    n = 0
    while n < 60:
        reads, k, thr, final = util.random_dna_case(r, big=(n % 4 == 3))     # 1 in 4: k = 15 .. 63 (two-word keys above 31)
        min_count = r.choice([1, 1, 2])
        try:
            exp = dbg_oracle.snapshot(reads, k, thr, final, max_paths=200, **({'min_count': min_count} if min_count != 1 else {}))
        except RuntimeError:
            continue                                                         # path enumeration above the cap: not a case
        n += 1
        try:
            got = util.product_snapshot(reads, k, thr, final, min_count=min_count)
        except ZeroDivisionError:
            continue                                                         # threshold 0 paths of the generator
        util.assert_same(got, exp, 'chaos %s seed %d case %d %r' % (layout, seed, n, (reads, k, thr, final, min_count)))
    assert _yields() > before, 'chaos mode did not switch a single fiber: DBG_EMU_CHAOS is not read'


def test_golden_cases_incl_peptides_under_chaos(emu, monkeypatch):
    """Every 6th golden fuzz case of the literal reference (DNA and peptide alphabets: both engines)."""
    monkeypatch.setenv('DBG_EMU_CHAOS', '3:500')
    for i, case in enumerate(util.load_golden('fuzz.json.gz')[::6]):
        got = util.product_snapshot(case['reads'], case['k'], case['threshold'], case['final'])
        util.assert_same(got, case['expect'], 'golden case %d under chaos' % (6 * i))


def test_serial_schedule_when_unset(emu, monkeypatch):
    monkeypatch.delenv('DBG_EMU_CHAOS', raising=False)
    before = _yields()
    util.product_snapshot(['ACGTACGGTCA', 'CGTACGGTCAT', 'ACGTACGGTCA'], 4, 2, False)
    assert _yields() == before
