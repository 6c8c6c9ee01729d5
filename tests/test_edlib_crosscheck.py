"""
Auto-activating cross-check against the REAL edlib.  edlib is a third-party dependency of
the reference that is absent from this image, so the aligner is pinned to a restatement of edlib's published rules
(oracle/badread_oracle.c; DESIGN.md section 1 says "parity unpinned at that boundary").  The day an `edlib` wheel is
importable these tests run by themselves and compare the oracle's paths with edlib's on ambiguous inputs: tie-breaks
of the traceback, and Hirschberg-sized pairs beyond edlib's 1 MiB traceback estimate.
"""
import random
import re

import pytest

from conftest import mutate, random_dna

edlib = pytest.importorskip('edlib', reason='edlib is not installed in this image (no network); see DESIGN.md section 1')
if not hasattr(edlib, 'align') or 'edlib_shim' in (getattr(edlib, '__file__', '') or ''):
    pytest.skip('the importable edlib is the oracle shim, not the real library', allow_module_level=True)


def _expand(cigar):
    return ''.join(op * int(n) for n, op in re.findall(r'(\d+)([=XID])', cigar))


def test_oracle_paths_equal_real_edlib():
    from oracle import oracle as O
    rnd = random.Random(5)
    cases = [('AB', 'BA'), ('AAAA', 'AAA'), ('ACGT', 'AGCT'), ('GATTACA', 'GCATGCU')]
    for _ in range(300):
        a = random_dna(rnd, rnd.randint(1, 60), 'AC')            # tiny alphabet: many co-optimal paths
        cases.append((a, mutate(rnd, a, 0.3)))
    for n in (1500, 2500, 6000):                                  # around and beyond the 1 MiB switch
        a = random_dna(rnd, n)
        cases.append((a, mutate(rnd, a, 0.1)))
        cases.append((mutate(rnd, a, 0.05), a))
    for q, t in cases:
        r = edlib.align(q, t, mode='NW', task='path')
        ops, dist = O.align_path(q, t)
        assert dist == r['editDistance']
        assert ops == _expand(r['cigar']), (len(q), len(t))
