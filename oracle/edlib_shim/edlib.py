"""
edlib.py - an `edlib`-compatible shim over the oracle's aligner (oracle/badread_oracle.c), so that the UNMODIFIED
reference under /root/reference can be imported and run in this container, where the real `edlib` wheel is absent.
TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py to generate tests/golden/ and to time the reference).

Only the call shape Badread uses is implemented: edlib.align(query, target, mode='NW', task='path'|'distance').
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.realpath(__file__)), '..', '..'))
from oracle import oracle as _oracle  # noqa: E402


def _compress(ops):
    out, i = [], 0
    while i < len(ops):
        j = i
        while j < len(ops) and ops[j] == ops[i]:
            j += 1
        out.append(f'{j - i}{ops[i]}')
        i = j
    return ''.join(out)


def align(query, target, mode='NW', task='distance', k=-1, additionalEqualities=None):
    if mode != 'NW' or additionalEqualities is not None or k != -1:
        raise NotImplementedError('edlib shim: only mode="NW", k=-1 without additionalEqualities')
    ops, dist = _oracle.align_path(query, target)
    result = {'editDistance': dist, 'alphabetLength': len(set(query) | set(target)),
              'locations': [(0, len(target) - 1)] if task != 'distance' else [(None, len(target) - 1)], 'cigar': None}
    if task == 'path' and ops is not None:
        result['cigar'] = _compress(ops)
    return result
