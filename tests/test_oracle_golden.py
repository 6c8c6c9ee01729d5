"""
The oracle pinned to the reference: oracle/badread_oracle.c in MT mode must reproduce, byte for byte, what the
unmodified reference (run with oracle/edlib_shim) produced for tests/golden/*.json (oracle/make_golden.py).
Also pins the host table builder and the precompiled model tables against the reference's own tables.
"""
import hashlib
import io
import json
import os

import pytest

from conftest import load_models

GOLDEN = os.path.join(os.path.dirname(os.path.realpath(__file__)), 'golden')


def _load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def test_sequence_fragment_matches_reference():
    from oracle import oracle as O
    cases = _load('golden_sequence_fragment.json')
    assert len(cases) >= 60
    oracles = {}
    for c in cases:
        key = (c['error_model'], c['qscore_model'])
        if key not in oracles:
            oracles[key] = O.Oracle(*load_models(*key))
        seq, qual, ident = oracles[key].sequence_fragment(c['fragment'], c['identity'], c['seed'], mode=O.RNG_MT)
        assert seq == c['seq'], key
        assert qual == c['qual'], key
        assert ident == c['actual_identity'], key


def test_get_qscores_matches_reference():
    from oracle import oracle as O
    for c in _load('golden_get_qscores.json'):
        em, qm = load_models('random', c['qscore_model'])
        orc = O.Oracle(em, qm)
        qual, matches, cols = orc.get_qscores(c['seq'], c['frag'], c['seed'], mode=O.RNG_MT)
        assert qual == c['qual']
        assert matches / cols == c['actual_identity']
        # identity_by_qscores is a pure function of the quality string (qscore_model.py:73)
        import statistics
        from badread_b200.qscore_model import qscore_char_to_error_prob
        assert 1.0 - statistics.mean(qscore_char_to_error_prob(q) for q in qual) == c['identity_by_qscores']


def _digest(d1, d2):
    h = hashlib.sha256()
    for key in sorted(d1.keys()):
        h.update(repr((key, d1[key], d2[key])).encode())
    return h.hexdigest()


@pytest.mark.parametrize('name', ['nanopore2018', 'nanopore2020', 'nanopore2023', 'pacbio2016', 'pacbio2021'])
def test_precompiled_tables_match_reference(name):
    from badread_b200.error_model import ErrorModel
    from badread_b200.qscore_model import QScoreModel
    want = _load('golden_tables.json')[name]
    out = io.StringIO()
    em = ErrorModel(name, out)
    assert em.kmer_size == want['kmer_size']
    assert _digest(em.alternatives, em.probabilities) == want['error']
    qm = QScoreModel(name, out)
    assert qm.kmer_size == want['qscore_kmer_size']
    assert _digest(qm.scores, qm.probabilities) == want['qscore']


def test_align_kmers_matches_reference():
    from badread_b200.error_model import align_kmers
    for c in _load('golden_align_kmers.json'):
        assert align_kmers(c['kmer'], c['alt']) == c['slots'], c
