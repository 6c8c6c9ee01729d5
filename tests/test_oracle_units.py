"""CPU checks of the oracle's building blocks: RNG restatements, aligner rules, edge cases the reference tests hold."""
import random

from conftest import mutate, random_dna


def test_mt19937_matches_cpython_random():
    from oracle import oracle as O
    for seed in (0, 1, 12345, 2 ** 32 - 1, 2 ** 40 + 7):
        random.seed(seed)
        r = O.Rng(O.RNG_MT, seed)
        assert [random.getrandbits(32) for _ in range(5)] == [r.u32() for _ in range(5)]
        assert random.random() == r.random()
        for n in (1, 2, 3, 4, 5, 7, 1000, 14999, 2 ** 31 - 1):
            assert random.randrange(n) == r.randbelow(n)


def test_philox_known_answers():
    # Random123 known-answer vectors for philox4x32-10
    import ctypes
    import numpy as np
    from oracle import oracle as O
    L = O.lib()

    def philox(ctr, key):
        c = np.asarray(ctr, dtype=np.uint32)
        k = np.asarray(key, dtype=np.uint32)
        out = np.zeros(4, dtype=np.uint32)
        L.bo_philox(c.ctypes.data_as(ctypes.c_void_p), k.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
        return [int(x) for x in out]

    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_banded_aligner_equals_full_matrix_rule():
    from oracle import oracle as O
    rnd = random.Random(5)
    try:
        for limit in (1024 * 1024, 2000, 300):  # force Hirschberg on small inputs too
            O.set_traceback_limit(limit)
            for it in range(150):
                n = rnd.randint(1, 300)
                a = random_dna(rnd, n)
                b = mutate(rnd, a, rnd.choice([0.02, 0.1, 0.3, 0.9])) if rnd.random() < 0.8 else random_dna(rnd, rnd.randint(1, 250), 'ACGTN')
                assert O.align_path(a, b) == O.align_path(a, b, naive=True)
    finally:
        O.set_traceback_limit(1024 * 1024)


def test_cigar_orientation_and_uniques():
    """Unique-answer cases of the reference's own tests: 'I' consumes a query base, 'D' a target base
    (test/test_qscore_model.py:360-388), identity = '=' / columns (test/test_misc.py:234-244)."""
    from oracle import oracle as O
    from badread_b200.misc import compress_cigar, identity_from_edlib_cigar
    ops, d = O.align_path('ACGTACGTAC', 'ACGTCGTAC')   # query has one extra base
    assert d == 1 and ops.count('I') == 1 and ops.count('D') == 0
    ops, d = O.align_path('ACGTCGTAC', 'ACGTACGTAC')
    assert d == 1 and ops.count('D') == 1 and ops.count('I') == 0
    assert compress_cigar('====X==II=D') == '4=1X2=2I1=1D'
    assert identity_from_edlib_cigar('5=5X') == 0.5
    assert identity_from_edlib_cigar('') == 0.0
    assert O.align_path('', 'ACGT') == (None, 4)  # edlib returns no path for empty input


def test_perfect_identity_returns_fragment():
    """test/test_simulate.py:33-50: identity 1.0 => seq == fragment, len(qual) == len(fragment)."""
    from oracle import oracle as O
    from conftest import load_models
    rnd = random.Random(2)
    for names in (('random', 'ideal'), ('nanopore2023', 'nanopore2023')):
        orc = O.Oracle(*load_models(*names))
        for mode in (O.RNG_MT, O.RNG_PHILOX):
            frag = random_dna(rnd, 700)
            seq, qual, ident = orc.sequence_fragment(frag, 1.0, 11, mode=mode)
            assert seq == frag and len(qual) == len(frag) and ident == 1.0


def test_oracle_identity_tracks_target():
    """test/test_simulate.py:53-161 in miniature: achieved error rate within +-50% of the target per read."""
    from oracle import oracle as O
    from conftest import load_models
    rnd = random.Random(3)
    orc = O.Oracle(*load_models('nanopore2023', 'nanopore2023'))
    for target in (0.9, 0.8):
        for i in range(5):
            frag = random_dna(rnd, 3000)
            seq, qual, ident = orc.sequence_fragment(frag, target, 99, read_index=i)
            assert 0.5 * (1 - target) <= 1 - ident <= 1.5 * (1 - target)


def test_philox_batch_equals_single_reads():
    from oracle import oracle as O
    from conftest import load_models
    rnd = random.Random(4)
    orc = O.Oracle(*load_models('random', 'ideal'))
    frags = [random_dna(rnd, n) for n in (5, 300, 1200, 2500)]
    out, total = orc.sequence_batch(frags, [0.9] * 4, 7, [10, 11, 12, 13], n_threads=3)
    for i, f in enumerate(frags):
        s, q, _ = orc.sequence_fragment(f, 0.9, 7, read_index=10 + i)
        assert (out[i][0], out[i][1]) == (s, q)
    assert total == sum(len(o[0]) for o in out)
