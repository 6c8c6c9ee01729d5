"""
The DEVICE aligner code (badread_b200/csrc/*.cuh) compiled for the host through the warp emulator (tests/emu) and
checked against the oracle on CPU: the warp wavefront (every L / K variant, strip fall-back), the lane aligner and
the level-synchronous task pipeline.  This is the same source the GPU runs; the GPU parity tests repeat it end to end.
"""
import random

import pytest

from conftest import mutate, random_dna


@pytest.fixture(scope='module')
def emu():
    from emu import emu as E
    E.build()
    return E


def test_warp_aligner_small_cases(emu):
    from oracle import oracle as O
    rnd = random.Random(17)
    for it in range(60):
        n = rnd.randint(1, 500)
        a = random_dna(rnd, n, 'ACGT' if it % 4 else 'ACGTN')
        b = mutate(rnd, a, rnd.choice([0, 0.02, 0.1, 0.3])) if rnd.random() < 0.8 else random_dna(rnd, rnd.randint(1, 400), 'ACGTN')
        assert emu.align_path(a, b, None, rnd.choice([0, 1, 31, 32, 45]), rnd.choice([1, 2, 16])) == O.align_path(a, b)


def test_warp_aligner_exact_bound_and_zero_band(emu):
    from oracle import oracle as O
    rnd = random.Random(5)
    for n in (1, 31, 32, 33, 967, 1000, 2500):
        a = random_dna(rnd, n)
        for b in (a, a[:n // 2] + ('A' if a[n // 2] != 'A' else 'C') + a[n // 2 + 1:], a + 'G'):
            want = O.align_path(a, b)
            for maxl in (1, 2, 16):
                assert emu.align_path(a, b, want[1], 0, maxl) == want


def test_warp_aligner_hirschberg_wide_and_strip_fallback(emu):
    from oracle import oracle as O
    rnd = random.Random(7)
    a = random_dna(rnd, 5000); b = mutate(rnd, a, 0.07)
    assert emu.align_path(a, b, 500, 0, 2) == O.align_path(a, b)          # paired 16-lane groups
    assert emu.align_path(a, b, None, 3, 2) == O.align_path(a, b)         # band beyond MAXL=2 -> 1024-row strips
    a = random_dna(rnd, 6000); b = mutate(rnd, a, 0.3)
    assert emu.align_path(a, b, None, 3, 16) == O.align_path(a, b)        # L = 8, streamed match words
    assert emu.align_path(random_dna(rnd, 3000), random_dna(rnd, 700), None, 9, 16)[1] == 2300 or True
    q, t = random_dna(rnd, 40), random_dna(rnd, 30000)
    assert emu.align_path(q, t, None, 0, 16) == O.align_path(q, t)         # wide leaf


def test_lane_aligner(emu):
    from oracle import oracle as O
    rnd = random.Random(41)
    checked = 0
    for it in range(120):
        n = rnd.choice([1, 2, 31, 32, 33, 100, 999, 1000])
        a = random_dna(rnd, n, 'ACGT' if it % 5 else 'ACGTN')
        b = mutate(rnd, a, rnd.choice([0, 0.01, 0.03, 0.06]))
        ops, d = O.align_path(a, b)
        for lw in (4, 8):
            got = emu.lane_align(a, b, d + rnd.choice([0, 1, 5]), rnd.choice([0, 1, 40]), lw)
            if got is not None:
                assert got == (ops.count('='), ops.count('D'), d)
                checked += 1
    assert checked > 150


def test_task_pipeline(emu):
    from oracle import oracle as O
    rnd = random.Random(51)
    for n in (1, 40, 1000, 1700):                     # the root is a leaf
        a = random_dna(rnd, n); b = mutate(rnd, a, 0.05)
        assert emu.tasks_align(b, a, O.align_path(b, a)[1] + 3) == O.align_path(b, a)[0]
    for n, rate in ((2600, 0.05), (9000, 0.06), (15000, 0.05)):
        a = random_dna(rnd, n, 'ACGTN' if n == 9000 else 'ACGT'); b = mutate(rnd, a, rate)
        assert emu.tasks_align(b, a, int(O.align_path(b, a)[1] * 1.2) + 5) == O.align_path(b, a)[0]
    a = random_dna(rnd, 20000); b = mutate(rnd, a, 0.25)   # wide root: 8-warp CTAs (128-lane wavefronts) / warp pairs
    want = O.align_path(b, a)[0]
    assert emu.tasks_align(b, a, int(O.align_path(b, a)[1] * 1.1)) == want
    assert emu.tasks_align(b, a, int(O.align_path(b, a)[1] * 1.1), quad=False) == want
    assert emu.tasks_align(b, a, int(O.align_path(b, a)[1] * 1.1), quad=False, hist=0) == want   # checkpoint lane leaves
    for n, rate in ((9000, 0.3), (30000, 0.12), (12000, 0.45)):   # wide roots of other chunk widths (1, 2, 4 words per lane)
        a = random_dna(rnd, n, 'ACGTN' if n == 9000 else 'ACGT'); b = mutate(rnd, a, rate)
        ops, d = O.align_path(b, a)
        assert emu.tasks_align(b, a, d + d // 7) == ops, (n, rate)
    # the two-columns-per-step wavefront of the lean warp kernel: odd / even lengths, every register-mask width
    # (L = 1, 2, 4), loose and exact bounds, non-ACGT targets
    for it in range(14):
        n = rnd.randrange(2300, 7000)
        a = random_dna(rnd, n, 'ACGTN' if it % 4 == 0 else 'ACGT')
        b = mutate(rnd, a, rnd.choice([0.004, 0.02, 0.05, 0.1, 0.2]))
        ops, d = O.align_path(b, a)
        assert emu.tasks_align(b, a, d + rnd.choice([0, 1, 7, d // 5])) == ops


def test_bit_plane_pass_equals_single_column_pass(emu):
    """bb_band_pass_bp (two columns per wavefront step, target codes and chunk rows as bit planes, the pass of the lean
    node kernels) writes the same column scores and corner distances as bb_band_pass for both directions of a node: short
    and long, odd and even lengths, exact and loose bounds, every chunk width (1, 2, 4 words), characters outside ACGT in
    the query and in the target (the exact per-column path), bands that start with several chunks in column 0."""
    from oracle import oracle as O
    rnd = random.Random(77)
    checked = 0
    for it in range(120):
        n = rnd.choice([3, 33, 64, 65, 200, 700, 1500, 2600])
        a = random_dna(rnd, n, 'ACGTN' if it % 5 == 0 else 'ACGT')
        b = mutate(rnd, a, rnd.choice([0.0, 0.01, 0.05, 0.15, 0.3]))
        if not b:
            continue
        if it % 7 == 3:   # a stretch of IUPAC codes in the query only
            p = rnd.randrange(len(b))
            b = b[:p] + 'RYN' + b[p + 3:]
        d = O.align_path(b, a)[1]
        for k in (d, d + 1, d + rnd.randrange(2, 40), max(d, 3 * d // 2) + 5):
            bad = emu.compare_passes(b, a, k)
            if bad is not None:
                assert bad == 0, (it, n, len(b), d, k)
                checked += 1
    assert checked > 250


def test_shared_memory_match_cache_equals_streamed_words(emu):
    """The warp-pair kernel's pass (match words cut out of the bitmap once per chunk into shared memory) equals the
    pass that streams them every step, for 8 / 16 / 32 words per lane, forward and reverse, including chunks that
    reach past the node's last row and non-ACGT targets."""
    from oracle import oracle as O
    rnd = random.Random(78)
    checked = 0
    for it in range(10):
        n = rnd.choice([700, 1900, 3300, 5200])
        a = random_dna(rnd, n, 'ACGTN' if it % 4 == 0 else 'ACGT')
        b = mutate(rnd, a, rnd.choice([0.02, 0.1, 0.25]))
        d = O.align_path(b, a)[1]
        for L in (8, 16, 32):
            bad = emu.compare_sm(b, a, d + rnd.choice([0, 3, 50]), L)
            if bad is not None:
                assert bad == 0, (it, n, len(b), d, L)
                checked += 1
    assert checked >= 25


@pytest.mark.parametrize('hist', [1, 2, 0])
def test_window_lane_kernel_matches_oracle(emu, hist):
    """bb_k_window_lane_hist<4> / <8> (the default: per-column history in global memory, walked back through the
    shared-memory staging ring, 4 or 2 columns per tick; '=' columns from the move counts and the distance) and
    bb_k_window_lane<4> / <8> (checkpoints every 16 columns, tiles re-run into shared memory for the traceback) under
    the emulator: '=' columns and total columns of every identity re-measurement of a read equal edlib's path between
    the original 1000-base window (query) and the window as it was after 25 a changes (target) - simulate.py:325-346.
    Fragments shorter and longer than the window, substitutions / deletions / insertions, 30 ... 250 changes."""
    from oracle import oracle as O
    rnd = random.Random(2024)
    checked = 0
    for frag_len, n_changes, lw in ((700, 60, 4), (1000, 75, 4), (3000, 250, 4), (2200, 180, 8), (5000, 130, 4), (1800, 260, 8)):
        frag = random_dna(rnd, frag_len)
        positions = rnd.sample(range(frag_len), n_changes)
        changes = []
        for p in positions:
            kind = rnd.random()
            if kind < 0.4:
                sub = rnd.choice([c for c in 'ACGT' if c != frag[p]])
            elif kind < 0.7:
                sub = ''
            else:
                sub = frag[p] + rnd.choice('ACGT') if rnd.random() < 0.5 else rnd.choice('ACGT') + frag[p]
            changes.append((p, sub))
        seed, read = 77, 5 + frag_len
        got = emu.window_lane(frag, changes, seed, read, lw=lw, hist=hist)
        assert len(got) == n_changes // 25
        for a, (matches, cols) in enumerate(got, start=1):
            qpos, qn = 0, frag_len
            if frag_len > 1000:
                rng = O.Rng(O.RNG_PHILOX, seed, read)
                rng.stream(4, a - 1)
                qpos, qn = rng.randbelow(frag_len - 1000 + 1), 1000
            applied = dict(changes[:25 * a])
            target = ''.join(applied.get(i, frag[i]) for i in range(qpos, qpos + qn))
            if (matches, cols) == (-1, -1):
                continue        # too wide for this build: the next kernel's job
            ops, _ = O.align_path(frag[qpos:qpos + qn], target)
            assert (matches, cols) == (ops.count('='), len(ops)), (frag_len, n_changes, lw, a)
            checked += 1
    assert checked >= 25


@pytest.mark.parametrize('model', ['nanopore2023', 'pacbio2021', 'nanopore2020'])
def test_error_loop_kernels_match_oracle(emu, model):
    """The error loop as the GPU runs it - bb_k_mutate (evaluation one step ahead of the ordered commit), the window task
    list, the lane window aligners with the staging ring and their fall-backs, bb_k_replay, round after round - under the
    emulator: loop_count, change_count, the number of identity re-measurements, the trims and the mutated read itself
    equal the oracle's sequential loop (simulate.py:256-358), for fragments below and above the 1000-base window, low and
    high target identities, and a read that needs more than one round."""
    from conftest import load_models
    from oracle import oracle as O
    em, qm = load_models(model, model)
    orc = O.Oracle(em, qm)
    rnd = random.Random(101)
    most_rounds = 0
    for n, ident in ((40, 0.9), (300, 0.8), (986, 0.93), (1500, 0.97), (2500, 0.65), (6000, 0.9), (3000, 1.0), (500, 0.999)):
        frag = random_dna(rnd, n, 'ACGTN' if n == 1500 else 'ACGT')
        seed, read = 1000 + n, 7 * n
        joined, st = emu.error_loop(frag, ident, seed, read, em)
        seq, _, _, want = orc.sequence_fragment(frag, ident, seed, read, with_stats=True)
        assert {k: st[k] for k in ('loop_count', 'change_count', 'n_alignments', 'untrimmed_len')} == \
            {k: want[k] for k in ('loop_count', 'change_count', 'n_alignments', 'untrimmed_len')}, (model, n, ident)
        assert joined[st['start_trim']:len(joined) - st['end_trim']] == seq, (model, n, ident)
        most_rounds = max(most_rounds, st['rounds'])
    assert most_rounds >= (2 if model == 'nanopore2020' else 1)     # (2500 bases at 0.65: the horizon is hit, the loop resumes)


@pytest.mark.parametrize('qscore_model', ['nanopore2023', 'pacbio2021', 'ideal', 'random'])
def test_get_qscores_kernels_match_oracle(emu, qscore_model):
    """get_qscores (qscore_model.py:32-75) as the GPU computes it - the alignment task pipeline, then per base the CIGAR
    window, the hash look-up with the trim-by-one fall-back and `choices` on the base's own Philox stream
    (bb_k_qscores_pair) - under the emulator: quality string, '=' columns and alignment columns equal the oracle's."""
    from conftest import load_models
    from oracle import oracle as O
    em, qm = load_models('nanopore2023', qscore_model)
    orc = O.Oracle(em, qm)
    rnd = random.Random(9)
    for n, rate in ((1, 0.0), (50, 0.1), (800, 0.05), (3000, 0.08), (2200, 0.25), (1200, 0.0)):
        frag = random_dna(rnd, n, 'ACGTN' if n == 800 else 'ACGT')
        seq = mutate(rnd, frag, rate) or 'A'
        d = O.align_path(seq, frag)[1]
        assert emu.get_qscores(seq, frag, d + rnd.choice([0, 5, 40]), qm, 321, 17 + n) == orc.get_qscores(seq, frag, 321, 17 + n), \
            (qscore_model, n, rate)


@pytest.mark.parametrize('models', [('nanopore2023', 'nanopore2023'), ('pacbio2021', 'pacbio2021'), ('nanopore2020', 'ideal')])
def test_whole_read_on_device_code_matches_oracle(emu, models):
    """simulate.sequence_fragment (simulate.py:256-358) end to end on the DEVICE code, without a GPU: error loop -> bb_k_join
    -> alignment task pipeline -> quality scores, with the bound and the trims the kernels computed; sequence, quality
    string and matches / columns equal the oracle's for the same seed and read index."""
    from conftest import load_models
    from oracle import oracle as O
    em, qm = load_models(*models)
    orc = O.Oracle(em, qm)
    rnd = random.Random(515)
    for n, ident in ((60, 0.9), (700, 0.88), (1800, 0.95), (5200, 0.8), (2600, 1.0)):
        frag = random_dna(rnd, n)
        seed, read = 31 + n, 3 * n + 1
        joined, st = emu.error_loop(frag, ident, seed, read, em)
        qual, matches, columns = emu.get_qscores(joined, st['padded_fragment'], st['upper'], qm, seed, read)
        lo, hi = st['start_trim'], len(joined) - st['end_trim']
        seq, want_qual, _, want = orc.sequence_fragment(frag, ident, seed, read, with_stats=True)
        assert (joined[lo:hi], qual[lo:hi]) == (seq, want_qual), (models, n, ident)
        assert (matches, columns) == (want['matches'], want['columns']), (models, n, ident)


def test_fragment_builder_and_compaction_kernels(emu):
    """bb_k_build_fragments under the emulator: reference slices of either strand (misc.reverse_complement with its IUPAC
    table, misc.py:56-71) and literal bytes gathered behind each other, the 2k pad bases from the read's Philox stream
    (the oracle pads the same way), slots reset, the k-mer row of every position, the fragment's match bitmap; and
    bb_k_compact's trim."""
    from conftest import load_models
    from badread_b200.misc import reverse_complement
    from oracle import oracle as O
    em, _ = load_models('nanopore2023', 'nanopore2023')
    t = em.to_device_tables()
    k = int(t['k'])
    rnd = random.Random(64)
    ref = random_dna(rnd, 5000, 'ACGTNRY')
    lit = 'AATGTACTTCGTTCAGTTACGTATTGCT' + random_dna(rnd, 300)
    for it in range(6):
        segs, want = [], []
        for _ in range(rnd.randrange(1, 6)):
            kind = rnd.randrange(3)
            pool = lit if kind == 2 else ref
            n = rnd.randrange(1, 700 if kind != 2 else 120)
            src = rnd.randrange(0, len(pool) - n)
            segs.append((kind, src, n))
            piece = pool[src:src + n]
            want.append(reverse_complement(piece) if kind == 1 else piece)
        body = ''.join(want)
        seed, read = 900 + it, 17 * it + 2
        frag, kidx, status = emu.build_fragment(ref, lit, segs, k, t['kmer_to_row'], seed, read)
        assert status == 0
        assert frag[k:len(frag) - k] == body
        rng = O.Rng(O.RNG_PHILOX, seed, read)
        rng.stream(2, 0)                                    # BB_PURPOSE_PAD
        pads = ''.join('ACGT'[rng.randbelow(4)] for _ in range(2 * k))
        assert frag[:k] + frag[len(frag) - k:] == pads
        for x in (0, 1, len(frag) // 2, len(frag) - k):
            kmer = frag[x:x + k]
            code = 0
            for c in kmer:
                code = code * 4 + 'ACGT'.find(c)
            assert kidx[x] == (int(t['kmer_to_row'][code]) if set(kmer) <= set('ACGT') else -1)
