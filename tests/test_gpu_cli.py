"""`badread simulate` end to end on the GPU: FASTQ layout, determinism, independence of the batch size, and every
emitted read equal to the oracle's sequence_fragment for the same (fragment, identity, seed, read index)."""
import io
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(tmp_path, extra=(), batch_reads=16384):
    from badread_b200.__main__ import check_simulate_args, parse_args
    from badread_b200.simulate import simulate
    rs = np.random.RandomState(11)
    ref = tmp_path / 'ref.fasta'
    if not ref.exists():
        ref.write_text('>chr circular=true\n' + bytes(np.frombuffer(b'ACGT', dtype=np.uint8)[rs.randint(0, 4, 40000)]).decode() +
                       '\n>lin depth=2\n' + bytes(np.frombuffer(b'ACGT', dtype=np.uint8)[rs.randint(0, 4, 15000)]).decode() + '\n')
    args = parse_args(['simulate', '--reference', str(ref), '--quantity', '6x', '--length', '2500,1500', '--seed', '5',
                       '--glitches', '2000,20,20', '--chimeras', '5', '--batch_reads', str(batch_reads)] + list(extra))
    check_simulate_args(args)
    out, err = io.StringIO(), io.StringIO()
    simulate(args, output=err, stdout=out)
    return args, out.getvalue(), err.getvalue()


def test_simulate_fastq_deterministic_and_oracle_exact(tmp_path):
    from badread_b200 import simulate as S
    from badread_b200.error_model import ErrorModel
    from badread_b200.fragment_lengths import FragmentLengths
    from badread_b200.identities import Identities
    from badread_b200.qscore_model import QScoreModel
    from oracle import oracle as O
    args, fastq, banner = _run(tmp_path)
    _, fastq_small_batches, _ = _run(tmp_path, batch_reads=7)
    assert fastq == fastq_small_batches            # output depends on --seed only, not on batching
    assert 'Target read set size: 330,000 bp' in banner and 'Badread v' in banner
    lines = fastq.strip().split('\n')
    assert len(lines) % 4 == 0 and len(lines) >= 4 * 50
    records = [(lines[i][1:], lines[i + 1], lines[i + 3]) for i in range(0, len(lines), 4)]
    assert all(lines[i + 2] == '+' for i in range(0, len(lines), 4))
    total = sum(len(r[1]) for r in records)
    assert total >= 330000 and total - len(records[-1][1]) < 330000   # stops after the read that reaches the target
    # re-plan the reads on the host and push each fragment through the oracle
    sink = io.StringIO()
    ref = S.Reference(args.reference, sink)
    fl = FragmentLengths(args.mean_frag_length, args.frag_length_stdev, sink)
    S.adjust_depths(ref, fl, args, np.random.RandomState(5))
    planner = S.ReadPlanner(args, ref, fl, Identities(args.mean_identity, args.identity_stdev, args.max_identity, sink), 5)
    orc = O.Oracle(ErrorModel(args.error_model, sink), QScoreModel(args.qscore_model, sink))
    by_name = {r[0].split(' ')[0]: r for r in records}
    checked = 0
    for idx in range(len(records) + 5):
        pieces, info, ident, name = planner.plan(idx)
        rec = by_name.get(str(name))
        if rec is None:
            continue
        seq, qual, actual = orc.sequence_fragment(planner.materialise(pieces), ident, 5, read_index=idx)
        assert rec[1] == seq and rec[2] == qual
        assert f'length={len(seq)}' in rec[0] and f'read_identity={actual * 100.0:.3f}%' in rec[0]
        checked += 1
    assert checked == len(records)


def test_single_read_api_matches_oracle():
    """sequence_fragment / get_qscores with the reference's signatures (batch of one on the GPU)."""
    import random
    from badread_b200 import engine
    from badread_b200.qscore_model import get_qscores
    from badread_b200.simulate import sequence_fragment
    from conftest import load_models, random_dna
    from oracle import oracle as O
    em, qm = load_models('nanopore2023', 'nanopore2023')
    orc = O.Oracle(em, qm)
    engine.set_seed(77)
    rnd = random.Random(1)
    for i in range(3):
        frag = random_dna(rnd, 1500 + 700 * i)
        seq, qual, ident, by_q = sequence_fragment(frag, 0.9, em, qm)
        s, q, a = orc.sequence_fragment(frag, 0.9, 77, read_index=i)
        assert (seq, qual, ident) == (s, q, a) and 0.0 < by_q < 1.0
    frag = random_dna(rnd, 800)
    seq = frag[:300] + 'A' + frag[300:500] + frag[503:]
    qual, ident, by_q = get_qscores(seq, frag, qm)
    q2, m, c = orc.get_qscores(seq, frag, 77, read_index=3)    # the module-level counter: three reads came before
    assert qual == q2 and ident == m / c
    qual_b, _, _ = get_qscores(seq, frag, qm)
    assert qual_b == orc.get_qscores(seq, frag, 77, read_index=4)[0] and qual_b != qual   # successive calls are independent


@pytest.mark.parametrize('extra', [
    # BASELINE.json configs[2] flavour: nanopore2020 models, lower identity, heavy glitches
    ['--error_model', 'nanopore2020', '--qscore_model', 'nanopore2020', '--identity', '90,98,5', '--glitches', '1000,100,100'],
    # configs[3] flavour: pacbio2021 models, many chimeras
    ['--error_model', 'pacbio2021', '--qscore_model', 'pacbio2021', '--chimeras', '10'],
    # random / ideal models, normal-distributed qscore identities, junk and random reads
    ['--error_model', 'random', '--qscore_model', 'ideal', '--identity', '12,3', '--junk_reads', '10', '--random_reads', '10'],
])
def test_simulate_config_variants_match_oracle(tmp_path, extra):
    """Other corners of BASELINE.json's configs at small scale: every emitted read equals the oracle's."""
    from badread_b200 import simulate as S
    from badread_b200.error_model import ErrorModel
    from badread_b200.fragment_lengths import FragmentLengths
    from badread_b200.identities import Identities
    from badread_b200.qscore_model import QScoreModel
    from oracle import oracle as O
    args, fastq, _ = _run(tmp_path, extra=['--quantity', '2x'] + extra)
    lines = fastq.strip().split('\n')
    records = {lines[i][1:].split(' ')[0]: (lines[i][1:], lines[i + 1], lines[i + 3]) for i in range(0, len(lines), 4)}
    assert len(records) >= 10
    sink = io.StringIO()
    ref = S.Reference(args.reference, sink)
    fl = FragmentLengths(args.mean_frag_length, args.frag_length_stdev, sink)
    S.adjust_depths(ref, fl, args, np.random.RandomState(5))
    planner = S.ReadPlanner(args, ref, fl, Identities(args.mean_identity, args.identity_stdev, args.max_identity, sink), 5)
    orc = O.Oracle(ErrorModel(args.error_model, sink), QScoreModel(args.qscore_model, sink))
    checked = 0
    for idx in range(len(records) + 50):
        pieces, info, ident, name = planner.plan(idx)
        rec = records.get(str(name))
        if rec is None:
            continue
        seq, qual, actual = orc.sequence_fragment(planner.materialise(pieces), ident, 5, read_index=idx)
        assert (rec[1], rec[2]) == (seq, qual)
        checked += 1
    assert checked == len(records)
