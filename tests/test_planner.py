"""
The native fragment builder / FASTQ assembly (csrc/bb_planner.cpp, no GPU needed) against the Python planner
(`simulate.ReadPlanner`, which draws from the real `random.Random` / numpy `RandomState`): descriptors, identities,
header info and read names must be identical read for read.  Covers the parameter sets of BASELINE.json's configs.
"""
import io
import os
import uuid

import numpy as np
import pytest

from badread_b200 import simulate as S
from badread_b200.__main__ import check_simulate_args, parse_args
from badread_b200.fragment_lengths import FragmentLengths
from badread_b200.identities import Identities


def _fasta(tmp_path, contigs):
    p = tmp_path / 'ref.fasta'
    rs = np.random.RandomState(7)
    with open(p, 'w') as f:
        for header, n in contigs:
            f.write(f'>{header}\n')
            f.write(np.frombuffer(b'ACGT', dtype=np.uint8)[rs.randint(0, 4, n)].tobytes().decode() + '\n')
    return str(p)


def _setup(tmp_path, contigs, extra, seed=5):
    from badread_b200.planner import NativePlanner
    path = _fasta(tmp_path, contigs)
    args = parse_args(['simulate', '--reference', path, '--quantity', '10x', '--seed', str(seed)] + extra)
    check_simulate_args(args)
    sink = io.StringIO()
    ref = S.Reference(args.reference, sink)
    fl = FragmentLengths(args.mean_frag_length, args.frag_length_stdev, sink)
    S.adjust_depths(ref, fl, args, np.random.RandomState(seed))
    ids = Identities(args.mean_identity, args.identity_stdev, args.max_identity, sink)
    return args, ref, S.ReadPlanner(args, ref, fl, ids, seed), NativePlanner(args, ref, fl, ids, seed, n_threads=4)


CASES = {
    'config2_defaults': ([('chr1 circular=true', 200000)], []),
    'config3_glitchy': ([('chr1 circular=true', 200000)], ['--identity', '90,98,5', '--glitches', '1000,100,100']),
    'config4_multi_contig_chimeras': ([('c1', 300000), ('c2', 250000), ('p1 depth=2 circular=true', 30000),
                                       ('p2 depth=10 circular=true', 5000)], ['--chimeras', '10', '--length', '8000,7000']),
    'config5_long': ([('a', 400000), ('b', 350000)], ['--length', '40000,20000']),
    'hairpins_linear_short': ([('h1 hairpin_left=true hairpin_right=true', 9000), ('l2', 7000)], ['--length', '6000,5000']),
    'adapters_half_no_glitch': ([('chr1 circular=true', 100000)],
                                ['--start_adapter', '50,50', '--end_adapter', '100,100', '--glitches', '0,0,0',
                                 '--junk_reads', '10', '--random_reads', '10', '--length', '3000,2000']),
    'constant_length_normal_identity': ([('chr1', 50000)], ['--length', '2000,0', '--identity', '15,3']),
    'constant_identity_small_glitches': ([('chr1 circular=true', 50000)],
                                         ['--identity', '92,92,0', '--glitches', '50,1,1', '--length', '1500,1200',
                                          '--start_adapter_seq', '', '--end_adapter', '20,80']),
}


@pytest.mark.parametrize('case', sorted(CASES))
def test_native_planner_matches_python_planner(tmp_path, case):
    contigs, extra = CASES[case]
    args, ref, py, nat = _setup(tmp_path, contigs, extra)
    n = 1200
    first, stride = 3, 2
    pb = nat.plan(first, n, stride=stride)
    assert len(pb) == n
    from badread_b200.engine import FragmentBatch
    batch = FragmentBatch()
    for i in range(n):
        idx = first + stride * i
        pieces, info, ident, name = py.plan(idx)
        py.add_to_batch(batch, idx, pieces, ident)
        assert pb.info_str(i) == ' '.join(info), (case, i)
        assert pb.name_str(i) == str(name), (case, i)
        assert pb.target_identity[i] == ident, (case, i)
        assert pb.frag_len[i] == sum(p.length for p in pieces)
        if i % 97 == 0:
            assert pb.fragment(i, ref.concat) == py.materialise(pieces), (case, i)
    ri, so, segs, lit, lit_len, ti = batch.arrays()
    assert np.array_equal(pb.read_index, ri)
    assert np.array_equal(pb.seg_off, so)
    want = np.frombuffer(segs, dtype=pb.segs.dtype)[:len(batch.seg_src)]
    # literal offsets are positions in the (identically built) literal pool
    assert np.array_equal(pb.segs['len'], want['len']) and np.array_equal(pb.segs['kind'], want['kind'])
    assert np.array_equal(pb.segs['src'], want['src'])
    assert pb.literal_len == lit_len and bytes(pb.literals[:lit_len]) == bytes(lit[:lit_len])
    nat.close()


def test_native_planner_is_thread_count_independent_and_reports_impossible_reads(tmp_path):
    contigs, extra = CASES['config4_multi_contig_chimeras']
    args, ref, py, nat = _setup(tmp_path, contigs, extra)
    a = nat.plan(0, 900)
    snap = (a.segs.copy(), a.target_identity.copy(), bytes(a.info), a.names.copy())
    nat.n_threads = 1
    b = nat.plan(0, 900)
    assert np.array_equal(snap[0], b.segs) and np.array_equal(snap[1], b.target_identity)
    assert snap[2] == bytes(b.info) and np.array_equal(snap[3], b.names)
    nat.close()
    # BASELINE.json configs[0]: three tiny circular contigs and 15 kb fragments -> the reference's own error message
    args, ref, py, nat = _setup(tmp_path, [('a circular=true', 30), ('b circular=true', 30), ('c circular=true', 30)],
                                ['--small_plasmid_bias'], seed=1)
    with pytest.raises(SystemExit) as e:
        nat.plan(0, 200)
    assert 'failed to generate any sequence fragments' in str(e.value)


def test_fastq_format_matches_python_records(tmp_path):
    """bb_fastq_format against the four print() calls of simulate.py:73-86, including the skip of empty reads and the
    stop at the target."""
    from badread_b200._lib import ReadResult
    from badread_b200.planner import fastq_format
    contigs, extra = CASES['config2_defaults']
    args, ref, py, nat = _setup(tmp_path, contigs, ['--length', '300,200'])
    n = 400
    pb = nat.plan(10, n)
    rs = np.random.RandomState(3)
    results = (ReadResult * n)()
    lens = rs.randint(0, 500, n)
    lens[::17] = 0
    off = np.concatenate([[0], np.cumsum(lens)])
    perm = rs.permutation(n)            # blocks are packed, not in batch order
    start = np.zeros(n, dtype=np.int64)
    pos = 0
    for r in perm:
        start[r] = pos
        pos += lens[r]
    seq = np.frombuffer(b'ACGT', dtype=np.uint8)[rs.randint(0, 4, pos)].copy()
    qual = (rs.randint(0, 60, pos) + 33).astype(np.uint8)
    for i in range(n):
        results[i].out_off, results[i].out_len, results[i].frag_len = int(start[i]), int(lens[i]), int(pb.frag_len[i])
        results[i].columns = int(lens[i]) + 7
        results[i].matches = int(rs.randint(0, lens[i] + 1))
    for target, sofar in ((10 ** 12, 0), (5000, 1200), (1, 0)):
        buf, n_emit, bases, nxt, _ = fastq_format(pb, results, seq, qual, 0, sofar, target, n_threads=3)
        want, total, count, i = [], sofar, 0, 0
        while i < n and total < target:
            r = results[i]
            if r.out_len:
                s = bytes(seq[r.out_off:r.out_off + r.out_len]).decode()
                q = bytes(qual[r.out_off:r.out_off + r.out_len]).decode()
                ident = r.matches / r.columns
                info = [pb.info_str(i), f'length={len(s)}', f'error-free_length={r.frag_len}', f'read_identity={ident * 100.0:.3f}%']
                want.append(f'@{uuid.UUID(pb.name_str(i))} {" ".join(info)}\n{s}\n+\n{q}\n')
                total += len(s)
                count += 1
            i += 1
        assert bytes(buf).decode('latin-1') == ''.join(want)
        assert (n_emit, bases, nxt) == (count, total - sofar, i)
    nat.close()
