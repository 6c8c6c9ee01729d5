"""Host-side logic (no GPU): C ABI exports, CLI validation messages, loaders, planner determinism, model surface."""
import ctypes
import io
import os
import random
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))


def test_library_exports_every_declared_symbol():
    from badread_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'badread_b200.h')).read()
    declared = set(re.findall(r'BB_API [^;(]*?\**(bb_[a-z_0-9]+)\(', header))
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.lib().bb_version().decode().startswith('badread_b200')


def test_no_gpu_means_loud_failure():
    """There is no CPU path: without a device bb_create fails and the Python layer raises."""
    from badread_b200.engine import Engine, EngineError
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip('a GPU is present')
    with pytest.raises(EngineError):
        Engine(device=0, seed=1)


def _args(extra):
    from badread_b200.__main__ import parse_args
    return parse_args(['simulate', '--reference', __file__, '--quantity', '10x'] + extra)


@pytest.mark.parametrize('extra,message', [
    (['--chimeras', '60'], 'Error: --chimeras cannot be greater than 50'),
    (['--junk_reads', '70', '--random_reads', '40'], 'Error: --junk_reads and --random_reads cannot sum to more than 100'),
    (['--length', 'abc'], 'Error: could not parse --length values'),
    (['--length', '50,10'], 'Error: mean read length must be at least 100'),
    (['--identity', '95,90,5'], 'Error: mean identity (95.0) cannot be larger than max identity (90.0)'),
    (['--identity', '40,99,2'], 'Error: mean read identity must be at least 50'),
    (['--glitches', '1,2'], 'Error: could not parse --glitches values'),
    (['--start_adapter_seq', 'ACGX'], 'Error: --start_adapter_seq must be a DNA sequence or a number'),
    (['--error_model', 'nonexistent_model'], 'Error: nonexistent_model is not a file\n  --error_model must be from'),
])
def test_cli_validation_messages(extra, message):
    """Same messages as /root/reference/badread/__main__.py:239-336 (test/test_cli.py)."""
    from badread_b200.__main__ import check_simulate_args
    with pytest.raises(SystemExit) as e:
        check_simulate_args(_args(extra))
    assert str(e.value).startswith(message)


def test_cli_defaults():
    from badread_b200.__main__ import check_simulate_args
    a = _args([])
    check_simulate_args(a)
    assert (a.mean_frag_length, a.frag_length_stdev) == (15000.0, 13000.0)
    assert (a.mean_identity, a.max_identity, a.identity_stdev) == (95.0, 99.0, 2.5)
    assert (a.glitch_rate, a.glitch_size, a.glitch_skip) == (10000.0, 25.0, 25.0)
    assert a.error_model == 'nanopore2023' and a.qscore_model == 'nanopore2023'


def test_target_size_known_answers():
    """test/test_target_size.py."""
    from badread_b200.simulate import get_target_size
    assert get_target_size(1000, '5000') == 5000
    assert get_target_size(1000, '25x') == 25000
    assert get_target_size(1000, '1.5X') == 1500
    assert get_target_size(1000, '250M') == 250000000
    assert get_target_size(1000, '2g') == 2000000000
    assert get_target_size(1000, '7.5k') == 7500
    with pytest.raises(SystemExit):
        get_target_size(1000, 'abc')


def test_fasta_loader_and_reverse_complement(tmp_path):
    from badread_b200.misc import load_fasta, reverse_complement
    p = tmp_path / 'ref.fasta'
    p.write_text('>A depth=2.5 circular=true\nacgt\nACGN\n>B hairpin_left=true HAIRPIN_RIGHT=TRUE\nTTTT\n>C depth=x\nGG\n')
    seqs, depths, circ, hl, hr = load_fasta(str(p))
    assert list(seqs.items()) == [('A', 'ACGTACGN'), ('B', 'TTTT'), ('C', 'GG')]
    assert depths == {'A': 2.5, 'B': 1.0, 'C': 1.0}
    assert circ == {'A': True, 'B': False, 'C': False}
    assert hl['B'] and hr['B'] and not hl['A']
    assert reverse_complement('ACGTNRYx') == 'NRYNACGT'
    import gzip
    g = tmp_path / 'ref.fasta.gz'
    with gzip.open(str(g), 'wt') as f:
        f.write('>Z\nACGT\n')
    assert load_fasta(str(g))[0]['Z'] == 'ACGT'


def _planner(tmp_path, extra=()):
    from badread_b200 import simulate as S
    from badread_b200.__main__ import check_simulate_args, parse_args
    from badread_b200.fragment_lengths import FragmentLengths
    from badread_b200.identities import Identities
    rs = np.random.RandomState(5)
    lines = ['>lin depth=1\n', bytes(np.frombuffer(b'ACGT', dtype=np.uint8)[rs.randint(0, 4, 20000)]).decode() + '\n',
             '>circ depth=3 circular=true\n', bytes(np.frombuffer(b'ACGT', dtype=np.uint8)[rs.randint(0, 4, 8000)]).decode() + '\n',
             '>hp hairpin_right=true hairpin_left=true\n', bytes(np.frombuffer(b'ACGT', dtype=np.uint8)[rs.randint(0, 4, 3000)]).decode() + '\n']
    ref_path = tmp_path / 'ref.fasta'
    ref_path.write_text(''.join(lines))
    args = parse_args(['simulate', '--reference', str(ref_path), '--quantity', '5x', '--length', '3000,2000', '--seed', '3',
                       '--glitches', '1000,25,25', '--chimeras', '20'] + list(extra))
    check_simulate_args(args)
    out = io.StringIO()
    ref = S.Reference(args.reference, out)
    fl = FragmentLengths(args.mean_frag_length, args.frag_length_stdev, out)
    S.adjust_depths(ref, fl, args, np.random.RandomState(3))
    ids = Identities(args.mean_identity, args.identity_stdev, args.max_identity, out)
    return S, S.ReadPlanner(args, ref, fl, ids, 3), ref


def test_planner_is_deterministic_per_read_and_descriptors_match_strings(tmp_path):
    """A read's plan depends only on (seed, read index): any sharding / batching gives the same reads, and the
    segment descriptors the GPU gathers reproduce the fragment string."""
    from badread_b200.engine import FragmentBatch
    from badread_b200.misc import reverse_complement
    S, planner, ref = _planner(tmp_path)
    plans = [planner.plan(i) for i in range(120)]
    again = [planner.plan(i) for i in reversed(range(120))][::-1]
    kinds = set()
    for a, b in zip(plans, again):
        assert planner.materialise(a[0]) == planner.materialise(b[0]) and a[1:] == b[1:]
    concat = ref.concat.tobytes()
    for i, (pieces, info, ident, name) in enumerate(plans):
        batch = FragmentBatch()
        planner.add_to_batch(batch, i, pieces, ident)
        ri, so, segs, lit, lit_len, ti = batch.arrays()
        rebuilt = []
        for s in range(so[0], so[1]):
            src, ln, kind = segs[s].src, segs[s].len, segs[s].kind
            kinds.add(kind)
            if kind == 0:
                rebuilt.append(concat[src:src + ln])
            elif kind == 1:
                rebuilt.append(reverse_complement(concat[src:src + ln]))
            else:
                rebuilt.append(bytes(lit[src:src + ln]))
        assert b''.join(rebuilt).decode() == planner.materialise(pieces)
        assert 0.0 <= ident <= 1.0
        assert any(tag in ' '.join(info) for tag in ('strand', 'junk_seq', 'random_seq'))
    assert kinds == {0, 1, 2}


def test_fragment_builder_semantics(tmp_path):
    """Circular wrap, hairpin read-through and whole-contig cases of get_real_fragment (simulate.py:183-246)."""
    S, planner, ref = _planner(tmp_path)
    rng = random.Random(1)
    seen = set()
    for _ in range(3000):
        pieces, info = planner.get_real_fragment(rng.choice([rng.randint(500, 6000), 25000]), rng)
        if not pieces:
            continue
        text = ','.join(info)
        total = sum(p.length for p in pieces)
        if 'hairpin' in text:
            seen.add('hairpin')
            assert len(pieces) == 2 and pieces[0].strand != pieces[1].strand
        elif info[0] == 'circ' and len(pieces) == 2:
            seen.add('wrap')
            start, end = [int(x) for x in info[2].split('-')]
            assert end - start == total and pieces[1].start == 0
        elif info[2].startswith('0-') and total == ref.lengths[ref.names.index(info[0])]:
            seen.add('whole')
    assert {'hairpin', 'wrap', 'whole'} <= seen


def test_model_plugin_surface():
    from badread_b200.error_model import ErrorModel, add_one_random_change
    from badread_b200.qscore_model import QScoreModel
    out = io.StringIO()
    em = ErrorModel('random', out)
    assert em.type == 'random' and em.kmer_size == 1
    random.seed(4)
    variants = {tuple(add_one_random_change('ACCA')) for _ in range(6000)}
    assert len(variants) == 44  # test/test_error_model.py: 44 distinct one-change variants of a 4-mer
    em = ErrorModel('nanopore2023', out)
    assert em.type == 'model' and em.kmer_size == 7 and len(em.alternatives) == 16384
    assert em.alternatives['AAAAAAA'][0] == list('AAAAAAA')
    random.seed(5)
    picks = [''.join(em.add_errors_to_kmer('ACGTACG')) for _ in range(300)]
    assert picks.count('ACGTACG') > 200
    qm = QScoreModel('ideal', out)
    assert qm.kmer_size == 9 and set(qm.scores) == {'X', 'I', '=', '===', '=====', '=======', '========='}
    random.seed(6)
    assert all(41 <= ord(qm.get_qscore('=========')) - 33 <= 50 for _ in range(50))
    assert all(1 <= ord(qm.get_qscore('==X==')) - 33 <= 3 for _ in range(50))
    t = qm.to_device_tables()
    assert t['n_keys'] == 7 and t['kmer_size'] == 9


def test_error_model_file_loading(tmp_path):
    """A user-supplied model file goes through the host table builder (error_model.py:111-133, 179-229)."""
    from badread_b200.error_model import ErrorModel
    p = tmp_path / 'model4'
    p.write_text('GCCA,0.80;GCCCA,0.05;GCA,0.05;GTCA,0.04;\nACGT,1.0;\nAAAA,0.7;AAA,0.2;AAAAA,0.1;\n')
    em = ErrorModel(str(p), io.StringIO())
    assert em.kmer_size == 4 and set(em.alternatives) == {'GCCA', 'ACGT', 'AAAA'}
    assert em.alternatives['GCCA'][1] in (['G', 'CC', 'C', 'A'], ['G', 'C', 'CC', 'A'])
    assert em.alternatives['GCCA'][2] in (['G', '', 'C', 'A'], ['G', 'C', '', 'A'])
    assert em.probabilities['GCCA'] == [0.80, 0.05, 0.05, 0.04]
    t = em.to_device_tables()
    flags = t['flags'][t['row_off'][0]:t['row_off'][1]]
    assert list(flags) == [1, 0, 0, 0, 2]       # remainder entry appended because the row sums to 0.94
    flags = t['flags'][t['row_off'][1]:t['row_off'][2]]
    assert list(flags) == [1]                   # sums to 1.0: no remainder entry


def test_library_rebuilds_when_any_device_header_changes():
    """The in-tree build must pick up edits to every header (a stale libbadread_b200.so silently runs old kernels):
    every object depends on all .cuh / .h files, and the library on every object."""
    import pathlib
    csrc = pathlib.Path(__file__).resolve().parent.parent / 'badread_b200' / 'csrc'
    mk = (csrc / 'Makefile').read_text()
    hdrs = [ln for ln in mk.splitlines() if ln.startswith('HDRS')]
    assert len(hdrs) == 1 and '$(wildcard *.cuh)' in hdrs[0] and '$(wildcard *.h)' in hdrs[0] and 'badread_b200.h' in hdrs[0]
    obj_rules = [ln for ln in mk.splitlines() if ln.startswith('$(BUILD)/') and ':' in ln]
    assert obj_rules and all('$(HDRS)' in ln for ln in obj_rules)
    assert '$(wildcard bb_tu_*.cu)' in mk and 'bb_api.cu' in mk and 'bb_host.o' in mk and 'bb_planner.o' in mk
    assert [ln for ln in mk.splitlines() if ln.startswith('$(OUT): $(OBJS)')]


def test_numpy_fasta_loader_matches_line_parser(tmp_path):
    """load_fasta_arrays (the loader `simulate` uses; numpy, no per-line Python) against the line-by-line restatement
    of misc.load_fasta (misc.py:122-153): header flags, depth parsing, upper-casing, blank lines, CRLF, a repeated
    name, an empty record, gzip."""
    import gzip
    import numpy as np
    from badread_b200.misc import load_fasta, load_fasta_arrays
    rs = np.random.RandomState(4)
    body = ''.join('acgtnACGTN'[i] for i in rs.randint(0, 10, 5000))
    txt = ('>c1 depth=2.5 circular=true extra\n' + '\n'.join(body[i:i + 70] for i in range(0, 3000, 70)) + '\n\n'
           '>c2 hairpin_left=TRUE hairpin_right=true\r\n' + body[3000:4000] + '\r\n' + body[4000:] + '\n>c3\n>c1 depth=x\nAAAA')
    for gz in (False, True):
        p = tmp_path / ('ref.fa' + ('.gz' if gz else ''))
        with (gzip.open(p, 'wt') if gz else open(p, 'w')) as f:
            f.write(txt)
        seqs, depths, circular, left, right = load_fasta(str(p))
        names, arrays, d2, c2, l2, r2 = load_fasta_arrays(str(p))
        assert names == list(seqs.keys())
        assert [a.tobytes().decode() for a in arrays] == [seqs[n] for n in names]
        assert (depths, circular, left, right) == (d2, c2, l2, r2)
