#!/usr/bin/env python3
"""
make_golden.py - generates tests/golden/*.json by running the UNMODIFIED reference (/root/reference) in this
container, with `edlib` supplied by oracle/edlib_shim (the real edlib wheel is absent here: parity is unpinned at
that boundary).  Run once here; the fixtures are committed, the reference is never needed at test time.

  golden_sequence_fragment.json  reference simulate.sequence_fragment outputs under random.seed(s)  (pins the oracle's
                                 MT mode: loop, samplers, trims, qscores)
  golden_tables.json             sha256 of the reference's ErrorModel.alternatives/.probabilities and QScoreModel
                                 tables for every built-in model (pins the precompiled tables in badread_b200/models)
  golden_align_kmers.json        error_model.align_kmers samples (pins the host table builder)
  golden_get_qscores.json        qscore_model.get_qscores outputs for hand-made pairs
"""
import hashlib
import io
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.realpath(__file__))
sys.path.insert(0, os.path.join(HERE, 'edlib_shim'))
sys.path.insert(0, os.path.join(HERE, '..'))
sys.path.insert(0, '/root/reference')

import badread.error_model as rem  # noqa: E402
import badread.qscore_model as rqm  # noqa: E402
import badread.simulate as rsim  # noqa: E402

OUT = os.path.join(HERE, '..', 'tests', 'golden')


def table_digest(d1, d2):
    h = hashlib.sha256()
    for key in sorted(d1.keys()):
        h.update(repr((key, d1[key], d2[key])).encode())
    return h.hexdigest()


def main():
    sink = io.StringIO()
    rnd = random.Random(20240924)
    ems = {name: rem.ErrorModel(name, sink) for name in ('random', 'nanopore2023', 'nanopore2020', 'pacbio2021')}
    qms = {name: rqm.QScoreModel(name, sink) for name in ('random', 'ideal', 'nanopore2023', 'nanopore2020', 'pacbio2021')}
    # digests first: add_errors_to_kmer mutates the reference's tables in place (error_model.py:151-154)
    tables = {}
    for name in ('nanopore2018', 'nanopore2020', 'nanopore2023', 'pacbio2016', 'pacbio2021'):
        em = ems.get(name) or rem.ErrorModel(name, sink)
        qm = qms.get(name) or rqm.QScoreModel(name, sink)
        tables[name] = {'error': table_digest(em.alternatives, em.probabilities), 'kmer_size': em.kmer_size,
                        'qscore': table_digest(qm.scores, qm.probabilities), 'qscore_kmer_size': qm.kmer_size}
    json.dump(tables, open(os.path.join(OUT, 'golden_tables.json'), 'w'), indent=1)

    cases = []
    combos = [('random', 'random'), ('random', 'ideal'), ('nanopore2023', 'nanopore2023'), ('nanopore2020', 'nanopore2020'),
              ('pacbio2021', 'pacbio2021'), ('nanopore2023', 'ideal')]
    for em_name, qm_name in combos:
        for length, ident in ((1, 0.9), (12, 0.8), (150, 0.95), (985, 0.9), (986, 0.92), (1000, 0.85), (1600, 0.97),
                              (3000, 0.9), (2500, 1.0), (4000, 0.6)):
            alphabet = 'ACGT' if length % 2 == 0 else 'ACGTN'
            frag = ''.join(rnd.choice(alphabet) for _ in range(length))
            seed = rnd.randint(0, 2 ** 32 - 1)
            random.seed(seed)
            seq, qual, actual, _ = rsim.sequence_fragment(frag, ident, ems[em_name], qms[qm_name])
            cases.append({'error_model': em_name, 'qscore_model': qm_name, 'fragment': frag, 'identity': ident,
                          'seed': seed, 'seq': seq, 'qual': qual, 'actual_identity': actual})
    json.dump(cases, open(os.path.join(OUT, 'golden_sequence_fragment.json'), 'w'))

    kmers = []
    for _ in range(400):
        k = rnd.choice([4, 5, 7, 7, 7])
        kmer = ''.join(rnd.choice('ACGT') for _ in range(k))
        inner = list(kmer[1:-1])
        for _ in range(rnd.randint(0, 3)):
            op = rnd.choice('sid')
            if op == 's' and inner:
                inner[rnd.randrange(len(inner))] = rnd.choice('ACGT')
            elif op == 'i':
                inner.insert(rnd.randint(0, len(inner)), rnd.choice('ACGT'))
            elif op == 'd' and inner:
                del inner[rnd.randrange(len(inner))]
        alt = kmer[0] + ''.join(inner) + kmer[-1]
        kmers.append({'kmer': kmer, 'alt': alt, 'slots': rem.align_kmers(kmer, alt)})
    json.dump(kmers, open(os.path.join(OUT, 'golden_align_kmers.json'), 'w'))

    qs = []
    for qm_name in ('ideal', 'nanopore2023'):
        for _ in range(6):
            n = rnd.choice([9, 40, 300, 1500])
            frag = ''.join(rnd.choice('ACGT') for _ in range(n))
            seq = list(frag)
            for _ in range(max(1, n // 15)):
                p = rnd.randrange(len(seq))
                op = rnd.choice('sid')
                if op == 's':
                    seq[p] = rnd.choice('ACGT')
                elif op == 'i':
                    seq.insert(p, rnd.choice('ACGT'))
                elif len(seq) > 2:
                    del seq[p]
            seq = ''.join(seq)
            seed = rnd.randint(0, 2 ** 32 - 1)
            random.seed(seed)
            qual, actual, by_q = rqm.get_qscores(seq, frag, qms[qm_name])
            qs.append({'qscore_model': qm_name, 'seq': seq, 'frag': frag, 'seed': seed, 'qual': qual,
                       'actual_identity': actual, 'identity_by_qscores': by_q})
    json.dump(qs, open(os.path.join(OUT, 'golden_get_qscores.json'), 'w'))
    print('wrote', len(cases), 'sequence_fragment cases,', len(kmers), 'align_kmers cases,', len(qs), 'get_qscores cases')


if __name__ == '__main__':
    main()
