#!/usr/bin/env python3
"""A small batch that reaches every alignment kernel (wide roots -> warp pairs, lean warp nodes, lane nodes and
leaves, window lanes and their warp fallback); meant to be run under compute-sanitizer."""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.realpath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.realpath(__file__)), '..', 'tests'))
from conftest import load_models, random_dna  # noqa: E402
from badread_b200.engine import Engine, FragmentBatch  # noqa: E402
from oracle import oracle as O  # noqa: E402

em, qm = load_models('nanopore2023', 'nanopore2023')
eng = Engine(0, seed=9)
eng.set_error_model(em)
eng.set_qscore_model(qm)
rnd = random.Random(3)
batch = FragmentBatch()
cases = [(24000, 0.75), (9000, 0.85), (6000, 0.95), (3000, 0.99), (1200, 0.9), (300, 0.8)]
frags = []
for i, (n, ident) in enumerate(cases):
    frags.append(random_dna(rnd, n, 'ACGTN' if i == 2 else 'ACGT'))
    batch.add_literal_read(70 + i, frags[-1], ident)
res, total = eng.sequence_batch(batch)
orc = O.Oracle(em, qm)
for i, (n, ident) in enumerate(cases):
    s, q, _ = orc.sequence_fragment(frags[i], ident, 9, read_index=70 + i)
    assert res.read(i) == (s, q), i
print('sanitize case ok:', total, 'bases,', eng.launch_count(), 'launches')
