#!/usr/bin/env python3
"""Latency chain of the longest reads: a batch of a few long reads alone on the GPU, with the launch trace
(BADREAD_B200_TRACE=1).  Shows what bounds a step from below whatever the throughput: the dependent stages of one read.
Usage (GPU box): BADREAD_B200_TRACE=1 BADREAD_B200_SUBBATCHES=1 python tools/chain_probe.py [length] [n_reads] [identity]"""
import csv
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.realpath(__file__)), '..'))
from badread_b200.engine import Engine, FragmentBatch  # noqa: E402
from badread_b200.error_model import ErrorModel  # noqa: E402
from badread_b200.qscore_model import QScoreModel  # noqa: E402

length = int(sys.argv[1]) if len(sys.argv) > 1 else 150000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ident = float(sys.argv[3]) if len(sys.argv) > 3 else 0.95
sink = io.StringIO()
eng = Engine(device=0, seed=1)
eng.set_error_model(ErrorModel('nanopore2023', sink))
eng.set_qscore_model(QScoreModel('nanopore2023', sink))
batch = FragmentBatch()
for i in range(n_reads):
    dna = np.frombuffer(b'ACGT', dtype=np.uint8)[np.random.RandomState(i).randint(0, 4, length)].tobytes().decode()
    batch.add_literal_read(i, dna, ident)
for _ in range(3):
    eng.upload_batch(batch)
    eng.run_batch()
    total, stages = eng.last_run_ms()
res, bases = eng.fetch_batch()
print(f'{n_reads} reads of {length} at identity {ident}: {total:.2f} ms', {k: round(v, 2) for k, v in stages.items()})
print('loop kcycles per read', [res.records[i].loop_kcycles for i in range(n_reads)])
path = '/tmp/chain_trace.csv'
eng.trace_dump(path)
rows = list(csv.DictReader(open(path)))
for r in rows:
    d = float(r['end_ms']) - float(r['begin_ms'])
    if d > 0.05 and r['name'] != 'fork':
        print(f"  s{r['stream']} {r['name']:14s} {float(r['begin_ms']):8.2f} -> {float(r['end_ms']):8.2f}  ({d:.2f} ms)")
