#!/usr/bin/env python3
"""Experiment: the bench workload as S sub-batches on S contexts of one GPU, driven by S host threads.
Usage: overlap_test.py [S ...]"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.realpath(__file__)), '..'))
import bench  # noqa: E402
from badread_b200.engine import Engine  # noqa: E402

planner, ref, models, plans, indices = bench.build_workload(0, 1)
for S in [int(a) for a in sys.argv[1:]] or [1, 2]:
    engines, batches = [], []
    for s in range(S):
        eng = Engine(0, seed=1)
        eng.upload_reference(ref.concat)
        eng.set_error_model(models[0])
        eng.set_qscore_model(models[1])
        engines.append(eng)
        batches.append(bench.make_batch(planner, plans[s::S], indices[s::S]))
    totals = [0] * S

    def work(s):
        _, totals[s] = engines[s].sequence_batch(batches[s])

    for rep in range(4):
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(s,)) for s in range(S)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        print(f'S={S} rep={rep} wall_ms={dt * 1e3:.1f} bases={sum(totals)} Gb/s={sum(totals) / dt / 1e9:.3f}', flush=True)
    del engines
