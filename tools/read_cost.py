#!/usr/bin/env python3
"""Per-read device cost of one bench-workload pass: prints how the error-loop / final-alignment kilo-cycles
distribute over read length (diagnostics for load balance and tail)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.realpath(__file__)), '..'))
import bench  # noqa: E402
from badread_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else None
planner, ref, models, plans, indices = bench.build_workload(0, 1, n_reads_override=n)
batch = bench.make_batch(planner, plans, indices)
eng = Engine(0, seed=1)
eng.upload_reference(ref.concat)
eng.set_error_model(models[0])
eng.set_qscore_model(models[1])
eng.sequence_batch(batch)
for _ in range(int(os.environ.get('READ_COST_REPEATS', '1'))):
    res, total = eng.sequence_batch(batch)
    t, st = eng.last_run_ms()
    print('total ms %.1f' % t, {k: round(v, 1) for k, v in st.items() if v >= 0.5})
rec = res.records
L = np.array([rec[i].frag_len for i in range(len(plans))])
ka = np.array([rec[i].align_kcycles for i in range(len(plans))], dtype=np.float64)
kl = np.array([rec[i].loop_kcycles for i in range(len(plans))], dtype=np.float64)
ident = np.array([rec[i].matches / max(1, rec[i].columns) for i in range(len(plans))])
print('reads', len(L), 'bases', total, 'sum align Gcycles', ka.sum() / 1e6, 'sum loop Gcycles', kl.sum() / 1e6)
order = np.argsort(-ka)
print('top reads by align kcycles: (len, identity, align_kc, loop_kc, ms@1.9GHz)')
for i in order[:12]:
    print(L[i], round(ident[i], 3), int(ka[i]), int(kl[i]), round(ka[i] * 1024 / 1.9e6, 1))
bins = [0, 2000, 5000, 10000, 20000, 30000, 50000, 80000, 10 ** 9]
for lo, hi in zip(bins[:-1], bins[1:]):
    m = (L >= lo) & (L < hi)
    if m.any():
        print(f'len [{lo},{hi}) n={m.sum()} bases={L[m].sum()} align cyc/base={ka[m].sum() * 1024 / L[m].sum():.0f} '
              f'loop cyc/base={kl[m].sum() * 1024 / L[m].sum():.0f}')
