#!/usr/bin/env python3
"""Product multi-GPU path check (run on a box with >= 2 GPUs):  `badread_b200 simulate --gpus N` must write the same
FASTQ, byte for byte, as `--gpus 1` (reads shard by index, the stop condition's SUM goes through NCCL inside the
library), and report its timing line.  Usage: python tools/multigpu_check.py [N] [quantity]"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.realpath(__file__)), '..')
n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
quantity = sys.argv[2] if len(sys.argv) > 2 else '20x'
rs = np.random.RandomState(1001)
fd, fasta = tempfile.mkstemp(suffix='.fasta')
with os.fdopen(fd, 'wb') as f:
    f.write(b'>chr1 circular=true\n' + np.frombuffer(b'ACGT', dtype=np.uint8)[rs.randint(0, 4, 5_000_000)].tobytes() + b'\n')
env = dict(os.environ, BADREAD_B200_TIMING='1', PYTHONPATH=ROOT)
out = {}
for g in (1, n_gpus):
    t0 = time.perf_counter()
    p = subprocess.run([sys.executable, '-m', 'badread_b200', 'simulate', '--reference', fasta, '--quantity', quantity,
                        '--seed', '1', '--gpus', str(g)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        print(p.stderr.decode()[-2000:])
        sys.exit(f'--gpus {g} failed with exit code {p.returncode}')
    line = [ln for ln in p.stderr.decode().splitlines() if ln.startswith('BADREAD_B200_TIMING ')][-1]
    st = json.loads(line.split(' ', 1)[1])
    out[g] = {'sha256': hashlib.sha256(p.stdout).hexdigest(), 'fastq_bytes': len(p.stdout), 'wall_s': round(wall, 2), **st}
    print(f'--gpus {g}:', json.dumps(out[g]))
os.unlink(fasta)
same = out[1]['sha256'] == out[n_gpus]['sha256']
print(json.dumps({'gpus': n_gpus, 'fastq_identical_to_1_gpu': same, 'nccl_stop_condition': out[n_gpus].get('nccl_stop_condition'),
                  'gbases_per_s_1': out[1]['bases'] / out[1]['batches_s'] / 1e9,
                  f'gbases_per_s_{n_gpus}': out[n_gpus]['bases'] / out[n_gpus]['batches_s'] / 1e9}))
sys.exit(0 if same else 1)
