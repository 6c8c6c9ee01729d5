#!/usr/bin/env python3
"""
bench.py - simulated Gbases/s of the Badread error-injection hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path (sequence_fragment for every read) over the BASELINE.json configs[1]
workload: 5 Mb synthetic circular reference (RandomState(1001)), 50x, nanopore2023 error + qscore models, default
identity / length / adapters / glitches / junk / random / chimeras, seed 1  (~17 k reads, ~250 Mbases).

  value     whole-job Gbases/s with the fragment descriptors already resident in HBM (bb_batch_run only)
  e2e       the same through bb_sequence_batch with HOST buffers: descriptor H2D + seq/qual D2H inside the timing
  roofline  dominant kernel: 3 algorithmic bytes per emitted base (1 reference read + 1 seq write + 1 qual write,
            SURVEY.md 8d) / its CUDA-event duration, against the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the CPU oracle port (oracle/badread_oracle.c, Philox mode, pthreads over reads) on a bounded
            sample of the same reads on this box's host cores
  --impl reference   times only that CPU port (the reference is pure Python + an un-vendored edlib and cannot
            travel to the GPU box; its C restatement is pinned byte-for-byte to it in tests/test_oracle_golden.py)

Multi-GPU (torchrun, one rank per GPU): reads shard by index (rank g owns indices g, g+N, ...), every rank
processes a full configs[1]-sized share (weak scaling); the only collectives are the barrier, the SUM of emitted
bases and the MAX of elapsed time (NCCL).
"""
import argparse
import io
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.realpath(__file__))
sys.path.insert(0, ROOT)

REF_BASES = 5_000_000
DEPTH = 50
SEED = 1


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_reference_fasta(path):
    rs = np.random.RandomState(1001)
    seq = np.frombuffer(b'ACGT', dtype=np.uint8)[rs.randint(0, 4, REF_BASES)].tobytes()
    with open(path, 'wb') as f:
        f.write(b'>chr1 circular=true\n')
        f.write(seq)
        f.write(b'\n')


def build_workload(rank, world, n_reads_override=None):
    """Plans this rank's reads of the configs[1] workload. Returns (planner, ref, models, plans, read indices)."""
    from badread_b200 import simulate as S
    from badread_b200.__main__ import check_simulate_args, parse_args
    from badread_b200.error_model import ErrorModel
    from badread_b200.fragment_lengths import FragmentLengths
    from badread_b200.identities import Identities
    from badread_b200.qscore_model import QScoreModel
    tmp = os.path.join(tempfile.gettempdir(), f'badread_b200_bench_ref_{os.getpid()}.fasta')
    make_reference_fasta(tmp)
    args = parse_args(['simulate', '--reference', tmp, '--quantity', f'{DEPTH}x', '--error_model', 'nanopore2023',
                       '--qscore_model', 'nanopore2023', '--seed', str(SEED)])
    check_simulate_args(args)
    sink = io.StringIO()
    ref = S.Reference(args.reference, sink)
    os.unlink(tmp)
    fl = FragmentLengths(args.mean_frag_length, args.frag_length_stdev, sink)
    S.adjust_depths(ref, fl, args, np.random.RandomState(SEED))
    ids = Identities(args.mean_identity, args.identity_stdev, args.max_identity, sink)
    em, qm = ErrorModel(args.error_model, sink), QScoreModel(args.qscore_model, sink)
    planner = S.ReadPlanner(args, ref, fl, ids, SEED)
    target = S.get_target_size(ref.size, args.quantity)
    # number of reads of the N=1 job: plan until the error-free lengths reach the target (reads come out ~1% shorter
    # or longer than their fragments, so this is the configs[1] read count to within a fraction of a percent)
    plans, indices, total, i = [], [], 0, 0
    while (n_reads_override is None and total < target) or (n_reads_override is not None and len(plans) < n_reads_override):
        idx = rank + world * i
        p = planner.plan(idx)
        plans.append(p)
        indices.append(idx)
        total += sum(x.length for x in p[0])
        i += 1
    return planner, ref, (em, qm), plans, indices


def make_batch(planner, plans, indices):
    from badread_b200.engine import FragmentBatch
    batch = FragmentBatch()
    for p, idx in zip(plans, indices):
        planner.add_to_batch(batch, idx, p[0], p[2])
    return batch


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(',')]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx = float(parts[2])
                except ValueError:
                    continue
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), parts[5:9]):
                    if val.lower() == 'active':
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out['sm_mhz'] = float(np.median(sm))
            out['sm_max_mhz'] = mx
            out['samples'] = len(sm)
        out['reasons'] = sorted(reasons)
        return out


def cpu_port_run(planner, models, plans, indices, n_threads, n_reads=None):
    """Runs the CPU oracle port (Philox mode, pthreads over reads) over the first n_reads reads of the workload (all of
    them by default).  Returns (outputs, bases, seconds, description); outputs[i] = (seq, qual, matches, columns)."""
    from oracle import oracle as O
    orc = O.Oracle(*models)
    n = len(plans) if n_reads is None else min(len(plans), n_reads)
    frs = [planner.materialise(p[0]) for p in plans[:n]]
    ids = [p[2] for p in plans[:n]]
    t0 = time.perf_counter()
    outs, bases = orc.sequence_batch(frs, ids, SEED, indices[:n], n_threads=n_threads)
    dt = time.perf_counter() - t0
    what = 'the whole workload' if n == len(plans) else f'the first {n} of {len(plans)} reads of the workload'
    return outs, bases, dt, f'{what}: {n} reads, {bases} bases in {dt:.2f} s on {n_threads} threads'


def parity_check(res, outs):
    """GPU reads (BatchResult) against the oracle's for the same read indices: sequences, quality strings, alignment
    counts.  Returns the parity object of the JSON line."""
    bad = []
    for i, o in enumerate(outs):
        rec = res.records[i]
        if res.read(i) != (o[0], o[1]) or (rec.matches, rec.columns) != (o[2], o[3]):
            bad.append(i)
    return {'reads_checked': len(outs), 'bases_checked': int(sum(len(o[0]) for o in outs)), 'mismatches': len(bad),
            'first_mismatching_reads': bad[:8], 'against': 'oracle/badread_oracle.c (Philox mode), same read indices'}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', type=str, default='b200', choices=['b200', 'reference'])
    ap.add_argument('--reads', type=int, default=None, help='override the number of reads per rank (debugging)')
    ap.add_argument('--profile', action='store_true', help='skip the e2e and CPU legs (for runs under ncu)')
    a = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    n_cores = os.cpu_count() or 1
    config = {'workload': 'BASELINE.json configs[1]: 5 Mb synthetic circular ref (RandomState(1001)), 50x, '
                          'nanopore2023 error+qscore, default identity/length/adapters/glitches, seed 1',
              'sharding': f'read index mod {world}', 'cache': 'inputs (>=250 MB of fragments per step) exceed the 126 MB L2'}

    if a.impl == 'reference':
        if rank != 0:
            return 0
        planner, ref, models, plans, indices = build_workload(0, 1, n_reads_override=a.reads)
        values = []
        for s_ in range(a.warmup + a.steps):   # every step = the same reads the GPU arm processes per step
            _, bases, dt, desc = cpu_port_run(planner, models, plans, indices, n_cores)
            if s_ >= a.warmup:
                values.append((bases / dt / 1e9, bases, dt, desc))
        tot_b = sum(v[1] for v in values)
        tot_t = sum(v[2] for v in values)
        val = tot_b / tot_t / 1e9
        line = {'impl': 'reference', 'metric': 'simulated Gbases/sec', 'value': val, 'unit': 'Gbases/s', 'n_gpus': a.gpus,
                'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': tot_t / max(1, a.steps) * 1e3,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/int32 (+f64 identity estimate)',
                'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': val, 'unit': 'Gbases/s', 'cores': n_cores, 'kind': 'port', 'sample': values[-1][3]},
                'e2e': {'value': val, 'unit': 'Gbases/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        print(json.dumps(line), flush=True)
        return 0

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    from badread_b200.engine import Engine
    t_plan = time.perf_counter()
    planner, ref, models, plans, indices = build_workload(rank, world, n_reads_override=a.reads)
    batch = make_batch(planner, plans, indices)
    log(f'[rank {rank}] planned {len(plans)} reads in {time.perf_counter() - t_plan:.1f} s')
    eng = Engine(device=local_rank, seed=SEED)
    eng.upload_reference(ref.concat)
    eng.set_error_model(models[0])
    eng.set_qscore_model(models[1])
    ri, so, segs, lit, lit_len, ti = batch.arrays()
    h2d = ri.nbytes + so.nbytes + len(batch.seg_src) * 16 + lit_len + ti.nbytes

    def barrier():
        eng.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- device-resident timing: descriptors uploaded once, K x bb_batch_run
    eng.upload_batch(batch)
    sampler = ClockSampler(local_rank)
    sampler.start()   # nvidia-smi needs ~0.5 s to deliver its first sample: it runs from the warm-up steps (same load) on
    for _ in range(a.warmup):
        eng.run_batch()
    eng.synchronize()
    res, bases = eng.fetch_batch()
    launches0 = eng.launch_count()
    barrier()
    t0 = time.perf_counter()
    stage_acc = {}
    dev_ms = 0.0
    for _ in range(a.steps):
        eng.run_batch()
        total_ms, stages = eng.last_run_ms()  # waits for the step's last event; steps are serial anyway
        dev_ms += total_ms
        for k, v in stages.items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    barrier()
    elapsed = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    recs = res.records if False else None

    # ---- end to end through bb_sequence_batch: host descriptors in, host seq/qual out, every step
    bases_e2e, e2e_elapsed = bases, float('nan')
    if not a.profile:
        eng.sequence_batch(batch)
        barrier()
        t1 = time.perf_counter()
        for _ in range(a.steps):
            res, bases_e2e = eng.sequence_batch(batch)
        barrier()
        e2e_elapsed = time.perf_counter() - t1
    d2h = 2 * bases_e2e + len(plans) * 40

    tot_bases, max_elapsed, max_e2e = float(bases), elapsed, e2e_elapsed
    if dist is not None:
        import torch
        tb = torch.tensor([float(bases)], device='cuda', dtype=torch.float64)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        tm = torch.tensor([elapsed, e2e_elapsed], device='cuda', dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        tot_bases, max_elapsed, max_e2e = float(tb.item()), float(tm[0].item()), float(tm[1].item())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    value = tot_bases * a.steps / max_elapsed / 1e9
    e2e_value = tot_bases * a.steps / max_e2e / 1e9
    kernels = {k: v / a.steps for k, v in stage_acc.items() if k not in ('total', 'host_scan')}
    dom = max(kernels, key=kernels.get)
    peak, peak_src = peaks()
    # algorithmic bytes of the dominant stage per base: the final alignment reads the mutated
    # sequence and the fragment (2 B) and writes one alignment op (1 B); the error loop reads the fragment and
    # writes the slot state (1 + 4 B)
    alg_bytes_per_base = {'final_align': 3.0, 'error_loop': 5.0}.get(dom, 3.0)
    achieved = alg_bytes_per_base * bases / (kernels[dom] * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:  # DRAM bytes of the stage's kernels over one step, from the committed ncu pass (profiles/)
        with open(os.path.join(ROOT, 'profiles', 'r1c_traffic.json')) as f:
            tj = json.load(f)
        traffic = float(tj['per_stage'][dom]['dram_GB']) * 1e9
        traffic_src = 'profiles/r1c_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, one step)'
    except Exception:
        pass
    alu = None
    try:  # integer ALU-pipe accounting: warp instructions of one step (committed ncu pass) over the measured step time
        with open(os.path.join(ROOT, 'profiles', 'r1c_inst.json')) as f:
            ginst = float(json.load(f)['warp_inst_G_per_step'])
        sm_count, smsp = 148, 4
        clk_ghz = (clocks.get('sm_mhz') or 1965.0) / 1e3
        peak_alu = sm_count * smsp * clk_ghz * 0.5   # LOP3/IADD3/SHF: one warp instruction per 2 clocks per SMSP
        ach = ginst / (dev_ms / a.steps * 1e-3)
        alu = {'achieved': ach, 'peak': peak_alu, 'unit': 'G warp-inst/s', 'frac': ach / peak_alu,
               'warp_inst_G_per_step': ginst, 'source': 'profiles/r1c_inst.json (ncu smsp__inst_executed.sum)'}
    except Exception:
        pass
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src, 'kernel_ms': kernels[dom],
                'algorithmic_bytes_per_step': alg_bytes_per_base * bases, 'alu_pipe': alu,
                'note': 'stage = all kernels of the Hirschberg task pipeline (bb_k_node_warp<4> dominant); the path is '
                        'bit-vector DP bound by the integer ALU pipe (0.5 warp-inst/clk/SMSP), not by HBM: the HBM '
                        'fraction is small by construction, see DESIGN.md section 5 for the ALU-pipe accounting'}
    cpu_g, cpu_desc, parity = None, 'skipped (--profile)', None
    if not a.profile:
        outs, cpu_bases, cpu_dt, cpu_desc = cpu_port_run(planner, models, plans, indices, n_cores)
        cpu_g = cpu_bases / cpu_dt / 1e9
        parity = parity_check(res, outs)   # res: the reads fetched by the last end-to-end step
    line = {'metric': 'simulated Gbases/sec', 'value': value, 'unit': 'Gbases/s', 'n_gpus': world, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': max_elapsed / a.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'u8/int32 (+f64 identity estimate)', 'data': 'synthetic', 'config': config,
            'reads_per_step_rank0': len(plans), 'bases_per_step_rank0': bases,
            'device_ms_per_step_rank0': dev_ms / a.steps, 'stage_ms_rank0': {k: v / a.steps for k, v in stage_acc.items()},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'Gbases/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
            'gpu_launches': int(launches),
            'roofline': roofline,
            'cpu_baseline': {'value': cpu_g, 'unit': 'Gbases/s', 'cores': n_cores, 'kind': 'port', 'sample': cpu_desc},
            'parity': parity}
    print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()
    if parity is not None and parity['mismatches']:
        log(f'PARITY FAILURE: {parity}')
        return 3
    return 0


if __name__ == '__main__':
    sys.exit(main())
