"""
world_size-2 `gloo` test of the multi-GPU plumbing on CPU: ranks shard reads by index (rank g owns g, g+N, ...), plan
their own reads independently, and the only collectives are SUM (emitted bases / read counts) and MAX (time), exactly
what bench.py does over NCCL.  The union of the shards must be the unsharded read set.
"""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))


def _worker(rank, world, port, ref_path, out):
    sys.path.insert(0, ROOT)
    import io
    import numpy as np
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from badread_b200 import simulate as S
    from badread_b200.__main__ import check_simulate_args, parse_args
    from badread_b200.fragment_lengths import FragmentLengths
    from badread_b200.identities import Identities
    args = parse_args(['simulate', '--reference', ref_path, '--quantity', '3x', '--length', '2000,1000', '--seed', '9'])
    check_simulate_args(args)
    sink = io.StringIO()
    ref = S.Reference(args.reference, sink)
    fl = FragmentLengths(args.mean_frag_length, args.frag_length_stdev, sink)
    S.adjust_depths(ref, fl, args, np.random.RandomState(9))
    planner = S.ReadPlanner(args, ref, fl, Identities(95, 2.5, 99, sink), 9)
    mine = [planner.plan(i) for i in range(rank, 64, world)]
    bases = sum(sum(p.length for p in pl[0]) for pl in mine)
    t = torch.tensor([float(bases), float(len(mine))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    mx = torch.tensor([float(rank)], dtype=torch.float64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    names = [str(pl[3]) for pl in mine]
    gathered = [None] * world
    dist.all_gather_object(gathered, names)
    # the native planner shards the same way (first index = rank, stride = world) and plans the same reads
    from badread_b200.planner import NativePlanner
    nat = NativePlanner(args, ref, fl, Identities(95, 2.5, 99, sink), 9, n_threads=2)
    pb = nat.plan(rank, len(mine), stride=world)
    assert [pb.name_str(i) for i in range(len(pb))] == names
    assert [pb.info_str(i) for i in range(len(pb))] == [' '.join(pl[1]) for pl in mine]
    assert [float(x) for x in pb.target_identity] == [pl[2] for pl in mine]
    nat_bases = torch.tensor([float(pb.frag_bases())], dtype=torch.float64)
    dist.all_reduce(nat_bases, op=dist.ReduceOp.SUM)
    assert int(nat_bases.item()) == int(t[0].item())
    nat.close()
    if rank == 0:
        serial = [planner.plan(i) for i in range(64)]
        assert int(t[1].item()) == 64
        assert int(t[0].item()) == sum(sum(p.length for p in pl[0]) for pl in serial)
        assert int(mx.item()) == world - 1
        merged = [None] * 64
        for g, lst in enumerate(gathered):
            for j, nm in enumerate(lst):
                merged[g + world * j] = nm
        assert merged == [str(pl[3]) for pl in serial]
        open(out, 'w').write('ok')
    dist.destroy_process_group()


def test_sharded_planning_over_gloo(tmp_path):
    torch = pytest.importorskip('torch')
    import numpy as np
    import torch.multiprocessing as mp
    rs = np.random.RandomState(2)
    ref_path = tmp_path / 'ref.fasta'
    ref_path.write_text('>c circular=true\n' + bytes(np.frombuffer(b'ACGT', dtype=np.uint8)[rs.randint(0, 4, 30000)]).decode() + '\n')
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = tmp_path / 'result'
    mp.spawn(_worker, args=(2, port, str(ref_path), str(out)), nprocs=2, join=True)
    assert out.read_text() == 'ok'
