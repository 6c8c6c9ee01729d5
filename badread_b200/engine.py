"""
Engine - Python owner of one `bb_ctx` (one GPU): uploads the reference and the model tables to HBM once and runs
`sequence_fragment` for batches of fragment descriptors through the C ABI (include/badread_b200.h).
No PyTorch, no CPU path: construction fails when the library or a GPU is missing.
"""
import ctypes
import os

import numpy as np

from . import _lib
from ._lib import BB_SEG_LITERAL, BB_SEG_REF_FWD, BB_SEG_REF_REV, ReadResult, Segment


_RESULT_DTYPE = np.dtype([('out_off', np.int64), ('out_len', np.int32), ('frag_len', np.int32), ('matches', np.int32),
                          ('columns', np.int32), ('loop_count', np.int32), ('change_count', np.int32),
                          ('n_alignments', np.int32), ('flags', np.int32), ('loop_kcycles', np.int32),
                          ('align_kcycles', np.int32)])
assert _RESULT_DTYPE.itemsize == ctypes.sizeof(ReadResult)


class EngineError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class FragmentBatch(object):
    """Flat fragment descriptors for a batch of reads: per read a run of segments (reference slices on either
    strand and literal bytes), the read's global index and its target identity."""

    def __init__(self):
        self.read_index = []
        self.target_identity = []
        self.seg_off = [0]
        self.seg_src, self.seg_len, self.seg_kind = [], [], []
        self.literals = bytearray()

    def __len__(self):
        return len(self.read_index)

    def add_literal_segment(self, data):
        if isinstance(data, str):
            data = data.encode('latin-1')
        self.seg_src.append(len(self.literals))
        self.seg_len.append(len(data))
        self.seg_kind.append(BB_SEG_LITERAL)
        self.literals += data

    def add_ref_segment(self, src, length, reverse):
        self.seg_src.append(int(src))
        self.seg_len.append(int(length))
        self.seg_kind.append(BB_SEG_REF_REV if reverse else BB_SEG_REF_FWD)

    def end_read(self, read_index, target_identity):
        self.read_index.append(int(read_index))
        self.target_identity.append(float(target_identity))
        self.seg_off.append(len(self.seg_src))

    def add_literal_read(self, read_index, fragment, target_identity):
        self.add_literal_segment(fragment)
        self.end_read(read_index, target_identity)

    def frag_bases(self):
        return sum(self.seg_len)

    def arrays(self):
        """The flat descriptor arrays bb_batch_upload takes (built once per batch state)."""
        key = (len(self.read_index), len(self.seg_src), len(self.literals))
        cached = getattr(self, '_arrays', None)
        if cached is not None and cached[0] == key:
            return cached[1]
        out = self._build_arrays()
        self._arrays = (key, out)
        return out

    def _build_arrays(self):
        n_seg = len(self.seg_src)
        segs = (Segment * max(n_seg, 1))()
        seg_np = np.frombuffer(segs, dtype=np.dtype([('src', np.int64), ('len', np.int32), ('kind', np.int32)]))
        if n_seg:
            seg_np['src'][:n_seg] = self.seg_src
            seg_np['len'][:n_seg] = self.seg_len
            seg_np['kind'][:n_seg] = self.seg_kind
        lit = np.frombuffer(bytes(self.literals), dtype=np.uint8) if self.literals else np.zeros(1, dtype=np.uint8)
        return (np.asarray(self.read_index, dtype=np.uint64), np.asarray(self.seg_off, dtype=np.int32), segs,
                np.ascontiguousarray(lit), len(self.literals), np.asarray(self.target_identity, dtype=np.float64))


class BatchResult(object):
    def __init__(self, results, seq, qual, n):
        self.records = results
        self.seq = seq
        self.qual = qual
        self.n = n

    def read(self, i):
        r = self.records[i]
        s = bytes(self.seq[r.out_off:r.out_off + r.out_len]).decode('latin-1')
        q = bytes(self.qual[r.out_off:r.out_off + r.out_len]).decode('latin-1')
        return s, q

    def identity(self, i):
        r = self.records[i]
        return r.matches / r.columns if r.columns else 0.0

    def table(self):
        """The records as a numpy structured array (no copy)."""
        return np.frombuffer(self.records, dtype=_RESULT_DTYPE, count=self.n)

    def total_bases(self):
        return int(self.table()['out_len'].sum()) if self.n else 0


class Engine(object):

    def __init__(self, device=0, seed=0):
        self._lib = _lib.lib()
        self._ctx = ctypes.c_void_p()
        rc = self._lib.bb_create(ctypes.byref(self._ctx), int(device), ctypes.c_uint64(int(seed) & (2 ** 64 - 1)))
        if rc != 0:
            msg = self._lib.bb_last_error(None)
            self._ctx = None
            raise EngineError(f'bb_create failed ({rc}): {msg.decode() if msg else ""}')
        self.device = device
        self.seed = seed
        self.error_model = None
        self.qscore_model = None
        self._out_cap = 0
        self._seq_buf = self._qual_buf = None
        self._pinned = []

    def close(self):
        if getattr(self, '_ctx', None):
            self._free_out()
            self._lib.bb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.bb_last_error(self._ctx)
            raise EngineError(f'{what} failed ({rc}): {msg.decode() if msg else ""}')

    # ---- uploads
    def upload_reference(self, bases):
        arr = np.frombuffer(bases, dtype=np.uint8) if isinstance(bases, (bytes, bytearray)) else np.ascontiguousarray(bases, dtype=np.uint8)
        self._ref_keepalive = arr
        self._check(self._lib.bb_upload_reference(self._ctx, _ptr(arr), arr.size), 'bb_upload_reference')

    def set_error_model(self, error_model):
        t = error_model.to_device_tables()
        if t['type'] == 0:
            rc = self._lib.bb_upload_error_model(self._ctx, 1, 0, None, 0, 0, None, None, None, None, None, 0)
        else:
            rc = self._lib.bb_upload_error_model(self._ctx, t['k'], 1, _ptr(t['kmer_to_row']), t['kmer_to_row'].size,
                                                 len(t['row_off']) - 1, _ptr(t['row_off']), _ptr(t['cum']),
                                                 _ptr(t['flags']), _ptr(t['slots']), _ptr(t['pool']), t['pool'].size)
        self._check(rc, 'bb_upload_error_model')
        self.error_model = error_model

    def set_qscore_model(self, qscore_model):
        t = qscore_model.to_device_tables()
        rc = self._lib.bb_upload_qscore_model(self._ctx, t['kmer_size'], t['n_keys'], _ptr(t['keys']),
                                              _ptr(t['row_off']), _ptr(t['scores']), _ptr(t['cum']))
        self._check(rc, 'bb_upload_qscore_model')
        self.qscore_model = qscore_model

    # ---- batch
    def _ensure_out(self, cap):
        if cap > self._out_cap:
            cap = int(cap * 1.25) + 4096
            # earlier (smaller) buffers stay allocated until close(): BatchResults handed out before still view them
            bufs = []
            for _ in range(2):  # page-locked, so that the device-to-host copies run at link rate
                p = ctypes.c_void_p()
                if self._lib.bb_host_alloc(ctypes.byref(p), cap) != 0:
                    raise EngineError(f'bb_host_alloc({cap}) failed')
                self._pinned.append(p)
                bufs.append(np.ctypeslib.as_array((ctypes.c_uint8 * cap).from_address(p.value)))
            self._seq_buf, self._qual_buf = bufs
            self._out_cap = cap

    def _free_out(self):
        self._seq_buf = self._qual_buf = None
        self._out_cap = 0
        for p in self._pinned:
            self._lib.bb_host_free(p)
        self._pinned = []

    def upload_batch(self, batch):
        ri, so, segs, lit, lit_len, ti = batch.arrays()
        self._batch_keepalive = (ri, so, segs, lit, ti)
        self._n = len(batch)
        rc = self._lib.bb_batch_upload(self._ctx, self._n, _ptr(ri), _ptr(so), ctypes.cast(segs, ctypes.c_void_p),
                                       _ptr(lit), lit_len, _ptr(ti))
        self._check(rc, 'bb_batch_upload')

    def run_batch(self):
        self._check(self._lib.bb_batch_run(self._ctx), 'bb_batch_run')

    def synchronize(self):
        self._check(self._lib.bb_synchronize(self._ctx), 'bb_synchronize')

    def last_run_ms(self):
        total = ctypes.c_float(0)
        stages = (ctypes.c_float * _lib.BB_N_STAGES)()
        self._check(self._lib.bb_last_run_ms(self._ctx, ctypes.byref(total), stages), 'bb_last_run_ms')
        names = [self._lib.bb_stage_name(i).decode() for i in range(_lib.BB_N_STAGES)]
        return total.value, dict(zip(names, [float(x) for x in stages]))

    def launch_count(self):
        return int(self._lib.bb_launch_count(self._ctx))

    def trace_dump(self, path):
        """Timeline of the last run (needs BADREAD_B200_TRACE=1 in the environment before the Engine is created)."""
        self._check(self._lib.bb_trace_dump(self._ctx, str(path).encode()), 'bb_trace_dump')

    def fetch_batch(self):
        n = self._n
        results = (ReadResult * n)()
        total = ctypes.c_int64(0)
        rc = self._lib.bb_fetch_last_batch(self._ctx, results, _ptr(self._seq_buf) if self._seq_buf is not None else None,
                                           _ptr(self._qual_buf) if self._qual_buf is not None else None,
                                           self._out_cap, ctypes.byref(total))
        if rc == _lib.BB_ERR_CAPACITY:
            self._ensure_out(total.value)
            rc = self._lib.bb_fetch_last_batch(self._ctx, results, _ptr(self._seq_buf), _ptr(self._qual_buf),
                                               self._out_cap, ctypes.byref(total))
        self._check(rc, 'bb_fetch_last_batch')
        return BatchResult(results, self._seq_buf, self._qual_buf, n), int(total.value)

    def sequence_batch(self, batch):
        """bb_sequence_batch (upload + run + fetch in one call: with several workers each block is copied out while
        the other workers still compute); returns (BatchResult, total_bases)."""
        ri, so, segs, lit, lit_len, ti = batch.arrays()
        self._batch_keepalive = (ri, so, segs, lit, ti)
        self._n = n = len(batch)
        if self._seq_buf is None:
            self._ensure_out(int(1.1 * batch.frag_bases()) + 4096)
        results = (ReadResult * n)()
        total = ctypes.c_int64(0)
        rc = self._lib.bb_sequence_batch(self._ctx, n, _ptr(ri), _ptr(so), ctypes.cast(segs, ctypes.c_void_p), _ptr(lit),
                                         lit_len, _ptr(ti), results, _ptr(self._seq_buf), _ptr(self._qual_buf),
                                         self._out_cap, ctypes.byref(total))
        if rc == _lib.BB_ERR_CAPACITY:
            self._ensure_out(total.value)
            rc = self._lib.bb_fetch_last_batch(self._ctx, results, _ptr(self._seq_buf), _ptr(self._qual_buf),
                                               self._out_cap, ctypes.byref(total))
        self._check(rc, 'bb_sequence_batch')
        return BatchResult(results, self._seq_buf, self._qual_buf, n), int(total.value)

    # ---- the one collective: SUM of emitted bases over the GPUs (stop condition, simulate.py:63)
    def comm_init_rank(self, unique_id, rank, world):
        """One process per GPU: joins the NCCL communicator identified by `unique_id` (128 bytes from
        `comm_unique_id()` on rank 0, shipped to the other ranks by the host program)."""
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self._check(self._lib.bb_comm_init_rank(self._ctx, buf, int(rank), int(world)), 'bb_comm_init_rank')

    def allreduce_bases(self, local):
        total = ctypes.c_int64(0)
        self._check(self._lib.bb_allreduce_bases(self._ctx, int(local), ctypes.byref(total)), 'bb_allreduce_bases')
        return int(total.value)

    # ---- single pair helpers
    def get_qscores(self, seq, frag, read_index=0):
        s = np.frombuffer(seq.encode('latin-1'), dtype=np.uint8)
        f = np.frombuffer(frag.encode('latin-1'), dtype=np.uint8)
        qual = np.empty(len(s), dtype=np.uint8)
        m, c = ctypes.c_int32(0), ctypes.c_int32(0)
        rc = self._lib.bb_get_qscores(self._ctx, ctypes.c_uint64(read_index), _ptr(s), len(s), _ptr(f), len(f), _ptr(qual),
                                      ctypes.byref(m), ctypes.byref(c))
        self._check(rc, 'bb_get_qscores')
        return bytes(qual).decode('latin-1'), m.value, c.value

    def align_path(self, query, target):
        """edlib.align(query, target, task='path') on the GPU -> (expanded ops string, edit distance)."""
        q = np.frombuffer(query.encode('latin-1') if isinstance(query, str) else bytes(query), dtype=np.uint8)
        t = np.frombuffer(target.encode('latin-1') if isinstance(target, str) else bytes(target), dtype=np.uint8)
        ops = np.empty(len(q) + len(t) + 16, dtype=np.uint8)
        n_ops, dist = ctypes.c_int64(0), ctypes.c_int32(0)
        rc = self._lib.bb_align_path(self._ctx, _ptr(q), len(q), _ptr(t), len(t), _ptr(ops), ops.size,
                                     ctypes.byref(n_ops), ctypes.byref(dist))
        self._check(rc, 'bb_align_path')
        return bytes(ops[:n_ops.value]).decode('ascii'), dist.value


def nccl_available():
    return bool(_lib.lib().bb_nccl_available())


def comm_unique_id():
    """128 bytes identifying a new NCCL communicator (ncclGetUniqueId)."""
    buf = ctypes.create_string_buffer(128)
    rc = _lib.lib().bb_comm_unique_id(buf)
    if rc != 0:
        raise EngineError(f'bb_comm_unique_id failed ({rc}): NCCL not available')
    return buf.raw


def comm_init_all(engines):
    """One process, several GPUs: one communicator over the engines' devices."""
    arr = (ctypes.c_void_p * len(engines))(*[e._ctx for e in engines])
    rc = _lib.lib().bb_comm_init_all(arr, len(engines))
    if rc != 0:
        msg = _lib.lib().bb_last_error(engines[0]._ctx)
        raise EngineError(f'bb_comm_init_all failed ({rc}): {msg.decode() if msg else ""}')


def allreduce_bases_all(engines, local):
    """SUM of the engines' emitted-base counts through one NCCL group call."""
    arr = (ctypes.c_void_p * len(engines))(*[e._ctx for e in engines])
    loc = (ctypes.c_int64 * len(engines))(*[int(x) for x in local])
    total = ctypes.c_int64(0)
    rc = _lib.lib().bb_allreduce_bases_all(arr, len(engines), loc, ctypes.byref(total))
    if rc != 0:
        msg = _lib.lib().bb_last_error(engines[0]._ctx)
        raise EngineError(f'bb_allreduce_bases_all failed ({rc}): {msg.decode() if msg else ""}')
    return int(total.value)


# ---- module-level default engine for the single-read convenience functions ---------------------------
_default = {'engine': None, 'seed': None, 'next_read': 0}


def set_seed(seed):
    """Seed of the per-read Philox streams used by the module-level convenience functions."""
    _default['seed'] = seed
    _default['next_read'] = 0
    if _default['engine'] is not None:
        _default['engine'].close()
        _default['engine'] = None


def default_engine(error_model=None, qscore_model=None):
    if _default['engine'] is None:
        seed = _default['seed']
        if seed is None:
            seed = int.from_bytes(os.urandom(8), 'little')
            _default['seed'] = seed
        _default['engine'] = Engine(device=int(os.environ.get('BADREAD_B200_DEVICE', '0')), seed=seed)
    eng = _default['engine']
    if error_model is not None and eng.error_model is not error_model:
        eng.set_error_model(error_model)
    if qscore_model is not None and eng.qscore_model is not qscore_model:
        eng.set_qscore_model(qscore_model)
    return eng


def next_read_index():
    i = _default['next_read']
    _default['next_read'] = i + 1
    return i
