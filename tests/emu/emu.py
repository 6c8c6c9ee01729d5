"""ctypes front end of the warp-emulated device aligner (tests/emu/emu_align.cpp). TEST INFRASTRUCTURE."""
import ctypes
import os
import pathlib
import subprocess

import numpy as np

HERE = pathlib.Path(os.path.dirname(os.path.realpath(__file__)))
LIB = HERE / 'libemu_align.so'


def build():
    csrc = HERE.parent.parent / 'badread_b200' / 'csrc'
    srcs = [HERE / 'emu_align.cpp', HERE / 'cuda_emu.h'] + sorted(csrc.glob('*.cuh'))
    if not LIB.is_file() or any(LIB.stat().st_mtime < s.stat().st_mtime for s in srcs):
        subprocess.run(['g++', '-O1', '-std=c++17', '-fPIC', '-shared', '-o', str(LIB), str(srcs[0])], check=True)
    return LIB


_lib = None


def align_path(query, target, k_upper=None, qabs_pad=0, maxl=16):
    """Device bb_align under the emulator -> (expanded ops string, distance)."""
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(LIB))
    q = query.encode('latin-1') if isinstance(query, str) else bytes(query)
    t = target.encode('latin-1') if isinstance(target, str) else bytes(target)
    n, m = len(q), len(t)
    ops = np.zeros(n, dtype=np.uint8)
    dcnt = np.zeros(n, dtype=np.uint32)
    out5 = np.zeros(5, dtype=np.int32)
    _lib.emu_align_path(q, n, t, m, max(n, m) if k_upper is None else k_upper, qabs_pad, maxl,
                        ops.ctypes.data_as(ctypes.c_void_p), dcnt.ctypes.data_as(ctypes.c_void_p),
                        out5.ctypes.data_as(ctypes.c_void_p))
    if out5[4]:
        raise RuntimeError(f'aligner error flags 0x{int(out5[4]):x}')
    sym = '=XI'
    parts = ['D' * int(out5[3])]
    for i in range(n):
        parts.append(sym[ops[i]])
        if dcnt[i]:
            parts.append('D' * int(dcnt[i]))
    s = ''.join(parts)
    assert len(s) == n + out5[1], (len(s), n, out5)
    assert s.count('=') == out5[0]
    return s, int(out5[2])


def lane_align(query, target, k_upper, qabs_pad=0, lw=4):
    """Lane-mode pass + traceback under the emulator -> (matches, dels, distance) or None if the band needs more
    than lw window words."""
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(LIB))
    q = query.encode('latin-1') if isinstance(query, str) else bytes(query)
    t = target.encode('latin-1') if isinstance(target, str) else bytes(target)
    out4 = np.zeros(4, dtype=np.int32)
    rc = _lib.emu_lane_align(q, len(q), t, len(t), k_upper, qabs_pad, lw, out4.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        return None
    if out4[3]:
        raise RuntimeError(f'lane aligner error flags 0x{int(out4[3]):x}')
    return int(out4[0]), int(out4[1]), int(out4[2])


def tasks_align(seq, frag, upper, quad=True, hist=1):
    """The level-synchronous alignment task pipeline under the emulator -> expanded ops string.
    quad: wide nodes by the 8-warp CTA kernel (bb_k_node_quad) instead of warp pairs (bb_k_node_pair).
    hist: lane leaves by the default build (global history + shared-memory staging ring) / 0: the checkpoint build."""
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(LIB))
    _lib.emu_set_quad(1 if quad else 0)
    _lib.emu_set_hist(int(hist))
    q = seq.encode('latin-1') if isinstance(seq, str) else bytes(seq)
    t = frag.encode('latin-1') if isinstance(frag, str) else bytes(frag)
    n, m = len(q), len(t)
    ops = np.zeros(n, dtype=np.uint8)
    dcnt = np.zeros(n, dtype=np.uint32)
    out5 = np.zeros(5, dtype=np.int32)
    _lib.emu_tasks_align(q, n, t, m, upper, ops.ctypes.data_as(ctypes.c_void_p), dcnt.ctypes.data_as(ctypes.c_void_p),
                         out5.ctypes.data_as(ctypes.c_void_p))
    if out5[4] or out5[2]:
        raise RuntimeError(f'task pipeline error flags 0x{int(out5[4]):x} overflow {int(out5[2])}')
    sym = '=XI'
    parts = ['D' * int(out5[3])]
    for i in range(n):
        parts.append(sym[ops[i]])
        if dcnt[i]:
            parts.append('D' * int(dcnt[i]))
    s = ''.join(parts)
    assert len(s) == n + out5[1], (len(s), n, out5)
    assert s.count('=') == out5[0], (s.count('='), out5)
    return s


def compare_passes(query, target, k):
    """Mismatch count between the two-column bit-plane pass of the lean node kernels (bb_band_pass_bp) and the
    single-column distance pass, both directions of a Hirschberg node side by side; None if the band does not fit the
    lean kernels."""
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(LIB))
    q = query.encode('latin-1') if isinstance(query, str) else bytes(query)
    t = target.encode('latin-1') if isinstance(target, str) else bytes(target)
    rc = _lib.emu_compare_passes(q, len(q), t, len(t), int(k))
    return None if rc < 0 else int(rc)


def compare_sm(query, target, k, L):
    """Mismatch count between the wide wavefront pass with the shared-memory match-word cache and the streaming
    variant (L = 8, 16 or 32 words per lane); None if the band does not fit."""
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(LIB))
    q = query.encode('latin-1') if isinstance(query, str) else bytes(query)
    t = target.encode('latin-1') if isinstance(target, str) else bytes(target)
    rc = _lib.emu_compare_sm(q, len(q), t, len(t), int(k), int(L))
    return None if rc < 0 else int(rc)


def window_lane(frag, changes, seed, read_index, lw=4, hist=0):
    """bb_k_window_lane<lw> (hist = 0: checkpoint build) or bb_k_window_lane_hist<lw> (hist = 1: the default build, 4
    columns per traceback tick; 2: two columns per tick) under the emulator for one read: changes = [(position, new string of <= 3 chars)] in the
    order they were applied -> [(matches, columns)] of the identity re-measurements after 25, 50, ... changes
    ((-1, -1): the window does not fit lw words and went to the fallback list)."""
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(LIB))
    f = frag.encode('latin-1') if isinstance(frag, str) else bytes(frag)
    pos = np.asarray([p for p, _ in changes], dtype=np.int32)
    enc = np.zeros(len(changes), dtype=np.uint32)
    for i, (_, sub) in enumerate(changes):
        b = sub.encode('latin-1')
        assert len(b) <= 3
        enc[i] = len(b) | sum(b[j] << (8 * (j + 1)) for j in range(len(b)))
    n_meas = len(changes) // 25
    out = np.full(2 * max(n_meas, 1), -7, dtype=np.int32)
    _lib.emu_set_hist(int(hist))
    flags = _lib.emu_window_lane(f, len(f), pos.ctypes.data_as(ctypes.c_void_p), enc.ctypes.data_as(ctypes.c_void_p),
                                 len(changes), ctypes.c_uint64(seed), ctypes.c_uint64(read_index), int(lw),
                                 out.ctypes.data_as(ctypes.c_void_p))
    if flags:
        raise RuntimeError(f'window kernel error flags 0x{int(flags):x}')
    return [(int(out[2 * a]), int(out[2 * a + 1])) for a in range(n_meas)]


def error_loop(fragment, target_identity, seed, read_index, error_model):
    """The error loop kernels (bb_k_mutate, bb_k_window_tasks, bb_k_window_lane_hist<4> / <8>, bb_k_window_warp,
    bb_k_replay) of one read under the emulator, round after round as bb_api.cu enqueues them, then bb_k_join (its read,
    match bitmap and distance bound are checked inside the harness).  error_model: a
    badread_b200.error_model.ErrorModel with tables.  Returns (joined read, stats dict)."""
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(LIB))
    t = error_model.to_device_tables()
    f = fragment.encode('latin-1') if isinstance(fragment, str) else bytes(fragment)
    k = int(t['k'])
    cap = 2 * (len(f) + 2 * k) + 64
    joined = np.zeros(cap, dtype=np.uint8)
    padded = np.zeros(len(f) + 2 * k, dtype=np.uint8)
    out8 = np.zeros(8, dtype=np.int32)
    arr = {name: np.ascontiguousarray(t[name]) for name in ('kmer_to_row', 'row_off', 'cum', 'flags', 'slots', 'pool')}
    _lib.emu_error_loop.restype = ctypes.c_int
    _lib.emu_error_loop.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_int32] + [ctypes.c_void_p] * 5 + [ctypes.c_void_p, ctypes.c_void_p,
                                                                                                 ctypes.c_int, ctypes.c_void_p]
    rounds = _lib.emu_error_loop(f, len(f), float(target_identity), seed, read_index, k,
                                 arr['kmer_to_row'].ctypes.data_as(ctypes.c_void_p), len(arr['row_off']) - 1,
                                 *(arr[n].ctypes.data_as(ctypes.c_void_p) for n in ('row_off', 'cum', 'flags', 'slots', 'pool')),
                                 out8.ctypes.data_as(ctypes.c_void_p), joined.ctypes.data_as(ctypes.c_void_p), cap,
                                 padded.ctypes.data_as(ctypes.c_void_p))
    if rounds < 0:
        raise RuntimeError(f'error loop under the emulator failed ({rounds})')
    if out8[7]:
        raise RuntimeError(f'error loop flags 0x{int(out8[7]):x}')
    stats = {'loop_count': int(out8[0]), 'change_count': int(out8[1]), 'n_alignments': int(out8[2]),
             'untrimmed_len': int(out8[3]), 'start_trim': int(out8[4]), 'end_trim': int(out8[5]), 'upper': int(out8[6]),
             'rounds': int(rounds), 'padded_fragment': bytes(padded).decode('latin-1')}
    return bytes(joined[:out8[3]]).decode('latin-1'), stats


def get_qscores(seq, frag, upper, qscore_model, seed, read_index):
    """get_qscores on the device code under the emulator: alignment task pipeline + bb_k_qscores_pair -> (quality
    string, '=' columns, alignment columns)."""
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(LIB))
    _lib.emu_set_quad(0)
    _lib.emu_set_hist(1)
    t = qscore_model.to_device_tables()
    q = seq.encode('latin-1') if isinstance(seq, str) else bytes(seq)
    f = frag.encode('latin-1') if isinstance(frag, str) else bytes(frag)
    qual = np.zeros(len(q), dtype=np.uint8)
    out5 = np.zeros(5, dtype=np.int32)
    arr = {name: np.ascontiguousarray(t[name]) for name in ('keys', 'row_off', 'scores', 'cum')}
    _lib.emu_get_qscores.restype = ctypes.c_int
    _lib.emu_get_qscores.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int32] + [ctypes.c_void_p] * 4 + [ctypes.c_uint64, ctypes.c_uint64,
                                                                               ctypes.c_void_p, ctypes.c_void_p]
    rc = _lib.emu_get_qscores(q, len(q), f, len(f), int(upper), int(t['kmer_size']), int(t['n_keys']),
                              *(arr[n].ctypes.data_as(ctypes.c_void_p) for n in ('keys', 'row_off', 'scores', 'cum')),
                              seed, read_index, qual.ctypes.data_as(ctypes.c_void_p), out5.ctypes.data_as(ctypes.c_void_p))
    if rc:
        raise RuntimeError(f'get_qscores under the emulator failed (flags 0x{int(out5[4]):x}, overflow {int(out5[2])})')
    return bytes(qual).decode('latin-1'), int(out5[0]), len(q) + int(out5[1])


def count_windows(which, flat, k, max_del=0, device=0, cap=None):
    """The counting kernels of the model builders (csrc/bb_models.cuh) under the emulator, in place of
    badread_b200.model_builders._count: same arguments, same return value."""
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(LIB))
    kind = 0 if which == 'kmers' else 1
    per_slot = 1 if kind == 0 else 94
    if cap is None:
        cap = 1 << 12
        while kind == 0 and cap < 2 * int(flat.ref_off[-1]) + 16:
            cap <<= 1
    ovf_cap = 1 << 16
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    _lib.emu_count_windows.restype = ctypes.c_int
    _lib.emu_count_windows.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int32] + [ctypes.c_void_p] * 9 + \
        [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p,
         ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]
    while True:
        keys = np.empty(cap, dtype=np.uint64); first = np.empty(cap, dtype=np.uint64)
        counts = np.empty(cap * per_slot, dtype=np.uint32)
        ovf = [np.empty(ovf_cap, dtype=np.int32) for _ in range(3)]
        overall = np.zeros(94, dtype=np.uint64)
        n, m = ctypes.c_int64(0), ctypes.c_int64(0)
        rc = _lib.emu_count_windows(kind, k, max_del, flat.n, p(flat.read), p(flat.qual), p(flat.read_off), p(flat.ref),
                                    p(flat.ref_off), p(flat.ops), p(flat.op_read0), p(flat.op_ref0), p(flat.ops_off), cap, p(keys),
                                    p(first), p(counts), ctypes.byref(n), p(overall), ovf_cap, p(ovf[0]), p(ovf[1]), p(ovf[2]),
                                    ctypes.byref(m))
        if rc == -4:          # table (or overflow list) full: the same retry as model_builders._count
            if m.value > ovf_cap:
                ovf_cap = int(m.value) + 16
            else:
                cap <<= 1
            continue
        if rc:
            raise RuntimeError(f'emu_count_windows failed ({rc})')
        return keys[:n.value], first[:n.value], counts[:n.value * per_slot].reshape(n.value, per_slot), overall, \
            [o[:m.value] for o in ovf]


def build_fragment(ref, literals, segments, k, kmer_to_row, seed, read_index):
    """bb_k_build_fragments (+ bb_k_compact on its output) for one read under the emulator.  segments: [(kind, src, len)]
    with kind 0 = reference slice, 1 = reverse complement of a reference slice, 2 = literal bytes.  Returns (padded
    fragment, k-mer row index per position, status: 0 = slots reset, bitmap and compaction right)."""
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(LIB))
    r = np.frombuffer(ref.encode('latin-1') if isinstance(ref, str) else bytes(ref), dtype=np.uint8)
    lit = np.frombuffer((literals.encode('latin-1') if isinstance(literals, str) else bytes(literals)) or b'\0', dtype=np.uint8)
    kind = np.asarray([s[0] for s in segments], dtype=np.int32)
    src = np.asarray([s[1] for s in segments], dtype=np.int64)
    ln = np.asarray([s[2] for s in segments], dtype=np.int32)
    n = int(ln.sum()) + 2 * k
    frag = np.zeros(n + 8, dtype=np.uint8)
    kidx = np.full(n + 8, -9, dtype=np.int32)
    k2r = np.ascontiguousarray(kmer_to_row, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    _lib.emu_build_fragment.restype = ctypes.c_int
    _lib.emu_build_fragment.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64,
                                                                ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
    status = _lib.emu_build_fragment(p(r), p(lit), p(kind), p(src), p(ln), len(segments), k, p(k2r), seed, read_index, p(frag),
                                     p(kidx))
    return bytes(frag[:n]).decode('latin-1'), kidx[:max(0, n - k + 1)].tolist(), int(status)
