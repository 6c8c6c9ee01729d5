"""
ctypes binding of libbadread_b200.so (the C ABI declared in include/badread_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C badread_b200/csrc`. There is no fallback:
if the shared object is missing this module raises, and if no GPU is usable `bb_create` fails.
"""
import ctypes
import os
import pathlib

_HERE = pathlib.Path(os.path.dirname(os.path.realpath(__file__)))
LIB_PATH = _HERE / 'libbadread_b200.so'

BB_OK = 0
BB_ERR_CUDA, BB_ERR_ARG, BB_ERR_STATE, BB_ERR_CAPACITY, BB_ERR_INTERNAL = -1, -2, -3, -4, -5
BB_SEG_REF_FWD, BB_SEG_REF_REV, BB_SEG_LITERAL = 0, 1, 2
BB_N_STAGES = 9


class Segment(ctypes.Structure):
    _fields_ = [('src', ctypes.c_int64), ('len', ctypes.c_int32), ('kind', ctypes.c_int32)]


class ReadResult(ctypes.Structure):
    _fields_ = [('out_off', ctypes.c_int64), ('out_len', ctypes.c_int32), ('frag_len', ctypes.c_int32),
                ('matches', ctypes.c_int32), ('columns', ctypes.c_int32), ('loop_count', ctypes.c_int32),
                ('change_count', ctypes.c_int32), ('n_alignments', ctypes.c_int32), ('flags', ctypes.c_int32),
                ('loop_kcycles', ctypes.c_int32), ('align_kcycles', ctypes.c_int32)]


class LibraryMissing(RuntimeError):
    pass


_lib = None


def lib():
    """Loads the shared library once. Raises LibraryMissing if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.is_file():
        raise LibraryMissing(f'{LIB_PATH} not found - build it with `python -c "import __graft_entry__ as g; '
                             f'g.build()"` or `make -C badread_b200/csrc` (there is no CPU fallback)')
    L = ctypes.CDLL(str(LIB_PATH))
    c = ctypes
    vp, i32, i64, u64, dbl = c.c_void_p, c.c_int32, c.c_int64, c.c_uint64, c.c_double
    P = c.POINTER
    sigs = {
        'bb_create': (c.c_int, [P(vp), c.c_int, u64]),
        'bb_destroy': (c.c_int, [vp]),
        'bb_last_error': (c.c_char_p, [vp]),
        'bb_version': (c.c_char_p, []),
        'bb_upload_reference': (c.c_int, [vp, vp, i64]),
        'bb_upload_error_model': (c.c_int, [vp, c.c_int, c.c_int, vp, i64, i32, vp, vp, vp, vp, vp, i64]),
        'bb_upload_qscore_model': (c.c_int, [vp, c.c_int, i32, vp, vp, vp, vp]),
        'bb_sequence_batch': (c.c_int, [vp, i32, vp, vp, vp, vp, i64, vp, vp, vp, vp, i64, P(i64)]),
        'bb_fetch_last_batch': (c.c_int, [vp, vp, vp, vp, i64, P(i64)]),
        'bb_batch_upload': (c.c_int, [vp, i32, vp, vp, vp, vp, i64, vp]),
        'bb_batch_run': (c.c_int, [vp]),
        'bb_synchronize': (c.c_int, [vp]),
        'bb_host_alloc': (c.c_int, [P(vp), i64]),
        'bb_host_free': (c.c_int, [vp]),
        'bb_last_run_ms': (c.c_int, [vp, P(c.c_float), P(c.c_float)]),
        'bb_stage_name': (c.c_char_p, [c.c_int]),
        'bb_launch_count': (i64, [vp]),
        'bb_get_qscores': (c.c_int, [vp, u64, vp, i32, vp, i32, vp, P(i32), P(i32)]),
        'bb_align_path': (c.c_int, [vp, vp, i32, vp, i32, vp, i64, P(i64), P(i32)]),
        'bb_host_align_kmers': (c.c_int, [c.c_int, i32, vp, vp, vp, vp, vp, vp, i64, P(i64)]),
        'bb_host_align_path': (c.c_int, [vp, i32, vp, i32, vp, i64, P(i64), P(i32)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = ['bb_create', 'bb_destroy', 'bb_last_error', 'bb_version', 'bb_upload_reference',
                    'bb_upload_error_model', 'bb_upload_qscore_model', 'bb_sequence_batch',
                    'bb_fetch_last_batch', 'bb_batch_upload', 'bb_batch_run', 'bb_synchronize', 'bb_host_alloc', 'bb_host_free',
                    'bb_last_run_ms', 'bb_stage_name', 'bb_launch_count', 'bb_get_qscores', 'bb_align_path',
                    'bb_host_align_kmers', 'bb_host_align_path']
