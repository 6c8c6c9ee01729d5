#!/usr/bin/env python3
"""
compile_models.py - turns Badread model text files (error_models/*.gz, qscore_models/*.gz of a Badread checkout)
into the precompiled table archives badread_b200 ships under badread_b200/models/ (<name>.error.npz,
<name>.qscore.npz). The error-model archives hold the flat device tables (slot strings already aligned by
error_model.align_kmers' rule), so loading a built-in model does not repeat ~425k alignments.

usage: tools/compile_models.py /path/to/Badread/badread [names...]
"""
import io
import os
import pathlib
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.realpath(__file__)), '..'))
from badread_b200.error_model import BUILTIN_MODELS, MODEL_DIR, ErrorModel  # noqa: E402
from badread_b200.qscore_model import QScoreModel  # noqa: E402


def main():
    src = pathlib.Path(sys.argv[1])
    names = sys.argv[2:] or list(BUILTIN_MODELS)
    MODEL_DIR.mkdir(exist_ok=True)
    sink = io.StringIO()
    for name in names:
        em = ErrorModel(str(src / 'error_models' / f'{name}.gz'), output=sink)
        em.save_tables(MODEL_DIR / f'{name}.error.npz')
        qm = QScoreModel(str(src / 'qscore_models' / f'{name}.gz'), output=sink)
        qm.save_tables(MODEL_DIR / f'{name}.qscore.npz')
        print(name, 'k =', em.kmer_size, 'rows =', len(em._tables['row_off']) - 1, 'qscore keys =', len(qm.scores))


if __name__ == '__main__':
    main()
