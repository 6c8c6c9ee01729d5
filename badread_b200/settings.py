"""
Constants of the hot path, mirroring /root/reference/badread/settings.py:24-51 (same names, same values).
ALIGNMENT_INTERVAL / ALIGNMENT_SIZE are compiled into the kernels (csrc/bb_kernels.cuh); they are listed here
because the reference exposes them as module attributes.

Derived from Badread (Copyright 2018 Ryan Wick, rrwick@gmail.com, https://github.com/rrwick/Badread), which is free
software under the GNU General Public License version 3 or later; this file mirrors the named parts of the
reference's interface and is distributed under the same licence (see LICENSE and NOTICE at the repository root).
"""
ALIGNMENT_INTERVAL = 25
ALIGNMENT_SIZE = 1000

MIN_MEAN_READ_LENGTH = 100
MIN_MEAN_READ_IDENTITY = 50
MIN_MEAN_READ_QSCORE = 5

RANDOM_QSCORE_MIN = 1
RANDOM_QSCORE_MAX = 20

IDEAL_QSCORE_RANK_1_MIN, IDEAL_QSCORE_RANK_1_MAX = 1, 3
IDEAL_QSCORE_RANK_2_MIN, IDEAL_QSCORE_RANK_2_MAX = 4, 7
IDEAL_QSCORE_RANK_3_MIN, IDEAL_QSCORE_RANK_3_MAX = 8, 20
IDEAL_QSCORE_RANK_4_MIN, IDEAL_QSCORE_RANK_4_MAX = 21, 30
IDEAL_QSCORE_RANK_5_MIN, IDEAL_QSCORE_RANK_5_MAX = 31, 40
IDEAL_QSCORE_RANK_6_MIN, IDEAL_QSCORE_RANK_6_MAX = 41, 50

CHIMERA_START_ADAPTER_CHANCE = 0.25
CHIMERA_END_ADAPTER_CHANCE = 0.25
