// bb_lane.cuh — lane-level banded alignment for narrow bands (one problem per THREAD).
//
// The identity re-measurements of the error loop (simulate.py:325-346) align a 1000-base window against its
// mutated copy: the Ukkonen band is only a few dozen rows wide, so a warp-wide wavefront would keep two or three
// lanes busy.  Here every lane owns a complete problem: it keeps a window of LW 32-row words that follows the
// band down the diagonal (shifted by one word every 32 columns) and advances one column per iteration with a
// single 32*LW-bit Myers step (one add carry chain, no shuffles).  Thirty-two reads progress per warp.
// Same path semantics as bb_align.cuh (edlib's traceback rule); the history layout is per lane:
// hist[c * LW + x] = (Pv, PhRaw) of window word x at column c, wtab[c] = the window's top word at column c.
#pragma once
#include <cstdint>

#include "bb_align.cuh"

// Window words needed for a band.  The window top follows the band lazily: lanes slide their windows only at
// warp-synchronous points (every 32 steps), so the top may lag one word behind max(0, (c - a) >> 5); rows up to
// c + b must still fit: ((a + b) >> 5) + 3 words.
__device__ __forceinline__ int bb_lane_words(int a, int b) { return ((a + b) >> 5) + 3; }


