"""
model_builders.py - `badread error_model` and `badread qscore_model` with the counting on the GPU.

Mirrors badread/error_model.py:31-83 (make_error_model), badread/qscore_model.py:78-175 (make_qscore_model,
print_qscore_fractions) and badread/alignment.py:23-100 (PAF records, best alignment per read): same arguments, same
messages, byte-identical model files (tests/test_model_builders.py compares with outputs of the unmodified reference).
The host parses the three input files and flattens the chosen alignments; libbadread_b200.so counts the windows
(csrc/bb_tu_models.cu: one CTA per alignment, one thread per window, 64-bit keys in an open-addressing table with the
first occurrence of every key); the host sorts and prints.  Windows whose content does not fit a key come back in an
overflow list and are evaluated here, exactly, from the same flat arrays.
"""
import collections
import ctypes
import re
import sys

import numpy as np

from . import _lib
from .misc import float_to_str, get_open_func, load_fasta, reverse_complement

_CIGAR_RUN = re.compile(r'(\d+)([A-Za-z=])')
_OP_CODE = {'M': 0, 'I': 1, 'D': 2}
_SYM = '=XID'
N_Q = 94


# ---------------------------------------------------------------------------------------------------- inputs
def load_fastq(filename, output=sys.stderr, dot_interval=1000):
    """misc.load_fastq (misc.py:97-119): {name: (upper-case sequence, qualities)}; name = first token of the header."""
    reads = {}
    print('Loading reads', end='', file=output, flush=True)
    with get_open_func(filename)(filename, 'rb') as handle:
        first = handle.read(1)
        if first != b'@':
            sys.exit('Error: {} is not FASTQ format'.format(filename))
        handle.seek(0)
        n = 0
        for line in handle:
            line = line.strip()
            if not line.startswith(b'@'):
                continue
            name = line[1:].split()[0].decode()
            seq = next(handle).strip().upper().decode()
            next(handle)
            qual = next(handle).strip().decode()
            reads[name] = (seq, qual)
            n += 1
            if n % dot_interval == 0:
                print('.', end='', file=output, flush=True)
    print('', file=output, flush=True)
    return reads


class Alignment(object):
    """One PAF line (alignment.py:23-76).  `runs` = the CIGAR runs in READ orientation (reversed for '-' strand hits)."""

    def __init__(self, paf_line):
        f = paf_line.strip().split('\t')
        if len(f) < 11:
            sys.exit('Error: alignment file does not seem to be in PAF format')
        self.read_name, self.read_start, self.read_end, self.strand = f[0], int(f[2]), int(f[3]), f[4]
        self.ref_name, self.ref_start, self.ref_end = f[5], int(f[7]), int(f[8])
        self.matching_bases, self.num_bases = int(f[9]), int(f[10])
        self.percent_identity = 100.0 * self.matching_bases / self.num_bases
        self.cigar, self.alignment_score = None, None
        for part in f:
            if part.startswith('cg:Z:'):
                self.cigar = part[5:]
            if part.startswith('AS:i:'):
                self.alignment_score = int(part[5:])
        if self.cigar is None:
            sys.exit('Error: no CIGAR string found')
        if self.alignment_score is None:
            sys.exit('Error: no alignment score')
        self.runs = [(int(n), t) for n, t in _CIGAR_RUN.findall(self.cigar)]
        self.max_indel = max([n for n, t in self.runs if t in 'ID'], default=0)
        if self.strand == '-':
            self.runs.reverse()

    def __repr__(self):
        return '%s:%d-%d(%s),%s:%d-%d(%.3f%%)' % (self.read_name, self.read_start, self.read_end, self.strand,
                                                   self.ref_name, self.ref_start, self.ref_end, self.percent_identity)


def load_alignments(filename, max_alignments=None, output=sys.stderr, dot_interval=1000):
    """alignment.py:79-105: the highest-scoring alignment of every read (the last one among equals), kept if it has
    more than 100 columns and more than 80 % identity; reads in order of first appearance."""
    print('Loading alignments', end='', file=output, flush=True)
    per_read = collections.OrderedDict()
    with get_open_func(filename)(filename, 'rt') as paf:
        for n, line in enumerate(paf, start=1):
            a = Alignment(line)
            per_read.setdefault(a.read_name, []).append(a)
            if n % dot_interval == 0:
                print('.', end='', file=output, flush=True)
            if n == max_alignments:
                break
    print('', file=output, flush=True)
    print('Choosing best alignment per read', end='', file=output, flush=True)
    chosen = []
    for alns in per_read.values():
        best = alns[0]
        for a in alns[1:]:
            if a.alignment_score >= best.alignment_score:
                best = a
        if best.num_bases > 100 and best.percent_identity > 80.0:
            chosen.append(best)
            if len(chosen) % dot_interval == 0:
                print('.', end='', file=output, flush=True)
    print('', file=output, flush=True)
    return chosen


class FlatAlignments(object):
    """The chosen alignments as the flat arrays bb_count_* take: per alignment the aligned slice of the read (+ its
    qualities), the aligned slice of the reference on the read's strand, and the CIGAR runs in read orientation with the
    offsets they start at."""

    def __init__(self, alignments, reads, refs, output, dot_interval):
        read_parts, qual_parts, ref_parts, ops, p0, r0 = [], [], [], [], [], []
        self.read_off, self.ref_off, self.ops_off = [0], [0], [0]
        print('Processing alignments', end='', file=output, flush=True)
        for n, a in enumerate(alignments, start=1):
            if a.read_name not in reads:
                sys.exit(f'\nError: could not find read {a.read_name}\nare you sure your read file and alignment file match?')
            if a.ref_name not in refs:
                sys.exit(f'\nError: could not find reference {a.ref_name}\nare you sure your reference file and '
                         f'alignment file match?')
            seq, qual = reads[a.read_name]
            read_seq, read_qual = seq[a.read_start:a.read_end], qual[a.read_start:a.read_end]
            ref_seq = refs[a.ref_name][a.ref_start:a.ref_end]
            if a.strand == '-':
                ref_seq = reverse_complement(ref_seq)
            rp = fp = 0
            for count, kind in a.runs:
                if kind not in _OP_CODE:
                    continue        # (alignment.align_sequences ignores every other CIGAR letter)
                ops.append((count << 2) | _OP_CODE[kind]); p0.append(rp); r0.append(fp)
                if kind != 'D':
                    rp += count
                if kind != 'I':
                    fp += count
            # the CIGAR may cover less than the slices (or more: the reference's slicing silently truncates, so do we)
            read_parts.append(read_seq[:rp].ljust(rp, '\0')); qual_parts.append(read_qual[:rp].ljust(rp, '\0'))
            ref_parts.append(ref_seq[:fp].ljust(fp, '\0'))
            self.read_off.append(self.read_off[-1] + rp)
            self.ref_off.append(self.ref_off[-1] + fp)
            self.ops_off.append(len(ops))
            if n % dot_interval == 0:
                print('.', end='', file=output, flush=True)
        print('', file=output, flush=True)
        self.n = len(alignments)
        self.read = np.frombuffer(''.join(read_parts).encode('latin-1') or b'\0', dtype=np.uint8)
        self.qual = np.frombuffer(''.join(qual_parts).encode('latin-1') or b'\0', dtype=np.uint8)
        self.ref = np.frombuffer(''.join(ref_parts).encode('latin-1') or b'\0', dtype=np.uint8)
        self.read_off = np.asarray(self.read_off, dtype=np.int64)
        self.ref_off = np.asarray(self.ref_off, dtype=np.int64)
        self.ops_off = np.asarray(self.ops_off, dtype=np.int64)
        self.ops = np.asarray(ops or [0], dtype=np.uint32)
        self.op_read0 = np.asarray(p0 or [0], dtype=np.int32)
        self.op_ref0 = np.asarray(r0 or [0], dtype=np.int32)

    # exact host evaluation of single windows (the overflow list)
    def columns(self, a):
        """Per read base of alignment a: symbol and the number of 'D' columns behind it; per reference base: the read
        offset at its column and whether that column holds a read base."""
        lo, hi = int(self.ops_off[a]), int(self.ops_off[a + 1])
        read = self.read[self.read_off[a]:self.read_off[a + 1]]
        ref = self.ref[self.ref_off[a]:self.ref_off[a + 1]]
        sym = np.zeros(len(read), dtype=np.uint8)
        dcount = np.zeros(len(read), dtype=np.int64)
        rp_at = np.zeros(len(ref), dtype=np.int64)
        is_m = np.zeros(len(ref), dtype=bool)
        lead = 0
        for o in range(lo, hi):
            count, kind, p, r = int(self.ops[o]) >> 2, int(self.ops[o]) & 3, int(self.op_read0[o]), int(self.op_ref0[o])
            if kind == 0:
                sym[p:p + count] = (read[p:p + count] != ref[r:r + count]).astype(np.uint8)
                rp_at[r:r + count] = np.arange(p, p + count); is_m[r:r + count] = True
            elif kind == 1:
                sym[p:p + count] = 2
            else:
                rp_at[r:r + count] = p
                if p > 0:
                    dcount[p - 1] += count
                else:
                    lead += count
        return read, ref, sym, dcount, rp_at, is_m, lead


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _count(which, flat, k, max_del=0, device=0):
    """Runs bb_count_kmer_alternatives / bb_count_cigar_qscores, growing the table until it fits."""
    L = _lib.lib()
    per_slot = 1 if which == 'kmers' else N_Q
    cap = 1 << 18       # slots; doubled until the distinct keys fit (k-mer pairs: at most one per window)
    while which == 'kmers' and cap < min(2 * int(flat.ref_off[-1]) + 16, 1 << 22):
        cap <<= 1
    ovf_cap = 1 << 16
    while True:
        keys = np.empty(cap, dtype=np.uint64); first = np.empty(cap, dtype=np.uint64)
        counts = np.empty(cap * per_slot, dtype=np.uint32)
        ovf = [np.empty(ovf_cap, dtype=np.int32) for _ in range(3)]
        overall = np.zeros(N_Q, dtype=np.uint64)
        n_entries, n_ovf = ctypes.c_int64(0), ctypes.c_int64(0)
        common = [_ptr(flat.ref), _ptr(flat.ref_off), _ptr(flat.ops), _ptr(flat.op_read0), _ptr(flat.op_ref0), _ptr(flat.ops_off),
                  cap, _ptr(keys), _ptr(first), _ptr(counts), ctypes.byref(n_entries)]
        tail = [ovf_cap, _ptr(ovf[0]), _ptr(ovf[1]), _ptr(ovf[2]), ctypes.byref(n_ovf)]
        if which == 'kmers':
            rc = L.bb_count_kmer_alternatives(device, k, flat.n, _ptr(flat.read), _ptr(flat.read_off), *common, *tail)
        else:
            rc = L.bb_count_cigar_qscores(device, k, max_del, flat.n, _ptr(flat.read), _ptr(flat.qual), _ptr(flat.read_off),
                                          *common, _ptr(overall), *tail)
        if rc == _lib.BB_ERR_CAPACITY:
            if n_ovf.value > ovf_cap:
                ovf_cap = int(n_ovf.value) + 16
            else:
                cap <<= 1
            continue
        if rc != _lib.BB_OK:
            raise RuntimeError('model builder: ' + L.bb_model_error().decode(errors='replace'))
        n, m = int(n_entries.value), int(n_ovf.value)
        return keys[:n], first[:n], counts[:n * per_slot].reshape(n, per_slot), overall, [o[:m] for o in ovf]


# ---------------------------------------------------------------------------------------------------- error model
def make_error_model(args, output=sys.stderr, dot_interval=1000):
    """error_model.py:31-83."""
    refs = load_fasta(args.reference)[0]
    reads = load_fastq(args.reads, output=output)
    alignments = load_alignments(args.alignment, args.max_alignments, output=output)
    if len(alignments) == 0:
        sys.exit('Error: no usable alignments')
    k = args.k_size
    flat = FlatAlignments(alignments, reads, refs, output, dot_interval)
    keys, first, counts, _, ovf = _count('kmers', flat, k)
    counts = counts[:, 0].astype(np.int64)
    shift_ref, shift_len = np.uint64(64 - 2 * k), np.uint64(58 - 2 * k)
    # read k-mers too long for a key (the overflow list): counted here, exactly; {reference k-mer: {read k-mer: [count, first]}}
    long_alts, cache = collections.defaultdict(dict), {}
    for a, r, _ in zip(*(o.tolist() for o in ovf)):
        if a not in cache:
            cache[a] = flat.columns(a)
        read, ref, _, _, rp_at, is_m, _ = cache[a]
        p_lo = 0 if r == 0 else int(rp_at[r])
        p_hi = int(rp_at[r + k - 1]) + int(is_m[r + k - 1])
        read_kmer = bytes(read[p_lo:p_hi]).decode('latin-1')
        if set(read_kmer) <= set('ACGT'):
            code = 0
            for c in bytes(ref[r:r + k]):
                code = code * 4 + b'ACGT'.index(c)
            entry = long_alts[code].setdefault(read_kmer, [0, (a << 32) | r])
            entry[0] += 1
            entry[1] = min(entry[1], (a << 32) | r)
    # per reference k-mer: total, the count of the unchanged k-mer, and the alternatives by (count, first occurrence) -
    # the order of the reference's stable sort by fraction over its insertion-ordered dict
    refcode = (keys >> shift_ref).astype(np.int64)
    totals = np.bincount(refcode, weights=counts, minlength=4 ** k).astype(np.int64)
    for code, alts in long_alts.items():
        totals[code] += sum(c for c, _ in alts.values())
    same = np.zeros(4 ** k, dtype=np.uint64)        # the read part of the key of an unchanged k-mer: base j at bits 2j
    codes = np.arange(4 ** k, dtype=np.uint64)
    for j in range(k):
        same |= ((codes >> np.uint64(2 * (k - 1 - j))) & np.uint64(3)) << np.uint64(2 * j)
    identity_key = (codes << shift_ref) | (np.uint64(k) << shift_len) | same
    is_identity = keys == identity_key[refcode]
    unchanged = np.zeros(4 ** k, dtype=np.int64)
    unchanged[refcode[is_identity]] = counts[is_identity]
    alt = np.flatnonzero(~is_identity)
    order = alt[np.lexsort((first[alt], -counts[alt], refcode[alt]))]
    group = refcode[order]
    starts = np.flatnonzero(np.concatenate([[True], group[1:] != group[:-1]])) if order.size else np.zeros(0, dtype=np.int64)
    rank = np.arange(order.size) - np.repeat(starts, np.diff(np.concatenate([starts, [order.size]])))
    # (a reference k-mer with alternatives in the overflow list keeps all of its entries: they are merged below)
    crowded = np.zeros(4 ** k, dtype=bool)
    crowded[list(long_alts)] = True
    kept = order[(rank < args.max_alt) | crowded[group]]
    lens = ((keys[kept] >> shift_len) & np.uint64(63)).astype(np.int64)
    width = int(lens.max()) if kept.size else 1
    letters = np.frombuffer(b'ACGT', dtype=np.uint8)[((keys[kept][:, None] >> (np.uint64(2) * np.arange(width, dtype=np.uint64))) &
                                                   np.uint64(3)).astype(np.int64)].tobytes()
    kept_code, kept_count, kept_first = refcode[kept].tolist(), counts[kept].tolist(), first[kept].tolist()
    per_code = collections.defaultdict(list)
    for i, n in enumerate(lens.tolist()):
        per_code[kept_code[i]].append((letters[i * width:i * width + n].decode(), kept_count[i], kept_first[i]))
    out = []
    totals_l, unchanged_l = totals.tolist(), unchanged.tolist()
    for code in np.flatnonzero(totals).tolist():
        kmer = ''.join('ACGT'[(code >> (2 * (k - 1 - j))) & 3] for j in range(k))
        total = totals_l[code]
        alts = per_code.get(code, [])
        if code in long_alts:
            alts = sorted(alts + [(a, c, s) for a, (c, s) in long_alts[code].items()], key=lambda x: (-x[1], x[2]))
        line = [f'{kmer},{unchanged_l[code] / total:.6f};']
        line.extend(f'{a},{c / total:.6f};' for a, c, _ in alts[:args.max_alt])
        out.append(''.join(line))
    print('\n'.join(out))


# ---------------------------------------------------------------------------------------------------- qscore model
def print_qscore_fractions(cigar, qscores, min_occur):
    """qscore_model.py:164-174; qscores: {quality value: count}."""
    total = sum(qscores.values())
    if total < min_occur:
        return
    fracs = ''.join(f'{q}:{float_to_str(qscores[q] / total, decimals=6, trim_zeros=True)},' for q in sorted(qscores))
    print(f'{cigar};{total};{fracs}')


def make_qscore_model(args, output=sys.stderr, dot_interval=1000):
    """qscore_model.py:78-161."""
    refs = load_fasta(args.reference)[0]
    reads = load_fastq(args.reads, output=output)
    alignments = load_alignments(args.alignment, args.max_alignments, output=output)
    if len(alignments) == 0:
        sys.exit('Error: no usable alignments')
    assert args.k_size % 2 == 1     # an odd size has a middle base to take the qscore from
    flat = FlatAlignments(alignments, reads, refs, output, dot_interval)
    keys, first, counts, overall, ovf = _count('cigars', flat, args.k_size, args.max_del)
    table = {}          # cigar -> [histogram, first occurrence]
    for key, stamp, hist in zip(keys.tolist(), first.tolist(), counts):
        n = key >> 58
        table[''.join(_SYM[(key >> (2 * j)) & 3] for j in range(n))] = [hist.astype(np.int64), stamp]
    overall = overall.astype(np.int64)
    cache = {}
    for a, i, kk in zip(*(o.tolist() for o in ovf)):    # CIGARs longer than a key holds (or odd quality characters)
        if a not in cache:
            cache[a] = flat.columns(a)
        _, _, sym, dcount, _, _, lead = cache[a]
        odd_quality = kk < 0
        kk = abs(kk)
        parts = ['D' * min(lead, args.max_del)] if i == 0 else []
        for j in range(kk):
            parts.append(_SYM[sym[i + j]])
            if j + 1 < kk:
                parts.append('D' * min(int(dcount[i + j]), args.max_del))
        cigar = ''.join(parts)
        q = int(flat.qual[flat.read_off[a] + i + (kk - 1) // 2]) - 33
        if odd_quality:
            sys.exit(f'Error: quality character {chr(q + 33)!r} outside the Phred+33 range')
        stamp = (a << 36) | (((kk - 1) // 2) << 32) | i
        entry = table.setdefault(cigar, [np.zeros(N_Q, dtype=np.int64), stamp])
        entry[0][q] += 1
        entry[1] = min(entry[1], stamp)
    print_qscore_fractions('overall', {q: int(c) for q, c in enumerate(overall) if c}, 0)
    order = sorted(table, key=lambda c: table[c][1])                       # insertion order of the reference's dict ...
    order.sort(key=lambda c: int(table[c][0].sum()), reverse=True)        # ... then stably by how common the CIGAR is
    for n, cigar in enumerate(order, start=1):
        print_qscore_fractions(cigar, {q: int(c) for q, c in enumerate(table[cigar][0]) if c}, args.min_occur)
        if n >= args.max_output:
            break
