#!/usr/bin/env python3
"""
make_golden_models.py - fixtures for the model builders (`badread error_model` / `badread qscore_model`, SURVEY.md 8f
row f4).  TEST INFRASTRUCTURE.

Writes a small synthetic data set of its own making into tests/golden/models/ -

  ref.fasta     two contigs of random ACGT (one with a stretch of N)
  reads.fastq   reads cut from them (both strands), with substitutions, insertions and deletions (runs of up to 12)
                and unaligned adapter-like ends
  reads.paf     minimap2-style PAF with cg:Z: and AS:i: tags: the true alignment of every read, a second, worse
                alignment for some reads, and alignments the reference filters out (short, < 80 % identity)

- and then runs the UNMODIFIED reference (/root/reference; its `edlib` import is satisfied by oracle/edlib_shim, which
these two commands never call) on them:

  error_model_k7.txt.gz, error_model_k5_alt3.txt.gz, error_model_k4_max50.txt.gz     badread.error_model.make_error_model
  qscore_model_k9.txt.gz, qscore_model_k5_del3.txt.gz, qscore_model_k9_max40.txt.gz, qscore_model_k9_all.txt.gz
                                                                                 badread.qscore_model.make_qscore_model

Run once in the build container; the fixtures are committed, the reference is not needed at test time.
"""
import contextlib
import gzip
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.realpath(__file__))
OUT = os.path.join(HERE, '..', 'tests', 'golden', 'models')
COMP = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', 'N': 'N'}


def revcomp(s):
    return ''.join(COMP[c] for c in reversed(s))


def make_inputs():
    rs = np.random.RandomState(4242)
    acgt = np.array(list('ACGT'))
    refs = {}
    for name, n in (('ctgA', 24000), ('ctgB', 16000)):
        refs[name] = ''.join(acgt[rs.randint(0, 4, n)])
    refs['ctgB'] = refs['ctgB'][:7000] + 'N' * 40 + refs['ctgB'][7040:]
    reads, paf = [], []
    for i in range(90):
        ctg = 'ctgA' if rs.rand() < 0.6 else 'ctgB'
        n = int(rs.randint(300, 2200))
        start = int(rs.randint(0, len(refs[ctg]) - n))
        strand = '+' if rs.rand() < 0.5 else '-'
        seg = refs[ctg][start:start + n]
        if strand == '-':
            seg = revcomp(seg)
        rate = [0.03, 0.08, 0.15, 0.3][int(rs.randint(0, 4))] if i % 10 else 0.45   # every tenth read: < 80 % identity
        out, ops = [], []   # ops in read orientation, one symbol per column: M (match or mismatch), I, D
        j = 0
        while j < len(seg):
            x = rs.rand()
            if x < rate * 0.4:
                out.append(acgt[rs.randint(0, 4)]); ops.append('M'); j += 1
            elif x < rate * 0.7:
                run = int(rs.randint(1, 4)) if rs.rand() < 0.9 else int(rs.randint(5, 13))
                run = min(run, len(seg) - j)
                ops.extend('D' * run); j += run
            elif x < rate:
                run = int(rs.randint(1, 4)) if rs.rand() < 0.9 else int(rs.randint(5, 13))
                out.extend(acgt[rs.randint(0, 4, run)]); ops.extend('I' * run)
            else:
                out.append(seg[j]); ops.append('M'); j += 1
        while ops and ops[0] != 'M':    # alignments start and end on an aligned column
            if ops[0] == 'I':
                out.pop(0)
            else:
                seg = seg[1:]
                if strand == '+':
                    start += 1
            ops.pop(0)
        while ops and ops[-1] != 'M':
            if ops[-1] == 'I':
                out.pop()
            else:
                seg = seg[:-1]
                if strand == '-':
                    start += 1
            ops.pop()
        aligned_read = ''.join(out)
        n_ref = sum(1 for o in ops if o != 'I')
        head = ''.join(acgt[rs.randint(0, 4, int(rs.randint(0, 40)))])
        tail = ''.join(acgt[rs.randint(0, 4, int(rs.randint(0, 40)))])
        read = head + aligned_read + tail
        qual = ''.join(chr(33 + int(q)) for q in rs.randint(1, 41, len(read)))
        name = f'read{i:03d}'
        reads.append((name, read, qual))
        runs = []
        for o in ops:
            if runs and runs[-1][0] == o:
                runs[-1][1] += 1
            else:
                runs.append([o, 1])
        if strand == '-':
            runs = runs[::-1]          # PAF: CIGAR along the forward strand of the reference
        cigar = ''.join(f'{c}{o}' for o, c in runs)
        matches = sum(1 for k in range(len(ops)) if ops[k] == 'M')   # (an upper bound is fine: only the ratio is used)
        # count true matches so that the identity filter sees the real identity
        rp = fp = true_m = 0
        for o in ops:
            if o == 'M':
                true_m += aligned_read[rp] == seg[fp]; rp += 1; fp += 1
            elif o == 'I':
                rp += 1
            else:
                fp += 1
        fields = [name, str(len(read)), str(len(head)), str(len(head) + len(aligned_read)), strand, ctg,
                  str(len(refs[ctg])), str(start), str(start + n_ref), str(int(true_m)), str(len(ops)), '60',
                  'tp:A:P', f'AS:i:{2 * int(true_m) - 4 * (len(ops) - int(true_m))}', f'cg:Z:{cigar}']
        paf.append('\t'.join(fields))
        if i % 7 == 3:    # a second, worse alignment of the same read (lower AS): must lose
            m2 = max(120, n_ref // 3)
            fields2 = [name, str(len(read)), str(len(head)), str(len(head) + m2), '+', 'ctgA', str(len(refs['ctgA'])), '100',
                       str(100 + m2), str(m2 // 2), str(m2), '0', 'tp:A:S', 'AS:i:-50', f'cg:Z:{m2}M']
            paf.insert(len(paf) - 1 if i % 2 else len(paf), '\t'.join(fields2))
    # two hand-made alignments for the rare paths: a 30-base insertion (read k-mers of more than 22 bases) and a stretch of
    # single aligned bases between 7-base deletions (CIGAR windows of more than 29 symbols)
    seg = refs['ctgA'][3000:3400]
    ins = ''.join(acgt[rs.randint(0, 4, 30)])
    read = seg[:200] + ins + seg[200:]
    reads.append(('readlongins', read, ''.join(chr(33 + int(q)) for q in rs.randint(1, 41, len(read)))))
    paf.append('\t'.join(['readlongins', str(len(read)), '0', str(len(read)), '+', 'ctgA', str(len(refs['ctgA'])), '3000', '3400',
                          '400', '430', '60', 'AS:i:700', 'cg:Z:200M30I200M']))
    seg = refs['ctgA'][5000:5400]
    keep, cigar, pos = [seg[:150]], ['150M'], 150
    for _ in range(12):
        cigar.append('7D'); pos += 7
        keep.append(seg[pos]); cigar.append('1M'); pos += 1
    keep.append(seg[pos:]); cigar.append(f'{len(seg) - pos}M')
    read = ''.join(keep)
    reads.append(('readgappy', read, ''.join(chr(33 + int(q)) for q in rs.randint(1, 41, len(read)))))
    paf.append('\t'.join(['readgappy', str(len(read)), '0', str(len(read)), '+', 'ctgA', str(len(refs['ctgA'])), '5000', '5400',
                          '330', '400', '60', 'AS:i:500', 'cg:Z:' + ''.join(cigar)]))   # (82.5 %: passes the identity filter)
    # a short alignment (<= 100 columns) of an extra read: filtered out
    reads.append(('readshort', refs['ctgA'][500:580], 'I' * 80))
    paf.append('\t'.join(['readshort', '80', '0', '80', '+', 'ctgA', str(len(refs['ctgA'])), '500', '580', '80', '80', '60',
                          'AS:i:160', 'cg:Z:80M']))
    return refs, reads, paf


def main():
    os.makedirs(OUT, exist_ok=True)
    refs, reads, paf = make_inputs()
    with open(os.path.join(OUT, 'ref.fasta'), 'w') as f:
        for name, seq in refs.items():
            f.write(f'>{name}\n')
            for i in range(0, len(seq), 80):
                f.write(seq[i:i + 80] + '\n')
    with open(os.path.join(OUT, 'reads.fastq'), 'w') as f:
        for name, seq, qual in reads:
            f.write(f'@{name} some description\n{seq}\n+\n{qual}\n')
    with open(os.path.join(OUT, 'reads.paf'), 'w') as f:
        f.write('\n'.join(paf) + '\n')

    sys.path.insert(0, os.path.join(HERE, 'edlib_shim'))
    sys.path.insert(0, '/root/reference')
    import badread.error_model as rem
    import badread.qscore_model as rqm

    def run(fn, out_name, max_alignments=None, **kw):
        args = types.SimpleNamespace(reference=os.path.join(OUT, 'ref.fasta'), reads=os.path.join(OUT, 'reads.fastq'),
                                     alignment=os.path.join(OUT, 'reads.paf'), max_alignments=max_alignments, **kw)
        buf, sink = io.StringIO(), io.StringIO()
        with contextlib.redirect_stdout(buf):
            fn(args, output=sink)
        with gzip.GzipFile(os.path.join(OUT, out_name + '.gz'), 'wb', mtime=0) as f:   # (mtime 0: reproducible bytes)
            f.write(buf.getvalue().encode())
        print(out_name, len(buf.getvalue().splitlines()), 'lines')

    run(rem.make_error_model, 'error_model_k7.txt', k_size=7, max_alt=25)
    run(rem.make_error_model, 'error_model_k5_alt3.txt', k_size=5, max_alt=3)
    run(rem.make_error_model, 'error_model_k4_max50.txt', max_alignments=50, k_size=4, max_alt=25)
    run(rqm.make_qscore_model, 'qscore_model_k9.txt', k_size=9, max_del=6, min_occur=3, max_output=10000)
    run(rqm.make_qscore_model, 'qscore_model_k5_del3.txt', k_size=5, max_del=3, min_occur=1, max_output=10000)
    run(rqm.make_qscore_model, 'qscore_model_k9_max40.txt', k_size=9, max_del=6, min_occur=100, max_output=40)
    run(rqm.make_qscore_model, 'qscore_model_k9_all.txt', k_size=9, max_del=6, min_occur=1, max_output=1000000)


if __name__ == '__main__':
    main()
