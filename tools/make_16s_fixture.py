#!/usr/bin/env python
"""Derive rambl_b200/data/ecoli_mg1655_16S.ungapped.fa from the reference's aligned FASTA
(/root/reference/scripts/ecoli_mg1655_16S.fasta: one record, alignment gaps as '-').
The reference tree does not exist on the GPU box, so the 1542 bp sequence is committed as data."""
import sys, textwrap
src = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/scripts/ecoli_mg1655_16S.fasta'
dst = sys.argv[2] if len(sys.argv) > 2 else 'rambl_b200/data/ecoli_mg1655_16S.ungapped.fa'
lines = open(src).read().split('\n')
seq = ''.join(lines[1:]).replace('-', '').replace('.', '')
open(dst, 'w').write('>S000529092 E. coli MG1655 16S rRNA (gaps of the reference alignment removed)\n'
                     + '\n'.join(textwrap.wrap(seq, 70)) + '\n')
