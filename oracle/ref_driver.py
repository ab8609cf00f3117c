"""TEST INFRASTRUCTURE -- restatement of the stages of the reference's pipeline driver that sit either
side of the scan/scoring path, so that table-level parity (combined_genes_<LOCUS>.txt) can be checked
on a machine that has no reference checkout (the GPU box).

Follows /root/reference/run_iterative_igdetective.py:
  GetRange               :41-49        GetPositionRange   :68-79      GetFastaID          :81-90
  RunIgDetective         :92-128       CombineIGGenes     :130-151    AlignGenesIteratively :153-185
  UpdateVGeneDF / UpdateDJGeneDF :200-220                             CollectLocusSummary :222-243
  main (from "running IgDetective" on) :262-289

The two stage scripts are reached exactly as the reference reaches them -- `python py/IGDetective.py ...`
through os.system with the working directory holding `py/` and `datafiles/`, and
`import extract_aligned_genes` from `py/` -- so `py_dir` decides which implementation runs: the
reference's own scripts (on oracle/bio_shim; tests/golden/make_golden.py validates this restatement
against the UNMODIFIED driver that way) or the product's same-name shims in the repository's `py/`.
The prefilter (minimap2 + py/analyze_matches.py, :255-260) is outside the path: its output directory
`ig_contigs/` is an input here.

Pinned by tests/golden/driver/: make_golden.py asserts that this file, run over the reference's
scripts, writes byte-identical files to the unmodified run_iterative_igdetective.py.
"""
import importlib
import os
import shutil
import sys

import numpy as np
import pandas as pd

LOCI = ['IGH', 'IGK', 'IGL', 'TRA', 'TRB', 'TRG']          # :263


def read_fasta_records(path):
    """[(id, seq)] like SeqIO.parse(path, 'fasta') -> r.id, str(r.seq)."""
    recs, rid, chunks = [], None, []
    with open(path) as f:
        for line in f:
            if line.startswith('>'):
                if rid is not None:
                    recs.append((rid, ''.join(chunks)))
                tok = line[1:].split()
                rid, chunks = (tok[0] if tok else ''), []
            elif rid is not None:
                chunks.append(''.join(line.split()))
    if rid is not None:
        recs.append((rid, ''.join(chunks)))
    return recs


def get_range(min_pos, max_pos, seq_len, max_len=10000000):          # :41-49
    prefix_len = max_pos
    suffix_len = seq_len - min_pos
    gap = 1000000
    if prefix_len > max_len and suffix_len > max_len:
        return (max(min_pos - gap, 0), min(prefix_len + gap, seq_len))
    if prefix_len < suffix_len:
        return (0, min(prefix_len + gap, seq_len))
    return (max(min_pos - gap, 0), seq_len)


def get_position_range(sorted_positions):                             # :68-79
    if len(sorted_positions) == 1:
        return sorted_positions[0], sorted_positions[0]
    distances = [sorted_positions[i] - sorted_positions[i - 1] for i in range(1, len(sorted_positions))]
    median = np.median(distances)
    start_pos, end_pos = sorted_positions[0], sorted_positions[-1]
    if distances[0] / median > 100:
        start_pos = sorted_positions[1]
    if distances[-1] / median > 100:
        end_pos = sorted_positions[-2]
    return start_pos, end_pos


def get_fasta_id(igcontig_dir, contig_id, locus):                     # :81-90
    for f in os.listdir(igcontig_dir):
        if f.find(contig_id) == -1:
            continue
        if f.split('_')[0].find(locus) == -1:
            continue
        return os.path.join(igcontig_dir, f)
    return ''


def run_igdetective(igcontig_dir, output_dir, locus, python=sys.executable):       # :92-128
    txt = os.path.join(igcontig_dir, '__summary.txt')
    if not os.path.exists(txt):
        return
    df = pd.read_csv(txt, sep='\t', dtype={'ContigID': str})
    igh_df = df.loc[df['Locus'] == locus]
    if len(igh_df) == 0:
        return
    fasta = os.path.join(output_dir, 'combined_contigs_' + locus + '.fasta')
    with open(fasta, 'w') as fh:
        for c in set(igh_df['ContigID']):
            c_df = igh_df.loc[igh_df['ContigID'] == c]
            contig_fasta = get_fasta_id(igcontig_dir, c, locus)
            if contig_fasta == '' or not os.path.exists(contig_fasta):
                continue
            seq, seq_id = '', ''
            for rid, s in read_fasta_records(contig_fasta):
                seq, seq_id = s, rid
            positions = sorted(c_df['Position'])
            min_pos, max_pos = get_position_range(positions)
            a, b = get_range(min_pos, max_pos, len(seq))
            fh.write('>' + seq_id + '|START:' + str(a) + '|END:' + str(b) + '\n')
            fh.write(seq[a:b] + '\n')
    igdetective_dir = os.path.join(output_dir, 'predicted_genes_' + locus)
    cmd = python + ' py/IGDetective.py -i ' + fasta + ' -o ' + igdetective_dir + ' -m 1 -l ' + locus
    rc = os.system(cmd + ' > ' + os.path.join(output_dir, 'predicted_genes_' + locus + '.out'))
    if rc != 0:
        raise RuntimeError('%s failed (%d); see %s' % (cmd, rc, os.path.join(output_dir, 'predicted_genes_' + locus + '.out')))


def combine_ig_genes(genes_fasta, igdetective_tsv, output_fasta):     # :130-151
    nucl_seqs = set()
    if os.path.exists(genes_fasta):
        for _, s in read_fasta_records(genes_fasta):
            nucl_seqs.add(s)
    if os.path.exists(igdetective_tsv):
        df = pd.read_csv(igdetective_tsv, sep='\t', dtype={'reference contig': str})
        for i in range(len(df)):
            nucl_seqs.add(df['gene sequence'][i])
    reduced = []
    for seq1 in nucl_seqs:
        if not any(seq1 != seq2 and seq2.find(seq1) != -1 for seq2 in nucl_seqs):
            reduced.append(seq1)
    with open(output_fasta, 'w') as fh:
        for i, seq in enumerate(reduced):
            fh.write('>seq_' + str(i) + '\n' + seq + '\n')


def align_genes_iteratively(tools, ref_gene_fasta, igdetective_tsv, genome_fasta, output_dir, gene_type,
                            num_iter=5, combine=combine_ig_genes):     # :153-185
    iter0_dir = os.path.join(output_dir, gene_type + '_iter0')
    tools.main(genome_fasta, ref_gene_fasta, iter0_dir)
    combined_fasta = os.path.join(output_dir, gene_type + '_combined.fasta')
    combine(os.path.join(iter0_dir, 'genes.fasta'), igdetective_tsv, combined_fasta)
    prev_n = len(read_fasta_records(combined_fasta))
    if prev_n == 0:
        return
    prev_fasta = combined_fasta
    num_gene_dict = dict()
    for i in range(num_iter):
        iter_dir = os.path.join(output_dir, gene_type + '_iter' + str(i + 1))
        tools.main(genome_fasta, prev_fasta, iter_dir)
        curr_fasta = os.path.join(iter_dir, 'genes.fasta')
        if not os.path.exists(curr_fasta):
            break
        curr_n = len(read_fasta_records(curr_fasta))
        num_gene_dict[iter_dir] = curr_n
        if prev_n >= curr_n and i != 0:
            break
        prev_n, prev_fasta = curr_n, curr_fasta
    best_iter = sorted(num_gene_dict, key=lambda x: num_gene_dict[x], reverse=True)[0]
    shutil.copytree(best_iter, os.path.join(output_dir, gene_type + '_final'))       # cp -r, :185


def collect_locus_summary(denovo_dir, iter_dir, locus, output_fname):  # :200-243
    gene_dict = {'V': os.path.join(iter_dir, locus + 'V_final', 'genes.tsv')}
    if locus == 'IGH':
        gene_dict['D'] = os.path.join(denovo_dir, 'genes_D.tsv')
    gene_dict['J'] = os.path.join(denovo_dir, 'genes_J.tsv')
    sum_df = {'GeneType': [], 'Contig': [], 'Pos': [], 'Strand': [], 'Sequence': [], 'Productive': []}
    for gene_type in ['V', 'D', 'J']:
        if gene_type not in gene_dict or not os.path.exists(gene_dict[gene_type]):
            continue
        df = pd.read_csv(gene_dict[gene_type], sep='\t', dtype={'Contig': str})
        for i in range(len(df)):
            sum_df['GeneType'].append(gene_type)
            if gene_type == 'V':                                      # UpdateVGeneDF :200-208
                sum_df['Contig'].append(str(df['Contig'][i]).replace('|', '_'))
                sum_df['Pos'].append(df['Pos'][i])
                sum_df['Strand'].append(df['Strand'][i])
                sum_df['Sequence'].append(df['Seq'][i])
                sum_df['Productive'].append(df['Productive'][i])
            else:                                                     # UpdateDJGeneDF :210-220
                splits = df['reference contig'][i].split('|')
                start_pos = int(splits[2].split(':')[1])
                sum_df['Contig'].append(str(splits[0].split(':')[1]).replace('|', '_'))
                sum_df['Pos'].append(start_pos + df['start of gene'][i])
                sum_df['Strand'].append(df['strand'][i])
                sum_df['Sequence'].append(df['gene sequence'][i])
                sum_df['Productive'].append('NA')
    sum_df = pd.DataFrame(sum_df)
    sum_df['Locus'] = locus
    sum_df = sum_df.sort_values(by=['Contig', 'Pos'])
    sum_df.to_csv(output_fname, sep='\t', index=False)


def run(genome_fasta, igcontig_dir, output_dir, ig_gene_dir, py_dir='py', loci=LOCI, python=sys.executable,
        combine=None):
    """main() of the reference from `#### running IgDetective` (:262) to the combined tables (:289).
    Must run with the working directory that holds `py_dir` and `datafiles/` (as the reference does);
    `combine` replaces CombineIGGenes (same signature) when given."""
    if os.path.abspath(py_dir) not in [os.path.abspath(p) for p in sys.path]:
        sys.path.append(py_dir)                                       # sys.path.append('py'), :16
    tools = importlib.import_module('extract_aligned_genes')
    igdetect_dir = os.path.join(output_dir, 'denovo_search')
    os.mkdir(igdetect_dir)
    for locus in loci:
        run_igdetective(igcontig_dir, igdetect_dir, locus, python)
    ig_genes = {f.split('.')[0]: os.path.join(ig_gene_dir, f) for f in os.listdir(ig_gene_dir)}      # ReadGeneDir :187-193
    iter_dir = os.path.join(output_dir, 'iterative_search')
    os.mkdir(iter_dir)
    for locus in loci:
        gene = locus + 'V'
        if gene not in ig_genes:
            continue
        tsv = os.path.join(igdetect_dir, 'predicted_genes_' + locus, 'genes_V.tsv')
        align_genes_iteratively(tools, ig_genes[gene], tsv, genome_fasta, iter_dir, gene,
                                combine=combine or combine_ig_genes)
    out = []
    for locus in loci:
        txt = os.path.join(output_dir, 'combined_genes_' + locus + '.txt')
        collect_locus_summary(os.path.join(igdetect_dir, 'predicted_genes_' + locus), iter_dir, locus, txt)
        out.append(txt)
    return out


if __name__ == '__main__':
    # python oracle/ref_driver.py genome.fasta ig_contigs_dir output_dir ig_gene_dir [py_dir [module:function]]
    # the optional last argument names a replacement of CombineIGGenes (same signature) to run in its place
    os.makedirs(sys.argv[3], exist_ok=True)
    combine = None
    if len(sys.argv) > 6:
        mod, _, fn = sys.argv[6].partition(':')
        combine = getattr(importlib.import_module(mod), fn)
    run(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else 'py', combine=combine)
