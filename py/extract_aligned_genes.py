"""Same-name replacement of the reference's py/extract_aligned_genes.py (the iterative stage).

The unmodified run_iterative_igdetective.py does `sys.path.append('py'); import extract_aligned_genes as
gene_finding_tools` (:16-17) and calls `gene_finding_tools.main(genome_fasta, gene_fasta, output_dir)`
(:156,171); py/IGDetective.py of the reference uses SetupAligner and BioAlign from here (:16).  All of
these names resolve to igdetective_b200.extract_aligned_genes, whose alignments run in libigd_b200.so.
"""
import os
import sys

# the igdetective_b200 package: installed / on PYTHONPATH, else $IGD_B200_HOME, else next to this py/ directory
# (symlinks resolved: the directory may be linked into a reference checkout)
for _home in (os.environ.get("IGD_B200_HOME"), os.path.dirname(os.path.dirname(os.path.realpath(__file__)))):
    if _home and os.path.isdir(os.path.join(_home, "igdetective_b200")) and _home not in sys.path:
        sys.path.insert(0, _home)

from igdetective_b200.extract_aligned_genes import (  # noqa: E402,F401
    Alignment, BioAlign, ComputeAlignment, GpuAligner, PrepareOutputDir, ProcessSamFile, SetupAligner, main)

if __name__ == '__main__':
    if len(sys.argv) != 4:
        print('Invalid arguments')
        print('python extract_aligned_genes.py genome.fasta reference_IG_genes.fasta output_dir')
        sys.exit(1)
    main(sys.argv[1], sys.argv[2], sys.argv[3])
