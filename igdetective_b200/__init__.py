"""igdetective_b200 -- B200-native RSS scan + V/D/J scoring behind IgDetective's own function seams.

Layout:
  csrc/                 hand-written sm_100a CUDA kernels + the C ABI (include/igd_b200.h)
  _lib.py               ctypes binding of libigd_b200.so (fails loudly if the library is missing)
  engine.py             Engine: one handle / one GPU; numpy in, numpy out
  rss.py                get_contigwise_rss / combine_D_RSS / fragments (py/IGDetective.py:138-300)
  scoring.py            ComputeAlignment / align_fragment_to_genes / BioAlign view (:307-360, BioAlign)
  igdetective.py        the de-novo stage: same CLI flags and TSV outputs as py/IGDetective.py
  extract_aligned_genes.py  the iterative stage: same main() and outputs as py/extract_aligned_genes.py
  shard.py, multi.py    contig sharding across torchrun ranks, deterministic merge, sharded CLI
  fasta.py              FASTA readers (Python, and NativeFasta over the library's parallel C++ parser)
  datafiles.py, order.py  motif / gene-set loading; replay of the reference's set iteration order
There is no CPU implementation of the scan or the aligner in this package.
"""
__version__ = "0.1.0"
