#!/usr/bin/env python3
"""Same-name replacement of the reference's py/IGDetective.py (the de-novo stage of IgDetective).

    python py/IGDetective.py -i genome.fasta -o out_dir -m 1 -l IGH [-r] [-g vdj]

Same flags, same working-directory convention (`datafiles/` is looked up under the directory the script
is started from, like py/IGDetective.py:54-58,133 of the reference) and the same output files
(genes_V.tsv, genes_D.tsv, genes_J.tsv; rss_<type>.csv with -r), so the unmodified
run_iterative_igdetective.py can shell out to it (:126-128).  The RSS scan and the candidate scoring
run in libigd_b200.so (hand-written sm_100a CUDA); there is no CPU fallback.
"""
import os
import sys

# the igdetective_b200 package: installed / on PYTHONPATH, else $IGD_B200_HOME, else next to this py/ directory
# (symlinks resolved: the directory may be linked into a reference checkout)
for _home in (os.environ.get("IGD_B200_HOME"), os.path.dirname(os.path.dirname(os.path.realpath(__file__)))):
    if _home and os.path.isdir(os.path.join(_home, "igdetective_b200")) and _home not in sys.path:
        sys.path.insert(0, _home)

from igdetective_b200.igdetective import main  # noqa: E402

if __name__ == "__main__":
    main()
