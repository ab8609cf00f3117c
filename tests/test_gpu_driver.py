"""Table-level parity: combined_genes_<LOCUS>.txt -- the output of the whole pipeline -- and every
intermediate table, through the repository's same-name stage scripts (py/IGDetective.py as a
subprocess, py/extract_aligned_genes.py imported), exactly as run_iterative_igdetective.py reaches them.

The golden files under tests/golden/driver/tri_locus were written by the UNMODIFIED reference driver
(/root/reference/run_iterative_igdetective.py over the reference's own stage scripts; Bio shim, stub
minimap2, no-op plotting modules -- tests/golden/make_golden.py:make_driver).  The GPU box has no
reference checkout, so the driver's glue (locus cropping with |START:|END: headers, the iterative loop
feeding growing gene FASTAs back into the scoring, CollectLocusSummary) is run from its restatement
oracle/ref_driver.py, which make_golden.py pins byte for byte against the unmodified driver; its
CombineIGGenes step runs the product's indexed version (igdetective_b200.iterative).  The prefilter's
output (ig_contigs/) is an input, rebuilt here from the golden summary and manifest."""
import gzip
import os
import shutil
import subprocess
import sys

import pytest

from conftest import DATAFILES, GOLDEN, ROOT

CASE = os.path.join(GOLDEN, "driver", "tri_locus")


def _prepare(tmp_path):
    work = tmp_path / "work"
    work.mkdir()
    os.symlink(os.path.join(ROOT, "py"), work / "py")
    os.symlink(DATAFILES, work / "datafiles")
    genome = work / "tri_locus.fa"
    genome.write_bytes(gzip.open(os.path.join(GOLDEN, "inputs", "tri_locus.fa.gz")).read())
    contigs = {}
    rid = None
    for line in genome.read_text().splitlines():
        if line.startswith(">"):
            rid = line[1:].split()[0]
            contigs[rid.replace("|", "_")] = []
        else:
            contigs[rid.replace("|", "_")].append(line)
    contigs = {k: "".join(v) for k, v in contigs.items()}
    igc = work / "ig_contigs"
    igc.mkdir()
    shutil.copyfile(os.path.join(CASE, "ig_contigs", "__summary.txt"), igc / "__summary.txt")
    for line in open(os.path.join(CASE, "ig_contigs", "manifest.tsv")):
        fn, title, length = line.rstrip("\n").split("\t")
        if "START_POS" in title:
            continue                       # sub-locus extracts: the driver never reads them (GetFastaID, :81-90)
        cid = title.split("|")[0].split(":", 1)[1]
        assert len(contigs[cid]) == int(length)
        (igc / fn).write_text(">%s\n%s\n" % (title, contigs[cid]))
    return work, genome, igc


def _golden_files():
    out = []
    for d, _, files in os.walk(CASE):
        for fn in files:
            rel = os.path.relpath(os.path.join(d, fn), CASE)
            if rel.startswith("ig_contigs") or rel.endswith(".headers") or rel == "final_iterations.txt":
                continue
            out.append(rel)
    return sorted(out)


@pytest.mark.gpu
def test_combined_genes_tables_through_the_stage_shims(tmp_path):
    work, genome, igc = _prepare(tmp_path)
    env = dict(os.environ, PYTHONHASHSEED="0", PATH=os.path.join(ROOT, "oracle", "stubs") + os.pathsep + os.environ["PATH"])
    env["PYTHONPATH"] = ROOT               # the product package only: no Bio shim anywhere near it
    env.pop("IGD_STUB_SAM", None)
    env.pop("IGD_DATAFILES", None)         # the shims find datafiles/ under the working directory, like the reference
    out = work / "out"
    out.mkdir()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_driver.py"), str(genome), str(igc), str(out),
                        "datafiles/combined_reference_genes", "py", "igdetective_b200.iterative:CombineIGGenes"],
                       cwd=work, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    files = _golden_files()
    assert any(f.startswith("combined_genes_IGH") for f in files) and any("iterative_search" in f for f in files)
    for rel in files:
        got = out / rel
        assert got.exists(), "missing " + rel
        assert got.read_bytes() == open(os.path.join(CASE, rel), "rb").read(), rel + " differs from the reference's"
    # the cropped contigs the de-novo stage was given: same headers (|START:|END:) and lengths
    for fn in os.listdir(os.path.join(CASE, "denovo_search")):
        if fn.endswith(".headers"):
            got = []
            title, n = None, 0
            for line in open(out / "denovo_search" / fn[:-len(".headers")]):
                if line.startswith(">"):
                    if title is not None:
                        got.append(">%s\t%d\n" % (title, n))
                    title, n = line[1:].split()[0], 0
                else:
                    n += len(line.strip())
            got.append(">%s\t%d\n" % (title, n))
            assert got == open(os.path.join(CASE, "denovo_search", fn)).readlines()
    # a table with rows for each of the three loci, or the scenario tests nothing
    for locus in ("IGH", "IGK", "IGL"):
        assert len(open(os.path.join(CASE, "combined_genes_%s.txt" % locus)).readlines()) > 3


def test_golden_driver_case_is_complete():
    """CPU: the committed case holds what the GPU test compares (guards against a half-written golden)."""
    files = _golden_files()
    for locus in ("IGH", "IGK", "IGL"):
        assert "combined_genes_%s.txt" % locus in files
        assert any(f.startswith("iterative_search/%sV_iter" % locus) and f.endswith("genes.tsv") for f in files)
    assert "denovo_search/predicted_genes_IGH/genes_D.tsv" in files
    assert open(os.path.join(GOLDEN, "BIOPYTHON")).read().split()[0] in ("real", "shim")
