"""Generates everything under tests/golden/ from the reference checkout.  Run HERE (the
container that has /root/reference); the GPU box only ever sees the committed outputs.

    python tests/golden/make_golden.py [--only data|igdetective|scoring_b] [--jobs 8]

1. data fixtures: motif sets, V/J reference gene sets and the three example loci, re-encoded
   (text / gzip) from /root/reference/datafiles and /root/reference/examples.
2. igdetective/: genes_{V,D,J}.tsv and rss_*.csv written by the UNMODIFIED reference script
   /root/reference/py/IGDetective.py running on oracle/bio_shim (no BioPython here).
3. scoring_b/: genes.tsv / genes.fasta written by the UNMODIFIED
   /root/reference/py/extract_aligned_genes.py main() on the shim, with oracle/stubs/minimap2
   serving a fixed SAM (minimap2 is not installed either).
"""
import argparse
import gzip
import os
import pickle
import shutil
import subprocess
import sys
import tempfile

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SHIM = os.path.join(ROOT, "oracle", "bio_shim")
STUBS = os.path.join(ROOT, "oracle", "stubs")

EXAMPLES = {"human": "human.fa", "mouse": "mouse.fa", "cow": "cow.fasta"}


def real_biopython():
    """Version string of a real BioPython >= 1.82 importable WITHOUT the shim on the path, else None."""
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, "-c", "import Bio; from Bio import Align, SeqIO; from Bio.Seq import Seq; print(Bio.__version__)"],
                       capture_output=True, text=True, env=env)
    if r.returncode != 0:
        return None
    v = r.stdout.strip()
    try:
        major, minor = (int(x) for x in v.split(".")[:2])
    except ValueError:
        return None
    return v if (major, minor) >= (1, 82) else None


# The moment a real BioPython is importable the reference scripts run on IT, not on oracle/bio_shim: the
# goldens are then BioPython's own output and every oracle / GPU parity test is pinned to the real library
# (tests/test_oracle_golden.py fails loudly on any difference).  tests/golden/BIOPYTHON records which it was.
REAL_BIO = real_biopython()
if REAL_BIO:
    SHIM = ""


def read_fasta(path):
    recs, title, chunks = [], None, []
    for line in open(path):
        if line.startswith(">"):
            if title is not None:
                recs.append((title, "".join(chunks)))
            title, chunks = line[1:].rstrip("\n"), []
        else:
            chunks.append(line.strip())
    if title is not None:
        recs.append((title, "".join(chunks)))
    return recs


def write_fasta_gz(path, recs):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as raw:
        with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as f:
            for t, s in recs:
                f.write((">%s\n%s\n" % (t, s)).encode())


def make_data():
    d = os.path.join(HERE, "datafiles")
    os.makedirs(d, exist_ok=True)
    motifs = pickle.load(open(os.path.join(REF, "datafiles", "motifs"), "rb"))
    with open(os.path.join(d, "motif_sets.txt"), "w") as f:
        for t in ("V", "D_left", "D_right", "J"):
            for k in ("7", "9"):
                for kmer in sorted(motifs[t][k]):
                    f.write("%s\t%s\t%s\n" % (t, k, kmer))
    for sub in ("combined_reference_genes", "human_reference_genes"):
        for fn in sorted(os.listdir(os.path.join(REF, "datafiles", sub))):
            if fn[3] in "VJ":      # C genes are used only by the minimap2 prefilter
                write_fasta_gz(os.path.join(d, sub, fn + ".gz"), read_fasta(os.path.join(REF, "datafiles", sub, fn)))
    for name, fn in EXAMPLES.items():
        write_fasta_gz(os.path.join(HERE, "inputs", name + ".fa.gz"),
                       read_fasta(os.path.join(REF, "examples", "igh_locus_examples", fn)))


def make_multi_contig_input():
    """A multi-contig FASTA that walks the reference's corner cases: slices of the three loci (one all
    lower-case), a repeated record id, contigs shorter than the 350-nt V flank with an RSS inside
    (negative slice start, IGDetective.py:265), an RSS within 350 nt of the start of a longer contig
    (empty fragment -> 'A' * len(gene), :339-340), an empty record."""
    motifs = pickle.load(open(os.path.join(REF, "datafiles", "motifs"), "rb"))
    ex = {n: read_fasta(os.path.join(REF, "examples", "igh_locus_examples", fn))[0][1] for n, fn in EXAMPLES.items()}
    v9, j9, j7 = sorted(motifs["V"]["9"]), sorted(motifs["J"]["9"]), sorted(motifs["J"]["7"])
    comp = str.maketrans("ACGT", "TGCA")
    v_unit = "CACAGTG" + "ACGTACGTTGCAACGTTGCAATC" + v9[10]                 # heptamer, 23 nt, nonamer
    j_unit = j9[3] + "TTGACCATGCAAGTCCGATAGCA" + j7[5]                      # nonamer, 23 nt, heptamer
    filler = ex["human"][500000:500600]
    recs = [
        ("ctgA human slice", ex["human"][1000000:1323718]),
        ("ctgB first copy, replaced below", ex["mouse"][:1000]),
        ("ctgC", ex["cow"][:250000]),
        ("short_v 300 nt, V RSS at 250", filler[:250] + v_unit + filler[:11]),
        ("early_v V RSS 100 nt into a 900-nt contig", filler[:100] + v_unit + filler[:600] + filler[:161]),
        ("rev_units", (filler[:80] + j_unit + filler[:400] + v_unit + filler[:90]).translate(comp)[::-1]),
        ("ctgB", ex["mouse"][2400000:2751186]),
        ("empty_record", ""),
        ("short_j 60 nt", j_unit + filler[:21]),
    ]
    path = os.path.join(HERE, "inputs", "multi.fa.gz")
    write_fasta_gz(path, recs)
    return path


def run_reference_on_file(fasta_gz, tag, jobs):
    out = os.path.join(HERE, "igdetective", tag)
    os.makedirs(out, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "datafiles"))
        os.symlink(os.path.join(REF, "datafiles", "motifs"), os.path.join(tmp, "datafiles", "motifs"))
        os.symlink(os.path.join(REF, "datafiles", "combined_reference_genes"),
                   os.path.join(tmp, "datafiles", "combined_reference_genes"))
        plain = os.path.join(tmp, "input.fa")
        with open(plain, "wb") as f:
            f.write(gzip.open(fasta_gz).read())
        env = dict(os.environ, PYTHONPATH=SHIM, PYTHONHASHSEED="0")
        for extra in ([], ["-r"]):
            cmd = [sys.executable, os.path.join(REF, "py", "IGDetective.py"), "-i", plain,
                   "-o", os.path.join(tmp, "out"), "-m", str(jobs), "-l", "IGH"] + extra
            subprocess.run(cmd, cwd=tmp, env=env, check=True, stdout=subprocess.DEVNULL)
        for fn in sorted(os.listdir(os.path.join(tmp, "out"))):
            shutil.copyfile(os.path.join(tmp, "out", fn), os.path.join(out, fn))
            print("  wrote", os.path.relpath(os.path.join(out, fn), ROOT))


def run_reference_igdetective(example, gene_dir, jobs, rss_only, locus="IGH"):
    """cwd gets a datafiles/ view whose combined_reference_genes points at `gene_dir`
    (the script hard-codes that name, IGDetective.py:133)."""
    out = os.path.join(HERE, "igdetective", "%s__%s" % (example, gene_dir) + ("" if locus == "IGH" else "__" + locus))
    os.makedirs(out, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "datafiles"))
        os.symlink(os.path.join(REF, "datafiles", "motifs"), os.path.join(tmp, "datafiles", "motifs"))
        os.symlink(os.path.join(REF, "datafiles", gene_dir),
                   os.path.join(tmp, "datafiles", "combined_reference_genes"))
        env = dict(os.environ, PYTHONPATH=SHIM, PYTHONHASHSEED="0")
        cmd = [sys.executable, os.path.join(REF, "py", "IGDetective.py"),
               "-i", os.path.join(REF, "examples", "igh_locus_examples", EXAMPLES[example]),
               "-o", os.path.join(tmp, "out"), "-m", str(jobs), "-l", locus] + (["-r"] if rss_only else [])
        subprocess.run(cmd, cwd=tmp, env=env, check=True, stdout=subprocess.DEVNULL)
        for fn in sorted(os.listdir(os.path.join(tmp, "out"))):
            shutil.copyfile(os.path.join(tmp, "out", fn), os.path.join(out, fn))
            print("  wrote", os.path.relpath(os.path.join(out, fn), ROOT))


def make_igdetective(jobs):
    for ex in EXAMPLES:
        run_reference_igdetective(ex, "combined_reference_genes", jobs, rss_only=True)
        run_reference_igdetective(ex, "combined_reference_genes", jobs, rss_only=False)
    run_reference_igdetective("human", "human_reference_genes", jobs, rss_only=False)
    # other loci select other gene sets (IGKV 358 / IGKJ 14 mixed case, IGLV 200 / IGLJ 20, TRBV 66 / TRBJ 36);
    # the spacers stay the IGH ones because InitializeVariables (:36-51) only assigns locals
    for locus in ("IGK", "IGL", "TRB"):
        run_reference_igdetective("cow", "combined_reference_genes", jobs, rss_only=False, locus=locus)
    run_reference_on_file(make_multi_contig_input(), "multi__combined_reference_genes", jobs)


# scoring-B scenarios: (name, example, gene fasta, 1-based SAM positions)
def scoring_b_scenarios():
    import csv
    sc = []
    for ex, gene_dir in (("human", "combined_reference_genes"), ("cow", "combined_reference_genes")):
        tsv = os.path.join(HERE, "igdetective", "%s__%s" % (ex, gene_dir), "genes_V.tsv")
        rows = list(csv.reader(open(tsv, newline=""), delimiter="\t"))[1:]
        starts = sorted(int(r[6]) for r in rows)
        n = 12 if ex == "human" else 5
        pos = [starts[i * len(starts) // n] + 1 + (37 * i) % 90 for i in range(n)]
        seqlen = len(read_fasta(os.path.join(REF, "examples", "igh_locus_examples", EXAMPLES[ex]))[0][1])
        pos += [5, 250, 390, 2000, 2300, seqlen - 120, seqlen - 500, seqlen // 2, seqlen // 3]
        pos += [pos[0]]                       # duplicate POS is dropped by ProcessSamFile
        sc.append(("%s_IGHV" % ex, ex, os.path.join(REF, "datafiles", gene_dir, "IGHV.fa"), pos))
    return sc


def make_scoring_b_multi():
    """scoring B on the multi-contig input against the short, lower-case IGHJ gene file: positions on
    several contigs (one all lower-case), near contig ends, on contigs shorter than the window"""
    import csv
    if SHIM:
        sys.path.insert(0, SHIM)
    sys.path.insert(0, os.path.join(REF, "py"))
    os.environ["PATH"] = STUBS + os.pathsep + os.environ["PATH"]
    import extract_aligned_genes as ref_mod
    rows = list(csv.reader(open(os.path.join(HERE, "igdetective", "multi__combined_reference_genes", "genes_J.tsv"),
                                newline=""), delimiter="\t"))[1:]
    hits = [(r[0], int(r[6]) + 1) for r in rows] + [("ctgC", 120000), ("ctgC", 120350), ("ctgC", 249990), ("ctgA", 5),
            ("short_j", 10), ("short_v", 200), ("rev_units", 300), ("ctgB", 1000), ("ctgB", 350000)]
    out = os.path.join(HERE, "scoring_b", "multi_IGHJ")
    os.makedirs(out, exist_ok=True)
    sam = os.path.join(out, "alignment.sam")
    with open(sam, "w") as f:
        f.write("@HD\tVN:1.6\n")
        for i, (c, p) in enumerate(hits):
            f.write("g%d\t0\t%s\t%d\t60\t*\t*\t0\t0\t*\t*\n" % (i, c, p))
    os.environ["IGD_STUB_SAM"] = sam
    with tempfile.TemporaryDirectory() as tmp:
        plain = os.path.join(tmp, "multi.fa")
        with open(plain, "wb") as f:
            f.write(gzip.open(os.path.join(HERE, "inputs", "multi.fa.gz")).read())
        ref_mod.main(plain, os.path.join(REF, "datafiles", "combined_reference_genes", "IGHJ.fa"), os.path.join(tmp, "o"))
        for fn in ("genes.tsv", "genes.fasta"):
            shutil.copyfile(os.path.join(tmp, "o", fn), os.path.join(out, fn))
            print("  wrote", os.path.relpath(os.path.join(out, fn), ROOT))


def make_scoring_b():
    if SHIM:
        sys.path.insert(0, SHIM)
    sys.path.insert(0, os.path.join(REF, "py"))
    os.environ["PATH"] = STUBS + os.pathsep + os.environ["PATH"]
    import extract_aligned_genes as ref_mod
    for name, ex, gene_fa, positions in scoring_b_scenarios():
        out = os.path.join(HERE, "scoring_b", name)
        os.makedirs(out, exist_ok=True)
        genome = os.path.join(REF, "examples", "igh_locus_examples", EXAMPLES[ex])
        contig = read_fasta(genome)[0][0].split()[0]
        sam = os.path.join(out, "alignment.sam")
        with open(sam, "w") as f:
            f.write("@HD\tVN:1.6\n")
            for i, p in enumerate(positions):
                f.write("g%d\t0\t%s\t%d\t60\t*\t*\t0\t0\t*\t*\n" % (i, contig, p))
            f.write("gu\t4\t*\t0\t0\t*\t*\t0\t0\t*\t*\n")
        os.environ["IGD_STUB_SAM"] = sam
        with tempfile.TemporaryDirectory() as tmp:
            ref_mod.main(genome, gene_fa, os.path.join(tmp, "o"))
            for fn in ("genes.tsv", "genes.fasta"):
                shutil.copyfile(os.path.join(tmp, "o", fn), os.path.join(out, fn))
                print("  wrote", os.path.relpath(os.path.join(out, fn), ROOT))



# ---- table level: the whole pipeline driver ----------------------------------------------------------
PYLIBS = os.path.join(STUBS, "pylibs")
DRIVER_KEEP = ("combined_genes_", "denovo_search/predicted_genes_", "iterative_search/")


def make_driver_input():
    """tri_locus: a real IGH locus slice (V, D and J genes of human.fa) plus two synthetic contigs with
    mutated IGK / IGL V and J genes and their RSSs planted on both strands, partly soft-masked; one id
    carries a '|' (the driver rewrites it to '_', run_iterative_igdetective.py:203,214)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    motifs = pickle.load(open(os.path.join(REF, "datafiles", "motifs"), "rb"))
    human = read_fasta(os.path.join(REF, "examples", "igh_locus_examples", EXAMPLES["human"]))[0][1]
    recs = [("chr14_igh_slice", human[850000:1030000])]
    for seed, locus, cid in ((31, "IGK", "ctgK|synthetic"), (32, "IGL", "ctgL")):
        gv = dict(read_fasta(os.path.join(REF, "datafiles", "combined_reference_genes", locus + "V.fa")))
        gj = dict(read_fasta(os.path.join(REF, "datafiles", "combined_reference_genes", locus + "J.fa")))
        gv = {k.split()[0]: v.upper() for k, v in gv.items()}
        gj = {k.split()[0]: v.upper() for k, v in gj.items()}
        contigs, _ = synth.assembly(seed, [60000], motifs, gv, gj, plant_every=0, n_runs=1, soft_mask=0.3,
                                    clusters=[(0, 5, 0, 2)])
        recs.append((cid, contigs[0].tobytes().decode()))
    path = os.path.join(HERE, "inputs", "tri_locus.fa.gz")
    write_fasta_gz(path, recs)
    return path


def _tree(root):
    out = {}
    for d, _, files in os.walk(root):
        for fn in files:
            p = os.path.join(d, fn)
            out[os.path.relpath(p, root)] = p
    return out


def make_driver(jobs):
    """combined_genes_<LOCUS>.txt and every intermediate table written by the UNMODIFIED
    /root/reference/run_iterative_igdetective.py on tri_locus (Bio shim, stub minimap2, no-op plotting
    modules), then the restated driver (oracle/ref_driver.py) over the same reference scripts, which
    must write the same bytes -- that is what pins the restatement the GPU-side test drives."""
    src = make_driver_input()
    out = os.path.join(HERE, "driver", "tri_locus")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    env = dict(os.environ, PYTHONPATH=SHIM + os.pathsep + PYLIBS, PYTHONHASHSEED="0",
               PATH=STUBS + os.pathsep + os.environ["PATH"])
    env.pop("IGD_STUB_SAM", None)
    with tempfile.TemporaryDirectory() as tmp:
        for name in ("run_iterative_igdetective.py", "py", "datafiles"):
            os.symlink(os.path.join(REF, name), os.path.join(tmp, name))
        genome = os.path.join(tmp, "tri_locus.fa")
        with open(genome, "wb") as f:
            f.write(gzip.open(src).read())
        subprocess.run([sys.executable, "run_iterative_igdetective.py", genome, os.path.join(tmp, "out")],
                       cwd=tmp, env=env, check=True, stdout=open(os.path.join(tmp, "driver.log"), "w"))
        ref_out = os.path.join(tmp, "out")
        # the driver deletes ig_contigs/*.fasta at its end (CleanLargeContigs, :195-198): re-run the prefilter's
        # analysis step on the alignments it left behind to get the directory back as RunIgDetective saw it
        subprocess.run([sys.executable, "py/analyze_matches.py", os.path.join(ref_out, "initial_alignments"),
                        os.path.join(tmp, "ig_contigs"), genome], cwd=tmp, env=env, check=True, stdout=subprocess.DEVNULL)
        assert open(os.path.join(tmp, "ig_contigs", "__summary.txt")).read() == \
            open(os.path.join(ref_out, "ig_contigs", "__summary.txt")).read()
        # the restated driver over the same reference scripts
        subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_driver.py"), genome, os.path.join(tmp, "ig_contigs"),
                        os.path.join(tmp, "out2"), "datafiles/combined_reference_genes"], cwd=tmp, env=env, check=True,
                       stdout=subprocess.DEVNULL)
        a, b = _tree(ref_out), _tree(os.path.join(tmp, "out2"))
        checked = 0
        for rel in sorted(a):
            if not rel.startswith(DRIVER_KEEP) or rel.endswith((".sam", ".out")):
                continue
            assert rel in b, "restated driver did not write " + rel
            assert open(a[rel], "rb").read() == open(b[rel], "rb").read(), "restated driver differs in " + rel
            checked += 1
        print("  restated driver == unmodified driver on %d files" % checked)
        # what the GPU-side test needs: the prefilter's summary + which FASTA files it wrote, and the tables
        os.makedirs(os.path.join(out, "ig_contigs"))
        shutil.copyfile(os.path.join(tmp, "ig_contigs", "__summary.txt"), os.path.join(out, "ig_contigs", "__summary.txt"))
        with open(os.path.join(out, "ig_contigs", "manifest.tsv"), "w") as mf:
            for fn in sorted(os.listdir(os.path.join(tmp, "ig_contigs"))):
                if fn.endswith(".fasta"):
                    recs = read_fasta(os.path.join(tmp, "ig_contigs", fn))
                    for title, s in recs:
                        mf.write("%s\t%s\t%d\n" % (fn, title, len(s)))
        for rel in sorted(a):
            keep = rel.startswith(DRIVER_KEEP) and rel.endswith((".txt", ".tsv", "_combined.fasta", "genes.fasta"))
            if rel.startswith("denovo_search/combined_contigs_"):
                keep = True
            if "_final/" in rel:
                keep = False                       # a copy of the best iteration
            if not keep:
                continue
            dst = os.path.join(out, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if rel.startswith("denovo_search/combined_contigs_"):      # headers only: the letters are slices of the input
                with open(dst + ".headers", "w") as f:
                    f.writelines(">%s\t%d\n" % (t, len(s)) for t, s in read_fasta(a[rel]))
            else:
                shutil.copyfile(a[rel], dst)
            print("  wrote", os.path.relpath(dst, ROOT))
        with open(os.path.join(out, "final_iterations.txt"), "w") as f:
            for rel in sorted(a):
                if "_final/genes.tsv" in rel:
                    gene = rel.split("/")[1][:-len("_final")]
                    want = open(a[rel], "rb").read()
                    its = [r for r in sorted(a) if r.startswith("iterative_search/%s_iter" % gene) and r.endswith("genes.tsv")
                           and open(a[r], "rb").read() == want]
                    f.write("%s\t%s\n" % (gene, ",".join(i.split("/")[1] for i in its)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--jobs", type=int, default=len(os.sched_getaffinity(0)))
    a = ap.parse_args()
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)
    print("BioPython:", "real " + REAL_BIO if REAL_BIO else "oracle/bio_shim (real BioPython not importable)")
    with open(os.path.join(HERE, "BIOPYTHON"), "w") as f:
        f.write(("real %s\n" % REAL_BIO) if REAL_BIO else "shim (oracle/bio_shim over oracle/gotoh_ref.c; BioPython not importable here)\n")
    if a.only in ("", "data"):
        make_data()
    if a.only in ("", "igdetective"):
        make_igdetective(a.jobs)
    if a.only in ("", "scoring_b"):
        make_scoring_b()
        make_scoring_b_multi()
    if a.only in ("", "driver"):
        make_driver(a.jobs)
