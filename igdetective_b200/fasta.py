"""FASTA reading with the semantics of SeqIO.parse(path, 'fasta') as the reference uses it
(py/IGDetective.py:129,134; py/extract_aligned_genes.py:143-151): record id = first
whitespace-delimited token of the header, sequence lines concatenated, case preserved."""
import gzip

_COMP = str.maketrans("ACGTMRWSYKVHDBXNacgtmrwsykvhdbxnUu", "TGCAKYWSRMBDHVXNtgcakywsrmbdhvxnAa")


def reverse_complement(s):
    """str(Seq(s).reverse_complement()): IUPAC-aware, case preserving."""
    return s.translate(_COMP)[::-1]


def read_fasta(path):
    """-> [(id, sequence)] in file order."""
    op = gzip.open if str(path).endswith(".gz") else open
    recs, title, chunks = [], None, []
    with op(path, "rt") as fh:
        for line in fh:
            if line.startswith(">"):
                if title is not None:
                    recs.append((title, "".join(chunks)))
                t = line[1:].rstrip()
                title = t.split(None, 1)[0] if t.strip() else ""
                chunks = []
            elif title is not None:
                chunks.append("".join(line.split()))
    if title is not None:
        recs.append((title, "".join(chunks)))
    return recs


def fasta_dict(path, upper=False):
    """{rec.id: seq} built like the reference's dict comprehensions: a repeated id keeps its first
    position and takes the last sequence."""
    d = {}
    for rid, s in read_fasta(path):
        d[rid] = s.upper() if upper else s
    return d


class ArraySeq:
    """str-like view of a contig held as a uint8 array (e.g. a slice of pinned host memory): len() and slicing with
    Python semantics return str; nothing else is copied.  What parent_seq[contig] is for callers that keep genomes
    as byte buffers."""

    def __init__(self, a):
        self.a = a

    def __len__(self):
        return len(self.a)

    def __getitem__(self, s):
        return self.a[s].tobytes().decode("latin-1")

    def __str__(self):
        return self.a.tobytes().decode("latin-1")


class ContigView:
    """One contig of a NativeFasta: len() and slicing with Python semantics (negative indices wrap,
    out-of-range bounds clamp), returning str -- what the reference does with parent_seq[contig]."""

    def __init__(self, owner, index, length):
        self._o, self._i, self._n = owner, index, length

    def __len__(self):
        return self._n

    def __getitem__(self, key):
        if isinstance(key, slice):
            a, b, step = key.indices(self._n)
            if step != 1:
                return str(self)[key]
            return self._o.read(self._i, a, max(0, b - a))
        if key < 0:
            key += self._n
        return self._o.read(self._i, key, 1)

    def __str__(self):
        return self._o.read(self._i, 0, self._n)


class NativeFasta:
    """FASTA file parsed by the C++ reader of libigd_b200.so (igd_fasta_*): a read-only mapping
    {record id: ContigView} in the reference's dict order.  The letters live in host memory owned by
    the library; Engine.load_fasta(self) sends them to the GPU without a Python copy."""

    def __init__(self, path, byte_range=None, margin=65536):
        """byte_range=(lo, hi): the range form (igd_fasta_open_range) -- records in file order, record 0 the lead;
        see RangeFasta for the bookkeeping of a sharded run."""
        import ctypes

        from . import _lib
        self._lib = _lib.load()
        self._ct = ctypes
        if byte_range is None:
            self._f = self._lib.igd_fasta_open(str(path).encode())
        else:
            self._f = self._lib.igd_fasta_open_range(str(path).encode(), int(byte_range[0]), int(byte_range[1]), int(margin))
        if not self._f:
            raise _lib.IgdError(_lib.IGD_ERR_ARG, self._lib.igd_last_error().decode("utf-8", "replace"))
        self.ids, self.lengths = [], []
        cid, n = ctypes.c_char_p(), ctypes.c_uint64()
        for i in range(self._lib.igd_fasta_count(self._f)):
            _lib.check(self._lib.igd_fasta_info(self._f, i, ctypes.byref(cid), ctypes.byref(n)))
            self.ids.append(cid.value.decode("latin-1"))
            self.lengths.append(n.value)
        self._views = {rid: ContigView(self, i, L) for i, (rid, L) in enumerate(zip(self.ids, self.lengths))}
        self.byte_range = None
        if byte_range is not None:
            lo, hi, a, b = (ctypes.c_uint64() for _ in range(4))
            _lib.check(self._lib.igd_fasta_range(self._f, ctypes.byref(lo), ctypes.byref(hi), ctypes.byref(a), ctypes.byref(b)))
            self.byte_range, self.own_letters = (lo.value, hi.value), (a.value, b.value)
            self.header_bytes = []
            hb = ctypes.c_uint64()
            for i in range(len(self.ids)):
                _lib.check(self._lib.igd_fasta_header_byte(self._f, i, ctypes.byref(hb)))
                self.header_bytes.append(None if hb.value == 2 ** 64 - 1 else hb.value)

    def view(self, i):
        """record i as a str-like view (the range form has no unique ids to index by)"""
        return ContigView(self, i, self.lengths[i])

    def read(self, i, start, n):
        from . import _lib
        buf = self._ct.create_string_buffer(n) if n else None
        _lib.check(self._lib.igd_fasta_read(self._f, i, start, n, buf))
        return buf.raw.decode("latin-1") if n else ""

    def close(self):
        f, self._f = getattr(self, "_f", None), None
        if f and getattr(self, "_lib", None) is not None:
            self._lib.igd_fasta_close(f)

    def __del__(self):
        try:
            self.close()
        except Exception:       # interpreter shutdown: the library may already be gone
            pass

    def __getitem__(self, rid):
        return self._views[rid]

    def __iter__(self):
        return iter(self.ids)

    def __len__(self):
        return len(self.ids)

    def __contains__(self, rid):
        return rid in self._views

    def keys(self):
        return list(self.ids)

    def items(self):
        return [(r, self._views[r]) for r in self.ids]

    def values(self):
        return [self._views[r] for r in self.ids]
