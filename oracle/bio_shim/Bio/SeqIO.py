"""Bio.SeqIO stand-in (test infrastructure): parse(handle_or_path, "fasta")."""
from .Seq import Seq


class SeqRecord:
    def __init__(self, seq, id="", description=""):
        self.seq = seq
        self.id = id
        self.name = id
        self.description = description


def parse(handle, fmt):
    if fmt != "fasta":
        raise ValueError("shim supports fasta only")
    close = False
    if isinstance(handle, (str, bytes)) or hasattr(handle, "__fspath__"):
        handle = open(handle)
        close = True
    try:
        title = None
        chunks = []
        for line in handle:
            if line.startswith(">"):
                if title is not None:
                    yield _record(title, chunks)
                title = line[1:].rstrip()
                chunks = []
            elif title is not None:
                chunks.append("".join(line.split()))
        if title is not None:
            yield _record(title, chunks)
    finally:
        if close:
            handle.close()


def _record(title, chunks):
    first = title.split(None, 1)[0] if title.strip() else ""
    return SeqRecord(Seq("".join(chunks)), id=first, description=title)
