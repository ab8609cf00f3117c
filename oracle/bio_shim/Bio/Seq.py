"""Bio.Seq stand-in (test infrastructure): an immutable string wrapper with the
methods the reference calls -- slicing, str(), len(), upper(),
reverse_complement() (IUPAC-aware, case preserving) and translate()."""

_COMP = str.maketrans("ACGTMRWSYKVHDBXNacgtmrwsykvhdbxnUu", "TGCAKYWSRMBDHVXNtgcakywsrmbdhvxnAa")

_BASES = "TCAG"
_AAS = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
_CODON = {a + b + c: _AAS[16 * i + 4 * j + k]
          for i, a in enumerate(_BASES) for j, b in enumerate(_BASES) for k, c in enumerate(_BASES)}
_AMBIG = {"A": "A", "C": "C", "G": "G", "T": "T", "U": "T", "M": "AC", "R": "AG", "W": "AT", "S": "CG",
          "Y": "CT", "K": "GT", "V": "ACG", "H": "ACT", "D": "AGT", "B": "CGT", "X": "ACGT", "N": "ACGT"}


def _translate_codon(codon):
    aa = _CODON.get(codon)
    if aa is not None:
        return aa
    try:
        options = [_AMBIG[c] for c in codon]
    except KeyError:
        raise ValueError("Codon '%s' is invalid" % codon)
    aas = {_CODON[a + b + c] for a in options[0] for b in options[1] for c in options[2]}
    if len(aas) == 1:
        return aas.pop()
    return "X"


class Seq:
    __slots__ = ("_data",)

    def __init__(self, data=""):
        self._data = data._data if isinstance(data, Seq) else str(data)

    def __str__(self):
        return self._data

    def __repr__(self):
        return "Seq(%r)" % self._data

    def __len__(self):
        return len(self._data)

    def __bytes__(self):
        return self._data.encode("ascii")

    def __hash__(self):
        return hash(self._data)

    def __eq__(self, other):
        return self._data == (other._data if isinstance(other, Seq) else other)

    def __iter__(self):
        return iter(self._data)

    def __getitem__(self, index):
        if isinstance(index, slice):
            return Seq(self._data[index])
        return self._data[index]

    def __add__(self, other):
        return Seq(self._data + str(other))

    def __contains__(self, item):
        return str(item) in self._data

    def find(self, sub, *args):
        return self._data.find(str(sub), *args)

    def upper(self):
        return Seq(self._data.upper())

    def lower(self):
        return Seq(self._data.lower())

    def complement(self):
        return Seq(self._data.translate(_COMP))

    def reverse_complement(self):
        return Seq(self._data.translate(_COMP)[::-1])

    def translate(self):
        s = self._data.upper().replace("U", "T")
        n = len(s) - len(s) % 3
        return Seq("".join(_translate_codon(s[i:i + 3]) for i in range(0, n, 3)))
