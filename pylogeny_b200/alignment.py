"""Minimal alignment reader standing in for the reference's p4-backed alignment.py.

Only what the scoring hot path touches is provided:
`getSize/getNumSeqs/getTaxa/getDataType/getSequenceString/data.sequenceSlice`, the
PHYLIP temp file handed to libpll (alignment.py:242-251, 278-289) and the
`phylipFriendlyAlignment` renaming to T0..Tn-1 (alignment.py:264-276).  FastTree,
NEXUS and the discrete-state model are out of scope.
"""
import os
from tempfile import NamedTemporaryFile as NTempFile

import numpy as np

_DNA = set("ACGTUNRYKMSWBDHV-?XO.")


class _Data(object):
    """What the callers use of p4.Alignment: taxNames, sequences, sequenceSlice, dataType, dim."""

    class _Seq(object):
        def __init__(self, name, sequence):
            self.name, self.sequence = name, sequence

    def __init__(self, names, seqs):
        self.taxNames = list(names)
        self.sequences = [self._Seq(n, s) for n, s in zip(names, seqs)]
        letters = set("".join(seqs).upper())
        self.dataType = 'dna' if letters <= _DNA else 'protein'
        self.dim = 4 if self.dataType == 'dna' else 20

    def sequenceSlice(self, site):
        return [s.sequence[site] for s in self.sequences]

    def __len__(self):
        return len(self.sequences[0].sequence) if self.sequences else 0


def read_fasta(path):
    names, seqs = [], []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line[0] == '>':
                names.append(line[1:].split()[0])
                seqs.append([])
            else:
                seqs[-1].append(line.replace(' ', ''))
    return names, ["".join(s) for s in seqs]


def read_phylip(path):
    """Sequential or interleaved PHYLIP, names separated from data by whitespace."""
    with open(path) as fh:
        lines = [l.rstrip("\n") for l in fh if l.strip()]
    n, m = [int(x) for x in lines[0].split()[:2]]
    names, seqs = [], []
    for i in range(n):
        parts = lines[1 + i].split(None, 1)
        names.append(parts[0])
        seqs.append(parts[1].replace(' ', '') if len(parts) > 1 else '')
    k = 1 + n
    while any(len(s) < m for s in seqs) and k < len(lines):
        for i in range(n):
            if k >= len(lines):
                break
            seqs[i] += lines[k].replace(' ', '')
            k += 1
    if any(len(s) != m for s in seqs):
        raise IOError("Alignment file could not be parsed.")
    return names, seqs


class alignment(object):
    """A biological sequence alignment (FASTA or PHYLIP file, or names+sequences)."""

    def __init__(self, inal=None, names=None, seqs=None):
        if inal is not None:
            with open(inal) as fh:
                head = fh.read(1)
            names, seqs = read_fasta(inal) if head == '>' else read_phylip(inal)
        if names is None:
            names, seqs = [], []
        if len(set(len(s) for s in seqs)) > 1:
            raise IOError("Alignment file could not be parsed.")
        self.data = _Data(names, seqs)
        self.paths = {}
        self._matrix = None

    def __getitem__(self, i):
        return self.data.sequences[i]

    def __len__(self):
        return len(self.data)

    def __iter__(self):
        for i in range(self.getNumSeqs()):
            yield self[i]

    def __str__(self):
        return '\n'.join(self.toStrList())

    def toStrList(self):
        return [self.getSequenceString(i) for i in range(self.getNumSeqs())]

    def getDataType(self):
        return self.data.dataType

    def getSize(self):
        return len(self)

    def getNumSeqs(self):
        return len(self.data.sequences)

    def getDim(self):
        return self.data.dim

    def getSequenceString(self, i):
        return self[i].sequence

    def getTaxa(self):
        return self.data.taxNames

    def as_matrix(self):
        """uint8 [n_rows, n_sites] of the raw characters (vectorised consumers)."""
        if self._matrix is None:
            n, m = self.getNumSeqs(), self.getSize()
            self._matrix = np.frombuffer("".join(self.toStrList()).encode("latin-1"), dtype=np.uint8).reshape(n, m)
        return self._matrix

    def _makePhylipFile(self):
        if 'phylip' in self.paths:
            return
        fh = NTempFile(mode='w', delete=False, suffix='.phy')
        self.paths['phylip'] = fh.name
        fh.write(" %d %d\n" % (self.getNumSeqs(), self.getSize()))
        for s in self.data.sequences:
            fh.write("%s %s\n" % (s.name, s.sequence))
        fh.close()

    def getPhylip(self):
        self._makePhylipFile()
        return self.paths['phylip']

    def close(self):
        for p in self.paths.values():
            if os.path.isfile(p):
                os.unlink(p)
        self.paths = {}


class phylipFriendlyAlignment(alignment):
    """Renames all taxa to T0..Tn-1 so a strict PHYLIP file can be written (alignment.py:223-276)."""

    def __init__(self, inal=None, names=None, seqs=None):
        super(phylipFriendlyAlignment, self).__init__(inal, names, seqs)
        self.namedict = {}
        for ind, item in enumerate(self.data.sequences):
            name = 'T' + str(ind)
            self.namedict[name] = item.name
            item.name = name
            self.data.taxNames[ind] = name

    def getProperName(self, n):
        return self.namedict.get(n)

    def getTaxa(self):
        # alignment.py:371-379 returns the dict's keys: insertion (= row) order on Python 3
        return [x for x in self.namedict]
