"""In-memory stand-in for the 12 ``pyfaidx`` calls the reference makes.

TEST INFRASTRUCTURE ONLY.  ``pyfaidx`` is a third-party dependency of the
reference (pyproject.toml:30-34) that is not installed in the build container
and cannot be fetched (no network).  It does no arithmetic on the hot path; it
only serves contig names, lengths, the line width and bases.  This module lets
the *unmodified* reference under /root/reference run here so that
``tests/golden/make_golden.py`` can produce golden vectors.  It is never
imported by the product package.

Surface (call sites: util.py:84-88, mutator.py:119,133-139,337-415,423,
it_mutator.py:56,137-151,162-211, rmt.py:315-426, vcf_writer.py:88-90):
``Fasta(path, one_based_attributes=False, as_raw=True,
sequence_always_upper=True, read_ahead=...)``; ``fasta[int|str]``; ``keys()``;
``get_seq(name, start1, end1)``; ``faidx.index[name].lenc``; ``close()``;
record ``len()``, ``[int]``, ``[a:b]``, ``.name``, ``.long_name``, ``iter``.
"""
from __future__ import annotations


class FastaIndexingError(Exception):
    pass


class FastaNotFoundError(Exception):
    pass


class _IndexEntry:
    __slots__ = ("rlen", "lenc", "lenb")

    def __init__(self, rlen, lenc, lenb):
        self.rlen, self.lenc, self.lenb = rlen, lenc, lenb


class _Faidx:
    def __init__(self):
        self.index = {}


class FastaRecord:
    def __init__(self, name, long_name, seq, lenc):
        self.name = name
        self.long_name = long_name
        self._seq = seq
        self._lenc = lenc

    def __len__(self):
        return len(self._seq)

    def __getitem__(self, n):
        if isinstance(n, slice):
            start, stop, step = n.start, n.stop, n.step
            if not start:
                start = 0
            if not stop:  # pyfaidx quirk: a falsy stop means "to the end"
                stop = len(self)
            if stop < 0:
                stop = len(self) + stop
            if start < 0:
                start = len(self) + start
            return self._seq[start:stop][::step]
        if n < 0:
            n = len(self) + n
        return self._seq[n]

    def __iter__(self):
        # pyfaidx iterates a record line by line
        w = self._lenc if self._lenc > 0 else max(1, len(self._seq))
        for i in range(0, len(self._seq), w):
            yield self._seq[i:i + w]

    def __str__(self):
        return self._seq


class Fasta:
    def __init__(self, filename, one_based_attributes=False, as_raw=True,
                 sequence_always_upper=True, read_ahead=None, **_):
        try:
            with open(filename, "r") as fh:
                text = fh.read()
        except FileNotFoundError:
            raise FastaNotFoundError(f"Cannot read FASTA from file {filename}")
        self.filename = filename
        self.faidx = _Faidx()
        self._records = {}
        self._order = []
        name = long_name = None
        lines = []

        def flush():
            if name is None:
                return
            if name in self._records:
                raise ValueError(f"Duplicate key \"{name}\"")
            widths = [len(l) for l in lines]
            # pyfaidx rejects ragged records (all but the last line equal)
            if len(widths) > 1 and any(w != widths[0] for w in widths[:-1]):
                raise FastaIndexingError(
                    f"Line length of fasta file is not consistent! ({name})")
            if len(widths) > 1 and widths[-1] > widths[0]:
                raise FastaIndexingError(
                    f"Line length of fasta file is not consistent! ({name})")
            seq = "".join(lines)
            if sequence_always_upper:
                seq = seq.upper()
            lenc = widths[0] if widths else 0
            self._records[name] = FastaRecord(name, long_name, seq, lenc)
            self._order.append(name)
            self.faidx.index[name] = _IndexEntry(len(seq), lenc, lenc + 1)

        for raw in text.split("\n"):
            line = raw.rstrip("\r")
            if line.startswith(">"):
                flush()
                long_name = line[1:]
                name = long_name.split()[0] if long_name.split() else ""
                lines = []
            elif name is not None:
                if line:
                    lines.append(line)
        flush()
        if not self._order:
            raise FastaIndexingError(f"No sequences in {filename}")

    def keys(self):
        return list(self._order)

    def __len__(self):
        return len(self._order)

    def __iter__(self):
        for k in self._order:
            yield self._records[k]

    def __contains__(self, k):
        return k in self._records

    def __getitem__(self, key):
        if isinstance(key, int):
            return self._records[self._order[key]]
        return self._records[key]

    def get_seq(self, name, start, end, rc=False):
        # 1-based inclusive, as_raw=True -> str
        return self._records[name]._seq[start - 1:end]

    def close(self):
        pass
