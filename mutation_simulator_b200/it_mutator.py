"""``ITMutator`` — interchromosomal translocations (it_mutator.py:20-220) on the B200 path.

Pairing of contigs is a host decision over at most a few thousand integers; breakpoints
are sampled on the GPU with the same position sampler as mutations
(sample_with_minimum_distance(1, len, n, 1), it_mutator.py:108-111); the alternating
own/partner concatenation is the splice kernel with raw far-copy records, and the output
FASTA image comes back complete.  Contigs without a partner are written once with normal
line wrapping (the reference's duplicated header is treated as a bug)."""
from __future__ import annotations

import random

import numpy as np

from .bedpe_writer import BedpeWriter, rows
from .engine import BUF_FASTA, Engine
from .fasta_writer import FastaWriter
from .mutator import run_seed
from .records import K_RAW, REC_DTYPE, T_IT
from .util import print_warning


class ITMutator:
    def __init__(self, args, fasta, sim):
        self._args, self._fasta, self._sim = args, fasta, sim
        self._fasta_writer = FastaWriter(args.outfastait)
        self._bedpe_writer = BedpeWriter(args.outbedpe)
        self._seed = run_seed(args)
        self._rng = random.Random(self._seed)
        # it_mutator.py:51-70: eligible contigs, then random disjoint pairs
        avail = [c.number for c in sim.chromosomes if c.it_rate is not None and len(fasta[c.number]) > 2]
        self._rng.shuffle(avail)
        self._partners = {}
        while len(avail) >= 2:
            a = avail.pop(0)
            b = avail.pop(self._rng.randrange(len(avail)))
            self._partners[a] = b
            self._partners[b] = a
        self._engine = None
        self.breakpoints = {}

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def close(self):
        self._fasta_writer.close()
        self._bedpe_writer.close()
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def _pairs_once(self):
        seen, out = set(), []
        for a, b in self._partners.items():
            if a not in seen:
                out.append((a, b))
                seen.update((a, b))
        return out

    def _generate_all_breakpoints(self, eng):
        fasta, sim, args = self._fasta, self._sim, self._args
        pairs, counts = [], []
        for a, b in self._pairs_once():
            la, lb = len(fasta[a]), len(fasta[b])
            n = int((la + lb - 4) / 2 * ((sim.chromosomes[a].it_rate + sim.chromosomes[b].it_rate) / 2))   # it_mutator.py:92
            ok = True
            if n > 0 and (n > la - n or n > lb - n or la - n <= 0 or lb - n <= 0):   # random.sample would raise (util.py:104)
                ok = False
                if not args.ignore_warnings:
                    print_warning(f"Interchromosomal translocation rate too high for sequence {a+1} and {b+1}.", args.no_color)
            if n <= 0 or not ok:
                if not args.ignore_warnings:
                    print_warning(f"No interchromosomal translocations could be generated between sequence {a+1} and {b+1} (it rates too low).",
                                  args.no_color)
                continue
            pairs.append((a, b))
            counts.append(n)
        bps = {}
        if pairs:
            bpa, bpb = eng.it_breakpoints(self._seed, [p[0] for p in pairs], [p[1] for p in pairs], counts)
            o = 0
            for (a, b), n in zip(pairs, counts):
                bps[a] = {"self": bpa[o:o + n], "partner": bpb[o:o + n]}
                bps[b] = {"self": bpb[o:o + n], "partner": bpa[o:o + n]}
                o += n
        return bps

    def _records(self, bps):
        fasta = self._fasta
        parts = []
        for c in sorted(bps):
            p = self._partners[c]
            a = np.concatenate(([0], bps[c]["self"].astype(np.int64), [len(fasta[c])]))
            b = np.concatenate(([0], bps[c]["partner"].astype(np.int64), [len(fasta[p])]))
            odd = np.arange(1, len(a) - 1, 2)          # intervals taken from the partner (it_mutator.py:133-137)
            r = np.zeros(len(odd), dtype=REC_DTYPE)
            r["pos"] = a[odd]
            r["cons"] = a[odd + 1] - a[odd]
            r["prod"] = b[odd + 1] - b[odd]
            r["src"] = fasta.goff[p] + b[odd]
            r["kind"] = K_RAW
            r["type"] = T_IT
            r["contig"] = c
            parts.append(r)
        return np.concatenate(parts) if parts else np.zeros(0, dtype=REC_DTYPE)

    def mutate(self):
        """Creates interchromosomal translocations and writes them to a Fasta and BEDPE file."""
        fasta = self._fasta
        eng = self._engine = Engine(getattr(self._args, "device", 0))
        fasta.upload(eng)
        bps = self.breakpoints = self._generate_all_breakpoints(eng)
        eng.load_records(self._records(bps))
        eng.apply()
        self._fasta_writer.write_image(eng.download(BUF_FASTA))
        for chrom in self._sim.chromosomes:            # FASTA order, one block of rows per paired contig
            c = chrom.number
            if c in bps:
                p = self._partners[c]
                self._bedpe_writer._f.write(rows(fasta[c].name, bps[c]["self"], len(fasta[c]), fasta[p].name,
                                                 bps[c]["partner"], len(fasta[p])))
