"""Synthetic StrainCall workloads: reads drawn from simulated 16S strains.

BASELINE.json's configs are all "synthetic reads drawn from simulated 16S strains
(scripts/ecoli_mg1655_16S.fasta, mutated)".  This module makes them, in the form the
reference holds them in memory right after ``load_mapping_reads`` (StrainCall.cpp:480-670):

  * unique ``AlignRead`` tuples (relative_pos, cigar, seq, "", copy_number), ordered the way
    ``std::map<AlignRead,...>`` orders them (StrainCall.cpp:532,594-614);
  * ``ReadPairs`` uid -> [mate uid or -1, one entry per duplicate copy], built in read-NAME
    order like StrainCall.cpp:629-665;
  * depth down-sampling with mt19937(1234) + uniform_real_distribution (StrainCall.cpp:491-493,
    528-529,589), reproduced bit-for-bit from the raw MT19937 stream.

The alignment of every read to the window's gene sequence is known by construction (strains
are edits of the gene), so no aligner is needed; the CIGARs use only M/I/D like the records
StrainCall keeps.  numpy only -- no torch, no CUDA.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
BASES = "ACGT"


def ecoli_16s() -> str:
    """The 1542 bp E. coli MG1655 16S gene (ungapped form of scripts/ecoli_mg1655_16S.fasta)."""
    with open(os.path.join(_DATA, "ecoli_mg1655_16S.ungapped.fa")) as f:
        return "".join(line.strip() for line in f if not line.startswith(">"))


# --------------------------------------------------------------------------------------
# std::mt19937 + std::uniform_real_distribution<double>(0,1) as libstdc++ draws it
# --------------------------------------------------------------------------------------
class StdMt19937:
    """Raw std::mt19937 stream (seeded like ``std::mt19937 gen(seed)``)."""

    def __init__(self, seed: int = 1234):
        self._rs = np.random.RandomState(seed)  # init_genrand(seed): same recurrence as std::mt19937

    def raw(self, n: int) -> np.ndarray:
        # RandomState.randint draws on the full uint32 range consume exactly one 32-bit output each.
        return self._rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint64)

    def canonical(self, n: int) -> np.ndarray:
        """n draws of std::generate_canonical<double,53>: two 32-bit words, low word first."""
        w = self.raw(2 * n)
        lo = w[0::2].astype(np.float64)
        hi = w[1::2].astype(np.float64)
        s = (lo + hi * 4294967296.0) / 18446744073709551616.0
        s[s >= 1.0] = np.nextafter(1.0, 0.0)
        return s


# --------------------------------------------------------------------------------------
@dataclass
class StrainTruth:
    """A simulated strain as an edit script against the gene (column-wise)."""
    seq: str
    # for every base of ``seq``: gene coordinate it is aligned to, or -1 for an inserted base
    gene_pos: np.ndarray
    abundance: float


@dataclass
class Subgroup:
    """One StrainCall problem: a gene window plus its de-duplicated aligned reads."""
    gene: str
    pos: List[int]
    cigar: List[str]
    seq: List[str]
    cn: List[int]
    pair_off: np.ndarray  # CSR over unique reads
    pair_val: np.ndarray
    n_raw_reads: int  # reads before de-duplication / down-sampling (the metric's unit)
    truth: List[StrainTruth] = field(default_factory=list)

    @property
    def n_unique(self) -> int:
        return len(self.pos)

    def packed(self) -> dict:
        """The reads as host buffers: positions and copies as int32 arrays, CIGARs and letters as byte arenas with
        int64 offsets (what rambl_batch_add_subgroup_packed takes).  Built once and kept."""
        pk = self.__dict__.get("_packed")
        if pk is None:
            def arena(strs):
                off = np.zeros(len(strs) + 1, dtype=np.int64)
                if strs:
                    off[1:] = np.cumsum([len(x) for x in strs])
                return off, "".join(strs).encode()
            co, cc = arena(self.cigar)
            so, sc = arena(self.seq)
            pk = dict(pos=np.asarray(self.pos, dtype=np.int32), cn=np.asarray(self.cn, dtype=np.int32),
                      cigar_off=co, cigar_chars=cc, seq_off=so, seq_chars=sc)
            self.__dict__["_packed"] = pk
        return pk

    @property
    def n_reads(self) -> int:
        return int(sum(self.cn))


def _mutate(gene: str, rng: np.random.Generator, divergence: float, indel_frac: float,
            homopolymer_bias: bool) -> Tuple[str, np.ndarray]:
    """Return (strain sequence, gene coordinate per strain base or -1)."""
    out: List[str] = []
    gpos: List[int] = []
    L = len(gene)
    i = 0
    while i < L:
        g = gene[i]
        r = rng.random()
        in_hp = homopolymer_bias and i > 0 and gene[i - 1] == g
        p_edit = divergence * (3.0 if in_hp else 1.0)
        # never edit the first/last 2 columns so that every read starts and ends in a match op
        if 2 <= i < L - 2 and r < p_edit:
            kind = rng.random()
            if kind >= indel_frac:  # substitution
                alt = BASES.replace(g, "")[int(rng.integers(3))]
                out.append(alt)
                gpos.append(i)
            elif kind < indel_frac / 2:  # deletion of 1-2 gene bases
                n = 1 + int(rng.integers(2))
                n = min(n, L - 2 - i)
                i += n
                continue
            else:  # insertion of 1-3 bases after this gene base
                out.append(g)
                gpos.append(i)
                n = 1 + int(rng.integers(3))
                for _ in range(n):
                    out.append(g if in_hp else BASES[int(rng.integers(4))])
                    gpos.append(-1)
        else:
            out.append(g)
            gpos.append(i)
        i += 1
    return "".join(out), np.asarray(gpos, dtype=np.int64)


def _cigar_from_ops(ops: Sequence[str]) -> str:
    if not ops:
        return ""
    out = []
    prev = ops[0]
    n = 1
    for o in ops[1:]:
        if o == prev:
            n += 1
        else:
            out.append(f"{n}{prev}")
            prev, n = o, 1
    out.append(f"{n}{prev}")
    return "".join(out)


def _emit_read(strain: StrainTruth, start: int, length: int, rng: np.random.Generator, sub_err: float,
               indel_err: float, max_ins: int) -> Optional[Tuple[int, str, str]]:
    """Sequence a read from ``strain`` and align it to the gene by construction.

    Returns (0-based gene position of the first aligned base, cigar, read sequence)."""
    s = strain.seq
    gp = strain.gene_pos
    end = min(len(s), start + length)
    # trim so that the read starts and ends on a base aligned to the gene
    while start < end and gp[start] < 0:
        start += 1
    while end > start and gp[end - 1] < 0:
        end -= 1
    if end - start < 20:
        return None
    ops: List[str] = []
    bases: List[str] = []
    prev_g = int(gp[start]) - 1
    for k in range(start, end):
        g = int(gp[k])
        interior = start + 1 < k < end - 2
        if g < 0:
            ops.append("I")
            bases.append(s[k])
            continue
        if g > prev_g + 1:
            ops.extend("D" * (g - prev_g - 1))
        prev_g = g
        b = s[k]
        if interior and indel_err > 0.0:
            r = rng.random()
            hp = s[k - 1] == b
            p = indel_err * (4.0 if hp else 0.25)
            if r < p / 2 and ops and ops[-1] == "M":  # read lost this base
                ops.append("D")
                continue
            if r < p:  # read gained a copy of this base
                ops.append("M")
                bases.append(b)
                ops.append("I")
                bases.append(b)
                continue
        if rng.random() < sub_err:
            b = BASES.replace(b, "")[int(rng.integers(3))]
        ops.append("M")
        bases.append(b)
    cigar = _cigar_from_ops(ops)
    # the same filter StrainCall applies (max_insert_size < max_ins, StrainCall.cpp:583-586)
    run = 0
    for o in ops:
        run = run + 1 if o == "I" else 0
        if run >= max_ins:
            return None
    return int(gp[start]), cigar, "".join(bases)


def simulate_raw_reads(n_reads: int = 2000, read_len: int = 100, n_strains: int = 3, divergence=(0.01, 0.03),
                  sub_err: float = 0.005, indel_err: float = 0.0, indel_frac: float = 0.1,
                  homopolymer_bias: bool = False, paired: bool = False, max_depth: int = 800,
                  gene: Optional[str] = None, window: Optional[Tuple[int, int]] = None, seed: int = 0,
                  max_ins: int = 10, abundances: Optional[Sequence[float]] = None) -> Tuple[str, List[Tuple[str, int, str, str]], List[StrainTruth]]:
    """Simulate the raw mapped reads of one subgroup: (gene window, [(name, pos, cigar, seq)], strains).
    Reads are in sequencing order; names of mates end in /1 and /2 (the suffix StrainCall derives from the
    SAM flag, StrainCall.cpp:552-562).  max_depth and abundances are accepted for a uniform signature."""
    rng = np.random.default_rng(seed)
    gene = ecoli_16s() if gene is None else gene
    if window is not None:
        gene = gene[window[0]:window[1]]
    L = len(gene)
    if abundances is None:
        ab = rng.dirichlet(np.full(n_strains, 2.0)) * 0.8 + 0.2 / n_strains
    else:
        ab = np.asarray(abundances, dtype=np.float64)
        ab = ab / ab.sum()
    strains: List[StrainTruth] = []
    for k in range(n_strains):
        d = divergence[0] + (divergence[1] - divergence[0]) * rng.random()
        if k == 0:
            d = 0.0  # the seed gene itself is one of the strains
        sseq, gp = _mutate(gene, rng, d, indel_frac, homopolymer_bias)
        strains.append(StrainTruth(sseq, gp, float(ab[k])))

    raw: List[Tuple[str, int, str, str]] = []  # (name, pos, cigar, seq)
    which = rng.choice(n_strains, size=n_reads, p=ab)
    rid = 0
    while rid < n_reads:
        st = strains[int(which[rid])]
        if paired and rid + 1 < n_reads:
            frag = int(rng.integers(read_len + 20, max(read_len + 21, min(3 * read_len, len(st.seq)))))
            a = int(rng.integers(0, max(1, len(st.seq) - frag + 1)))
            r1 = _emit_read(st, a, read_len, rng, sub_err, indel_err, max_ins)
            r2 = _emit_read(st, max(a, a + frag - read_len), read_len, rng, sub_err, indel_err, max_ins)
            nm = f"r{rid // 2:07d}"
            if r1 is not None:
                raw.append((nm + "/1",) + r1)
            if r2 is not None:
                raw.append((nm + "/2",) + r2)
            rid += 2
        else:
            a = int(rng.integers(0, max(1, len(st.seq) - read_len + 1)))
            r = _emit_read(st, a, read_len, rng, sub_err, indel_err, max_ins)
            if r is not None:
                raw.append((f"s{rid:07d}",) + r)
            rid += 1

    return gene, raw, strains


def make_subgroup(n_reads: int = 2000, read_len: int = 100, n_strains: int = 3, divergence=(0.01, 0.03),
                  sub_err: float = 0.005, indel_err: float = 0.0, indel_frac: float = 0.1,
                  homopolymer_bias: bool = False, paired: bool = False, max_depth: int = 800,
                  gene: Optional[str] = None, window: Optional[Tuple[int, int]] = None, seed: int = 0,
                  max_ins: int = 10, abundances: Optional[Sequence[float]] = None) -> Subgroup:
    """Simulate one taxonomic subgroup (a seed gene window and the reads mapped to it), in the form
    StrainCall holds it after load_mapping_reads.

    ``window`` is a 0-based half-open slice of the gene (default: the whole gene, the way
    scripts/rambl.py:181-187 runs StrainCall with ``-w 5000``)."""
    gene, raw, strains = simulate_raw_reads(n_reads, read_len, n_strains, divergence, sub_err, indel_err, indel_frac,
                                            homopolymer_bias, paired, max_depth, gene, window, seed, max_ins, abundances)
    L = len(gene)
    # ---- depth down-sampling, StrainCall.cpp:505-529,586-592 (window = whole gene here)
    depth = 0
    for (_, p, cg, _s) in raw:
        ref_len = 0
        num = ""
        for ch in cg:
            if ch.isdigit():
                num += ch
            else:
                if ch in "MD":
                    ref_len += int(num)
                num = ""
        depth += ref_len
    depth //= max(L, 1)
    rho = min(1.0, max_depth / (depth + 0.0)) if depth > 0 else 1.0
    # StrainCall meets the reads in the order of the position-sorted BAM
    raw = sorted(raw, key=lambda r: r[1])
    u = StdMt19937(1234).canonical(len(raw))
    kept = [r for r, x in zip(raw, u) if not (x > rho)]

    # ---- de-duplication into std::map<AlignRead, vector<name>> order
    groups = {}
    for (nm, p, cg, sq) in kept:
        groups.setdefault((p, cg, sq), []).append(nm)
    keys = sorted(groups.keys())
    pos = [k[0] for k in keys]
    cigar = [k[1] for k in keys]
    seq = [k[2] for k in keys]
    cn = [len(groups[k]) for k in keys]
    name_uid = {}
    for uid, k in enumerate(keys):
        for nm in groups[k]:
            name_uid[nm] = uid
    pairs: List[List[int]] = [[] for _ in keys]
    for nm in sorted(name_uid.keys()):  # std::map<string,int> iteration order
        uid = name_uid[nm]
        mate = -1
        if nm.endswith("/1"):
            mate = name_uid.get(nm[:-2] + "/2", -1)
        elif nm.endswith("/2"):
            mate = name_uid.get(nm[:-2] + "/1", -1)
        pairs[uid].append(mate)
    off = np.zeros(len(keys) + 1, dtype=np.int32)
    for i, pl in enumerate(pairs):
        off[i + 1] = off[i] + len(pl)
    val = np.asarray([m for pl in pairs for m in pl], dtype=np.int32)
    return Subgroup(gene, pos, cigar, seq, cn, off, val, len(raw), strains)


def write_cli_fixture(dirname: str, gene_name: str, gene: str, raw) -> Tuple[str, str]:
    """Write what the StrainCall CLI reads: <dir>/genes.fa (+ .fai) and <dir>/reads.sam -- SAM text records
    sorted by position, standing in for the BAM (tests/samtools_shim/samtools serves them).
    Returns (fasta path, sam path)."""
    os.makedirs(dirname, exist_ok=True)
    fa = os.path.join(dirname, "genes.fa")
    with open(fa, "w") as f:
        f.write(">%s\n" % gene_name)
        for i in range(0, len(gene), 60):
            f.write(gene[i:i + 60] + "\n")
    with open(fa + ".fai", "w") as f:
        f.write("%s\t%d\t%d\t60\t61\n" % (gene_name, len(gene), len(gene_name) + 2))
    sam = os.path.join(dirname, "reads.sam")
    recs = []
    for (nm, p, cg, sq) in raw:
        flag, base = 0, nm
        if nm.endswith("/1"):
            flag, base = 65, nm[:-2]
        elif nm.endswith("/2"):
            flag, base = 129, nm[:-2]
        recs.append((p, "%s\t%d\t%s\t%d\t30\t%s\t*\t0\t0\t%s\t%s" % (base, flag, gene_name, p + 1, cg, sq, "I" * len(sq))))
    recs.sort(key=lambda r: r[0])
    with open(sam, "w") as f:
        for _, line in recs:
            f.write(line + "\n")
    return fa, sam


# Named workloads of BASELINE.json["configs"] ------------------------------------------------
def config2_subgroup(k: int, window: Optional[Tuple[int, int]] = None) -> Subgroup:
    """Subgroup k of BASELINE configs[2] ("500 synthetic taxonomic subgroups x 5k reads"): 5 000 150 bp reads from
    2-6 strains of the 16S gene.  With ``window`` the same subgroup restricted to a slice of the gene at the same
    depth (the read count shrinks with the slice) -- the bounded samples the CPU reference is timed on."""
    if window is None:
        return make_subgroup(5000, 150, 2 + (k % 5), seed=k)
    w = window[1] - window[0]
    return make_subgroup(max(50, int(5000 * w / 1542)), min(150, w), 2 + (k % 5), seed=k, window=window)


def config_workload(idx: int, seed: int = 0, scale: float = 1.0) -> List[Subgroup]:
    """BASELINE.json configs[idx] as a list of subgroups (scale<1 shrinks read counts for tests)."""
    if idx == 0:  # 2k 100bp reads from 3 mutated strains (CPU-runnable)
        return [make_subgroup(int(2000 * scale), 100, 3, seed=seed)]
    if idx == 1:  # single subgroup, 20k 150bp reads, 10 strains at 1-3% divergence
        return [make_subgroup(int(20000 * scale), 150, 10, divergence=(0.01, 0.03), seed=seed)]
    if idx == 2:  # 500 subgroups x 5k reads
        n = max(1, int(500 * scale))
        return [config2_subgroup(seed * 100003 + k) for k in range(n)]
    if idx == 3:  # one deep subgroup, 1M 150bp reads, 50 strains (down-sampled to -D like StrainCall)
        return [make_subgroup(int(1000000 * scale), 150, 50, divergence=(0.01, 0.03), seed=seed)]
    if idx == 4:  # 250bp MiSeq-like reads, indel-rich strains (homopolymer-biased indels); see tools/ref_config4.py for what
        # the reference does when the READS carry indel errors as well (it does not finish)
        return [make_subgroup(int(5000 * scale), 250, 4, indel_err=0.0, indel_frac=0.4,
                              homopolymer_bias=True, seed=seed)]
    raise ValueError(idx)
