"""Host-side mirror of StrainCall's in-memory interface over librambl_b200.so (include/rambl_b200.h).

The names follow the reference (StrainCall/PartialOrderGraph.hpp:231-357,
MultipleSequenceAlignment.hpp:87-107):

    pog = PartialOrderGraph(G, R)                  # R: AlignRead tuples (pos, cigar, seq, qual, copies)
    pog.output_edge()                              # the text StrainCall -G prints
    strains = pog.infer_strains(read_pairs, n=5000, e=0.01, tau=0.02, diff=0.01)
    pog.read_assign(strains, R, read_pairs, n=5000)
    MultipleSequenceAlignmentSP().align(seqs)      # rows = MSA::get(t)

plus ``StrainCallBatch`` for many subgroups at once (what scripts/rambl.py does with a process
pool, rambl.py:165-194).  Everything computes on the GPU through the C ABI; there is no CPU
fallback -- without the built library or without a device the calls raise ``RamblError``.
This module imports only ctypes and numpy.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librambl_b200.so")

RAMBL_OK, RAMBL_ERR_CUDA, RAMBL_ERR_INVALID, RAMBL_ERR_CAPACITY, RAMBL_ERR_NO_STRAINS, RAMBL_ERR_STATE = range(6)


class RamblError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("rambl_b200 error %d: %s" % (code, msg))
        self.code = code


class Stats(C.Structure):
    _fields_ = [("gpu_launches", C.c_int32), ("level_steps", C.c_int32), ("draws", C.c_int64),
                ("loglik_updates", C.c_int64), ("msa_dp_cells", C.c_int64), ("msa_problems", C.c_int32),
                ("msa_kernel_ms", C.c_float), ("infer_gpu_ms", C.c_float), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("gibbs_kernel_ms", C.c_float), ("gibbs_launches", C.c_int32),
                ("gibbs_alg_bytes", C.c_int64), ("gibbs_rounds", C.c_int64), ("gibbs_passes", C.c_int64),
                ("dpm_kernel_ms", C.c_float), ("dpm_launches", C.c_int32), ("dpm_alg_bytes", C.c_int64),
                ("offtable_levels", C.c_int32), ("reserved", C.c_int32)]


_lib: Optional[C.CDLL] = None
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_strp = C.POINTER(C.c_char_p)

# every symbol include/rambl_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "rambl_last_error": (C.c_char_p, []),
    "rambl_device_count": (C.c_int, []),
    "rambl_set_device": (C.c_int, [C.c_int32]),
    "rambl_free": (None, [C.c_void_p]),
    "rambl_release_cached_memory": (None, []),
    "rambl_cached_host_bytes": (C.c_int64, []),
    "rambl_solve_layout": (C.c_int32, [C.c_int32, C.c_int32, _i32p, C.c_int32]),
    "rambl_batch_walk_plan": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]),
    "rambl_set_gibbs_blocks": (C.c_int, [C.c_int32]),
    "rambl_set_host_threads": (C.c_int, [C.c_int32]),
    "rambl_set_walk_mode": (C.c_int, [C.c_int32]),
    "rambl_set_walk_blocks": (C.c_int, [C.c_int32]),
    "rambl_set_walk_cluster": (C.c_int, [C.c_int32]),
    "rambl_msa_rows_capacity": (C.c_int64, [C.c_int32, _i32p, _i32p]),
    "rambl_msa_sp_align_batch": (C.c_int, [C.c_int32, _i32p, _i32p, C.c_char_p, _i32p, _i64p, _i32p, C.c_char_p,
                                           C.POINTER(C.c_uint64), C.POINTER(C.c_float)]),
    "rambl_batch_create": (C.c_void_p, []),
    "rambl_batch_destroy": (None, [C.c_void_p]),
    "rambl_batch_add_subgroup": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32, _i32p, _strp, _strp, _i32p, _i32p, _i32p]),
    "rambl_batch_add_subgroup_packed": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32, _i32p, _i64p, C.c_char_p, _i64p, C.c_char_p,
                                                  _i32p, _i32p, _i32p]),
    "rambl_batch_add_graph": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint8), _i32p, C.c_char_p, _i32p, _i32p,
                                        _i32p, _i32p, _i32p, _i32p, C.c_char_p, _i32p, _i32p, _i32p]),
    "rambl_batch_build_graphs": (C.c_int, [C.c_void_p]),
    "rambl_batch_thread_reads": (C.c_int, [C.c_void_p]),
    "rambl_batch_msa_problems_text": (C.c_void_p, [C.c_void_p]),
    "rambl_batch_finish_graphs_with_rows": (C.c_int, [C.c_void_p, C.c_char_p]),
    "rambl_batch_infer": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int32]),
    "rambl_batch_solve": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int32]),
    "rambl_batch_num_subgroups": (C.c_int32, [C.c_void_p]),
    "rambl_batch_num_nodes": (C.c_int32, [C.c_void_p, C.c_int32]),
    "rambl_batch_graph_text": (C.c_void_p, [C.c_void_p, C.c_int32, C.c_int32]),
    "rambl_batch_status": (C.c_int32, [C.c_void_p, C.c_int32]),
    "rambl_batch_num_strains": (C.c_int32, [C.c_void_p, C.c_int32]),
    "rambl_batch_strain": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _f64p, _f64p, _i32p]),
    "rambl_batch_strain_path": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _i32p]),
    "rambl_batch_strain_sequence": (C.c_void_p, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "rambl_batch_strain_sub": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _f64p]),
    "rambl_batch_strain_loglik": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _f64p, C.c_int32]),
    "rambl_batch_order": (C.c_int, [C.c_void_p, C.c_int32, _i32p]),
    "rambl_batch_strains_text": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "rambl_batch_fasta": (C.c_void_p, [C.c_void_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32, C.c_float]),
    "rambl_batch_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
}


def lib() -> C.CDLL:
    """Load librambl_b200.so (built in-tree by __graft_entry__.build()); fail loudly when absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RamblError(RAMBL_ERR_CUDA, "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                             "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc: int):
    if rc != RAMBL_OK:
        raise RamblError(rc, (lib().rambl_last_error() or b"").decode())


def _take(p) -> str:
    if not p:
        raise RamblError(RAMBL_ERR_INVALID, (lib().rambl_last_error() or b"").decode())
    s = C.string_at(p).decode()
    lib().rambl_free(p)
    return s


def _strs(xs: Sequence[str]):
    arr = (C.c_char_p * max(1, len(xs)))()
    for i, x in enumerate(xs):
        arr[i] = x.encode()
    return arr


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def device_count() -> int:
    return lib().rambl_device_count()


def release_cached_memory() -> None:
    """Cached device, pinned and pageable host blocks go back to the driver / the allocator."""
    lib().rambl_release_cached_memory()


def solve_layout(n_subgroups: int, sms: int = 148) -> List[int]:
    """Chunk bounds rambl_batch_solve uses for a batch of n_subgroups on a device with `sms` SMs (no device needed)."""
    buf = np.zeros(64, dtype=np.int32)
    n = lib().rambl_solve_layout(n_subgroups, sms, buf.ctypes.data_as(_i32p), len(buf))
    if n < 0:
        _check(-n)
    return [int(x) for x in buf[:min(n, len(buf))]]


def cached_host_bytes() -> int:
    """Pageable host memory the library's block cache holds right now."""
    return int(lib().rambl_cached_host_bytes())


# ------------------------------------------------------------------------------------------------
class MultipleSequenceAlignmentSP:
    """MultipleSequenceAlignmentSP<Index2D,SimpleScoreModel,vector,string,char> with SimpleDnaScore defaults."""

    def align(self, data: Sequence[str]) -> List[str]:
        """Align ``data`` in order; row t of the result is MSA::get(t)."""
        return msa_align_batch([list(data)])[0][0]


def msa_align_batch(problems: Sequence[Sequence[str]]) -> Tuple[List[List[str]], Dict[str, float]]:
    """Many independent alignments in one device launch -> (rows per problem, {dp_cells, kernel_ms})."""
    L = lib()
    P = len(problems)
    pso = np.zeros(P + 1, dtype=np.int32)
    seqs: List[str] = []
    for p, pr in enumerate(problems):
        seqs.extend(pr)
        pso[p + 1] = len(seqs)
    so = np.zeros(len(seqs) + 1, dtype=np.int32)
    for i, s in enumerate(seqs):
        so[i + 1] = so[i] + len(s)
    letters = "".join(seqs).encode()
    cap = L.rambl_msa_rows_capacity(P, pso.ctypes.data_as(_i32p), so.ctypes.data_as(_i32p))
    rows = C.create_string_buffer(max(1, int(cap)))
    width = np.zeros(max(P, 1), dtype=np.int32)
    row_off = np.zeros(max(P, 1), dtype=np.int64)
    stride = np.zeros(max(P, 1), dtype=np.int32)
    cells = C.c_uint64(0)
    ms = C.c_float(0)
    _check(L.rambl_msa_sp_align_batch(P, pso.ctypes.data_as(_i32p), so.ctypes.data_as(_i32p), letters,
                                      width.ctypes.data_as(_i32p), row_off.ctypes.data_as(_i64p),
                                      stride.ctypes.data_as(_i32p), rows, C.byref(cells), C.byref(ms)))
    raw = rows.raw
    out: List[List[str]] = []
    for p, pr in enumerate(problems):
        base, st, w = int(row_off[p]), int(stride[p]), int(width[p])
        out.append([raw[base + t * st: base + t * st + w].decode() for t in range(len(pr))])
    return out, {"dp_cells": int(cells.value), "kernel_ms": float(ms.value)}


# ------------------------------------------------------------------------------------------------
class Strain:
    """What StrainCall reads from a reference ``Strain``: abundance, path, strain_seq(), plain_seq()."""

    def __init__(self, abundance_infer: float, abundance: float, path: List[int], seq: str, plain: str,
                 sub: np.ndarray, loglik: Optional[np.ndarray]):
        self.abundance_infer = abundance_infer
        self.abundance = abundance
        self.path = path
        self._seq = seq
        self._plain = plain
        self.sub_count = sub  # 6x6 over A,C,G,T,-,=
        self.read_loglik = loglik

    def strain_seq(self) -> str:
        return self._seq

    def plain_seq(self) -> str:
        return self._plain


class StrainCallBatch:
    """Any number of subgroups built and solved together on the device."""

    def __init__(self):
        self._h = lib().rambl_batch_create()
        if not self._h:
            raise RamblError(RAMBL_ERR_INVALID, "could not create a batch")
        self._nreads: List[int] = []
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            lib().rambl_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_subgroup(self, gene: str, pos, cigar: Sequence[str], seq: Sequence[str], copies,
                     pair_off=None, pair_val=None) -> int:
        n = len(pos)
        p, c = _i32(pos), _i32(copies)
        po = _i32(pair_off) if pair_off is not None else None
        pv = _i32(pair_val) if pair_val is not None else None
        rc = lib().rambl_batch_add_subgroup(
            self._h, gene.encode(), n, p.ctypes.data_as(_i32p), _strs(cigar), _strs(seq), c.ctypes.data_as(_i32p),
            po.ctypes.data_as(_i32p) if po is not None else None, pv.ctypes.data_as(_i32p) if pv is not None else None)
        if rc < 0:
            _check(-rc)
        self._nreads.append(n)
        return rc

    def add_graph(self, nodes: Sequence[dict], read_copies, pair_off=None, pair_val=None) -> int:
        """Add a subgroup whose graph was built elsewhere.  ``nodes`` is the `nodes` vector in order, each a dict
        with st (AlignState 0..3), label, out (ordered successor ids) and pool (ordered (rid, letters, copies))."""
        n = len(nodes)
        st = np.asarray([nd["st"] for nd in nodes], dtype=np.uint8)
        lab_off, out_off, pool_off, str_off = [0], [0], [0], [0]
        labs, outs, rid, cn, strs = [], [], [], [], []
        for nd in nodes:
            labs.append(nd["label"]); lab_off.append(lab_off[-1] + len(nd["label"]))
            outs.extend(nd["out"]); out_off.append(len(outs))
            for (r, s, c) in nd["pool"]:
                rid.append(r); cn.append(c); strs.append(s); str_off.append(str_off[-1] + len(s))
            pool_off.append(len(rid))
        rc_ = _i32(read_copies)
        arrs = [_i32(x) for x in (lab_off, out_off, outs or [0], pool_off, rid or [0], cn or [0], str_off)]
        po = _i32(pair_off) if pair_off is not None else None
        pv = _i32(pair_val) if pair_val is not None else None
        p = lambda a_: a_.ctypes.data_as(_i32p)
        rc = lib().rambl_batch_add_graph(
            self._h, n, len(rc_), st.ctypes.data_as(C.POINTER(C.c_uint8)), p(arrs[0]), "".join(labs).encode(), p(arrs[1]),
            p(arrs[2]), p(arrs[3]), p(arrs[4]), p(arrs[5]), p(arrs[6]), "".join(strs).encode(), p(rc_),
            p(po) if po is not None else None, p(pv) if pv is not None else None)
        if rc < 0:
            _check(-rc)
        self._nreads.append(len(rc_))
        return rc

    def add_subgroup_packed(self, gene: str, pos, cigar_off, cigar_chars: bytes, seq_off, seq_chars: bytes, copies,
                            pair_off=None, pair_val=None) -> int:
        """add_subgroup with the CIGAR and read strings as two byte arenas + offsets (no per-read objects)."""
        p, c = _i32(pos), _i32(copies)
        co = np.ascontiguousarray(cigar_off, dtype=np.int64)
        so = np.ascontiguousarray(seq_off, dtype=np.int64)
        po = _i32(pair_off) if pair_off is not None else None
        pv = _i32(pair_val) if pair_val is not None else None
        rc = lib().rambl_batch_add_subgroup_packed(
            self._h, gene.encode(), len(p), p.ctypes.data_as(_i32p), co.ctypes.data_as(_i64p), cigar_chars,
            so.ctypes.data_as(_i64p), seq_chars, c.ctypes.data_as(_i32p),
            po.ctypes.data_as(_i32p) if po is not None else None, pv.ctypes.data_as(_i32p) if pv is not None else None)
        if rc < 0:
            _check(-rc)
        self._nreads.append(len(p))
        return rc

    def add(self, sg) -> int:
        """Add a rambl_b200.synth.Subgroup (through its packed host buffers, built once per subgroup)."""
        pk = sg.packed()
        return self.add_subgroup_packed(sg.gene, pk["pos"], pk["cigar_off"], pk["cigar_chars"], pk["seq_off"], pk["seq_chars"],
                                        pk["cn"], sg.pair_off, sg.pair_val)

    def walk_plan(self, sg: int) -> dict:
        """What the device-resident walk will do with subgroup sg (no device needed once the graph is built)."""
        out = (C.c_int64 * 8)()
        _check(lib().rambl_batch_walk_plan(self._h, sg, out))
        keys = ("eligible", "handoff", "reason", "levels", "entries", "max_entries", "max_draws", "offtable_levels")
        d = {k: int(v) for k, v in zip(keys, out)}
        d["eligible"], d["handoff"] = bool(d["eligible"]), bool(d["handoff"])
        return d

    def build_graphs(self):
        _check(lib().rambl_batch_build_graphs(self._h))

    # the construction with the device step supplied by the caller (see include/rambl_b200.h)
    def thread_reads(self):
        _check(lib().rambl_batch_thread_reads(self._h))

    def msa_problems(self) -> List[List[str]]:
        txt = _take(lib().rambl_batch_msa_problems_text(self._h))
        out: List[List[str]] = []
        lines = txt.split("\n")
        i = 0
        while i < len(lines):
            if lines[i].startswith("P "):
                n = int(lines[i].split()[2])
                out.append(lines[i + 1:i + 1 + n])
                i += 1 + n
            else:
                i += 1
        return out

    def finish_graphs_with_rows(self, rows: Sequence[Sequence[str]]):
        txt = "".join("P %d %d\n%s\n" % (p, len(r), "\n".join(r)) for p, r in enumerate(rows))
        _check(lib().rambl_batch_finish_graphs_with_rows(self._h, txt.encode()))

    def infer(self, n: int = 5000, e: float = 0.01, tau: float = 0.02, diff: float = 0.01, assign: bool = True,
              keep_loglik: bool = False):
        _check(lib().rambl_batch_infer(self._h, n, e, tau, diff, int(assign), int(keep_loglik)))

    def solve(self, n: int = 5000, e: float = 0.01, tau: float = 0.02, diff: float = 0.01, assign: bool = True,
              keep_loglik: bool = False):
        """build_graphs() + infer() as one call that overlaps host graph construction with the device strain search."""
        _check(lib().rambl_batch_solve(self._h, n, e, tau, diff, int(assign), int(keep_loglik)))

    # results ------------------------------------------------------------------------------------
    def num_subgroups(self) -> int:
        return lib().rambl_batch_num_subgroups(self._h)

    def num_nodes(self, sg: int) -> int:
        return lib().rambl_batch_num_nodes(self._h, sg)

    def graph_dump(self, sg: int) -> str:
        return _take(lib().rambl_batch_graph_text(self._h, sg, 0))

    def output_edge(self, sg: int) -> str:
        return _take(lib().rambl_batch_graph_text(self._h, sg, 1))

    def status(self, sg: int) -> int:
        return lib().rambl_batch_status(self._h, sg)

    def strains_text(self, sg: int) -> str:
        return _take(lib().rambl_batch_strains_text(self._h, sg))

    def fasta(self, sg: int, gene_name: str, p0: int, p1: int, tau: float = 0.02) -> str:
        return _take(lib().rambl_batch_fasta(self._h, sg, gene_name.encode(), p0, p1, tau))

    def order(self, sg: int) -> List[int]:
        n = lib().rambl_batch_num_strains(self._h, sg)
        o = np.zeros(max(n, 1), dtype=np.int32)
        _check(lib().rambl_batch_order(self._h, sg, o.ctypes.data_as(_i32p)))
        return [int(x) for x in o[:n]]

    def strains(self, sg: int, with_loglik: bool = False) -> List[Strain]:
        L = lib()
        n = L.rambl_batch_num_strains(self._h, sg)
        if n < 0:
            _check(-n)
        out = []
        for k in range(n):
            a0, a1, pl = C.c_double(), C.c_double(), C.c_int32()
            _check(L.rambl_batch_strain(self._h, sg, k, C.byref(a0), C.byref(a1), C.byref(pl)))
            path = np.zeros(max(1, pl.value), dtype=np.int32)
            _check(L.rambl_batch_strain_path(self._h, sg, k, path.ctypes.data_as(_i32p)))
            sub = np.zeros(36, dtype=np.float64)
            _check(L.rambl_batch_strain_sub(self._h, sg, k, sub.ctypes.data_as(_f64p)))
            ll = None
            if with_loglik:
                ll = np.zeros(max(1, self._nreads[sg]), dtype=np.float64)
                _check(L.rambl_batch_strain_loglik(self._h, sg, k, ll.ctypes.data_as(_f64p), self._nreads[sg]))
            out.append(Strain(a0.value, a1.value, [int(x) for x in path[:pl.value]],
                              _take(L.rambl_batch_strain_sequence(self._h, sg, k, 0)),
                              _take(L.rambl_batch_strain_sequence(self._h, sg, k, 1)), sub.reshape(6, 6), ll))
        return out

    def stats(self) -> Dict[str, float]:
        st = Stats()
        _check(lib().rambl_batch_stats(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in Stats._fields_}


# ------------------------------------------------------------------------------------------------
class PartialOrderGraph:
    """PartialOrderGraph(G, R) for one subgroup (PartialOrderGraph.hpp:231-357).

    ``R`` holds AlignRead tuples (pos, cigar, seq, qual, copies).  ``read_pairs`` is the reference's
    ReadPairs (dict uid -> list with one mate uid or -1 per copy); it can be given here or to
    infer_strains / read_assign like in the reference -- the graph does not depend on it."""

    def __init__(self, G: str, R: Sequence[tuple], read_pairs=None):
        self._G = G
        self._pos = [r[0] for r in R]
        self._cigar = [r[1] for r in R]
        self._seq = [r[2] for r in R]
        self._cn = [r[4] if len(r) > 4 else 1 for r in R]
        self._pairs = None
        self._b = None
        self._make(read_pairs)

    def _csr(self, read_pairs):
        off, val = [0], []
        for u in range(len(self._cn)):
            mates = read_pairs.get(u) if isinstance(read_pairs, dict) else read_pairs[u]
            mates = list(mates) if mates is not None else [-1] * self._cn[u]
            if len(mates) < self._cn[u]:
                raise RamblError(RAMBL_ERR_INVALID, "ReadPairs needs one entry per read copy (uid %d)" % u)
            val.extend(int(m) for m in mates)
            off.append(len(val))
        return off, val

    def _make(self, read_pairs):
        if self._b is not None and (read_pairs is None or read_pairs == self._pairs):
            return
        off = val = None
        if read_pairs is not None:
            off, val = self._csr(read_pairs)
        b = StrainCallBatch()
        b.add_subgroup(self._G, self._pos, self._cigar, self._seq, self._cn, off, val)
        b.build_graphs()
        self._b, self._pairs = b, read_pairs

    @property
    def N(self) -> int:
        return self._b.num_nodes(0)

    def output_edge(self) -> str:
        return self._b.output_edge(0)

    def infer_strains(self, read_pairs=None, n: int = 5000, e: float = 0.01, tau: float = 0.02, diff: float = 0.01):
        """streaming_clustering: the strains in the order the reference leaves them in `strains`."""
        self._make(read_pairs)
        self._b.infer(n, e, tau, diff, assign=False)
        if self._b.status(0) != RAMBL_OK:
            raise RamblError(self._b.status(0), "no strain survived (the reference is undefined on this input)")
        return self._b.strains(0)

    def read_assign(self, strains=None, reads=None, read_pairs=None, n: int = 5000, e: float = 0.01,
                    tau: float = 0.02, diff: float = 0.01):
        """infer_strains + read_assign + the abundance sort of StrainCall's main(); returns the sorted strains
        (the device keeps the per-read likelihoods of the strains, so the two steps run as one call)."""
        self._make(read_pairs)
        self._b.infer(n, e, tau, diff, assign=True)
        if self._b.status(0) != RAMBL_OK:
            raise RamblError(self._b.status(0), "no strain survived (the reference is undefined on this input)")
        st = self._b.strains(0)
        return [st[k] for k in self._b.order(0)]
