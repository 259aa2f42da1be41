"""ctypes access to oracle/_ref (the UNMODIFIED reference compiled by oracle/Makefile).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under rambl_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: Dict[str, C.CDLL] = {}


def lib_path(variant: str = "") -> str:
    """variant '' | 'O0' | 'plain' -> oracle/_ref builds of the reference; 'oracle' -> our CPU restatement."""
    if variant == "oracle":
        return os.path.join(_HERE, "liboracle.so")
    name = "libstraincall_ref%s.so" % (("_" + variant) if variant else "")
    return os.path.join(_HERE, "_ref", name)


def available(variant: str = "") -> bool:
    return os.path.exists(lib_path(variant))


def _lib(variant: str = "") -> C.CDLL:
    if variant not in _LIBS:
        raw = C.CDLL(lib_path(variant))
        prefix = "orc_" if variant == "oracle" else "ref_"

        class _L:  # the two libraries export the same calls under different prefixes
            pass
        lib = _L()
        for fn in ("msa_align", "pog_build", "pog_free", "pog_num_nodes", "pog_dump", "pog_edges", "infer", "free"):
            setattr(lib, "ref_" + fn, getattr(raw, prefix + fn))
        lib.ref_msa_align.restype = C.c_void_p
        lib.ref_msa_align.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        lib.ref_pog_build.restype = C.c_void_p
        lib.ref_pog_build.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_char_p),
                                      C.POINTER(C.c_char_p), C.POINTER(C.c_int)]
        lib.ref_pog_free.argtypes = [C.c_void_p]
        lib.ref_pog_num_nodes.argtypes = [C.c_void_p]
        lib.ref_pog_dump.restype = C.c_void_p
        lib.ref_pog_dump.argtypes = [C.c_void_p]
        lib.ref_pog_edges.restype = C.c_void_p
        lib.ref_pog_edges.argtypes = [C.c_void_p]
        lib.ref_infer.restype = C.c_void_p
        lib.ref_infer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_double,
                                  C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double)]
        lib.ref_free.argtypes = [C.c_void_p]
        _LIBS[variant] = lib
    return _LIBS[variant]


def _take(lib, p) -> str:
    s = C.string_at(p).decode()
    lib.ref_free(p)
    return s


def _strs(xs: Sequence[str]):
    arr = (C.c_char_p * len(xs))()
    arr[:] = [x.encode() for x in xs]
    return arr


def msa_align(seqs: Sequence[str], variant: str = "") -> List[str]:
    """MultipleSequenceAlignmentSP::align + MSA::get(t) for every t."""
    lib = _lib(variant)
    txt = _take(lib, lib.ref_msa_align(len(seqs), _strs(seqs)))
    lines = txt.split("\n")
    ncol = int(lines[0])
    rows = lines[1:1 + len(seqs)]
    assert all(len(r) == ncol for r in rows), (ncol, rows)
    return rows


class RefPog:
    """A reference PartialOrderGraph built from in-memory reads."""

    def __init__(self, gene: str, pos, cigar, seq, cn, variant: str = ""):
        self.lib = _lib(variant)
        n = len(pos)
        self._pos = (C.c_int * n)(*[int(x) for x in pos])
        self._cn = (C.c_int * n)(*[int(x) for x in cn])
        self.h = self.lib.ref_pog_build(gene.encode(), n, self._pos, _strs(cigar), _strs(seq), self._cn)

    def close(self):
        if self.h:
            self.lib.ref_pog_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def num_nodes(self) -> int:
        return self.lib.ref_pog_num_nodes(self.h)

    def dump(self) -> str:
        return _take(self.lib, self.lib.ref_pog_dump(self.h))

    def edges(self) -> str:
        return _take(self.lib, self.lib.ref_pog_edges(self.h))

    def infer(self, pair_off, pair_val, n: int = 5000, e: float = 0.01, tau: float = 0.02, diff: float = 0.01,
              do_assign: bool = True, with_loglik: bool = True):
        po = np.ascontiguousarray(pair_off, dtype=np.int32)
        pv = np.ascontiguousarray(pair_val, dtype=np.int32)
        ms = (C.c_double * 2)()
        txt = _take(self.lib, self.lib.ref_infer(
            self.h, len(po) - 1, po.ctypes.data_as(C.POINTER(C.c_int)), pv.ctypes.data_as(C.POINTER(C.c_int)),
            n, e, tau, diff, int(do_assign), int(with_loglik), ms))
        return parse_strain_dump(txt), (ms[0], ms[1])


def parse_graph_dump(txt: str) -> List[dict]:
    nodes = []
    for line in txt.split("\n"):
        if not line.startswith("NODE "):
            continue
        head, out, inn, sib, pool = [x.strip() for x in line.split("|")]
        h = head.split()
        nodes.append(dict(
            id=int(h[1]), st=int(h[2]), label=h[3], level=int(h[4]),
            out=[int(x) for x in out.split()[1:]],
            inn=[int(x) for x in inn.split()[1:]],
            sib=[int(x) for x in sib.split()[1:]],
            pool=[(int(a), b, int(c)) for a, b, c in (x.split(":") for x in pool.split()[1:])],
        ))
    return nodes


def parse_strain_dump(txt: str) -> Dict[str, List[dict]]:
    """Stages ('infer','assign','final') -> list of strain dicts (abundance as float + exact string)."""
    stages: Dict[str, List[dict]] = {}
    cur: Optional[List[dict]] = None
    for line in txt.split("\n"):
        if line.startswith("STAGE "):
            cur = []
            stages[line.split()[1]] = cur
        elif line.startswith("STRAIN "):
            f = line.split()
            cur.append(dict(abundance_ld=f[2], abundance=float(f[3]), Z=float(f[4])))
        elif line.startswith("PATH"):
            cur[-1]["path"] = [int(x) for x in line.split()[1:]]
        elif line.startswith("SEQ "):
            cur[-1]["seq"] = line[4:]
        elif line.startswith("PLAIN"):
            cur[-1]["plain"] = line[6:]
        elif line.startswith("SUB "):
            cur[-1]["sub"] = [float(x) for x in line.split()[1:]]
        elif line.startswith("LOGLIK "):
            f = line.split()
            cur[-1]["loglik"] = {int(a): float(b) for a, b in (x.split(":") for x in f[2:])}
    return stages
