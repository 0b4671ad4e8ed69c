"""ctypes binding of the CPU ORACLE (oracle/cpo.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package.  It exposes the reference's
function names (partition_stripe, pack_stripe, partition_plaid, oracle_stripe, bound_stripe,
netcount, ...) over the same host-side types as the product so parity tests read
``gpu.partition_stripe(A, K, mtd).spl == ref.partition_stripe(A, K, mtd).spl``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
import time
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from chainb200 import types as T  # noqa: E402  (plain data types only; no CUDA involved)

I64 = np.int64
_LIB_PATH = os.path.join(_HERE, "_build", "libcpo.so")


def build(force: bool = False) -> str:
    """Compile oracle/_build/libcpo.so with g++ (Makefile next to this file)."""
    srcs = [os.path.join(_HERE, f) for f in ("cpo.cpp", "cpo_core.hpp", "cpo_solvers.hpp", "cpo.h")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _CSC(ctypes.Structure):
    _fields_ = [
        ("m", ctypes.c_longlong),
        ("n", ctypes.c_longlong),
        ("nnz", ctypes.c_longlong),
        ("colptr", ctypes.c_void_p),
        ("rowval", ctypes.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.cpo_last_error.restype = ctypes.c_char_p
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError("oracle: " + lib().cpo_last_error().decode())


def _csc(A: T.SparseMatrixCSC) -> _CSC:
    return _CSC(A.m, A.n, A.nnz, A.colptr.ctypes.data, A.rowval.ctypes.data)


def _arr(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=I64)


def _ptr(a: Optional[np.ndarray]):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def _pi(Pi):
    if Pi is None:
        return None, 0
    if not isinstance(Pi, T.SplitPartition):
        raise TypeError("row partition must be a SplitPartition")
    return _arr(Pi.spl), Pi.K


# ----------------------------------------------------------------------------- structures


def adjointpattern(A: T.SparseMatrixCSC) -> T.SparseMatrixCSC:
    pos = np.empty(A.m + 1, dtype=I64)
    idx = np.empty(A.nnz, dtype=I64)
    _check(lib().cpo_adjointpattern(ctypes.byref(_csc(A)), _ptr(pos), _ptr(idx)))
    return T.SparseMatrixCSC(A.n, A.m, pos, idx)


def dominancecount(A, i, j, hint=T.NoHint(), b=0, H=0, bp=0) -> np.ndarray:
    i, j = _arr(i), _arr(j)
    out = np.empty(len(i), dtype=I64)
    _check(lib().cpo_dominancecount(hint.code, ctypes.byref(_csc(A)), int(b), int(H), int(bp), ctypes.c_longlong(len(i)), _ptr(i), _ptr(j), _ptr(out)))
    return out


def dominancesum(A, val, i, j, b=0, H=0, bp=0) -> np.ndarray:
    """``dominancesum(hint, A; b, H, b')[i, j]`` through the restated DominanceSum structure (SparsePrefixMatrices.jl:1-254)."""
    i, j = _arr(i), _arr(j)
    val = np.asarray(val)
    dtype = np.uint64 if val.dtype == np.uint64 else I64
    v = np.ascontiguousarray(val.astype(dtype, copy=False)).view(I64)
    out = np.empty(len(i), dtype=I64)
    _check(lib().cpo_dominancesum(ctypes.byref(_csc(A)), _ptr(v), int(b), int(H), int(bp), ctypes.c_longlong(len(i)), _ptr(i), _ptr(j), _ptr(out)))
    return out.view(dtype)


def prefix_query(m, n, N, pos, idx, val, i, j) -> np.ndarray:
    """dominancesum / rookcount / rooksum entries by the offline sweep of cpo_prefix_query (pos None: rook form; val None: counts)."""
    i, j, idx = _arr(i), _arr(j), _arr(idx)
    pos = None if pos is None else _arr(pos)
    dtype = I64
    if val is not None:
        val = np.asarray(val)
        dtype = np.uint64 if val.dtype == np.uint64 else I64
        val = np.ascontiguousarray(val.astype(dtype, copy=False)).view(I64)
    out = np.empty(len(i), dtype=I64)
    L = ctypes.c_longlong
    _check(lib().cpo_prefix_query(L(m), L(n), L(N), _ptr(pos), _ptr(idx), _ptr(val), L(len(i)), _ptr(i), _ptr(j), _ptr(out)))
    return out.view(dtype)


def dominancecount_walk(A, i, j) -> np.ndarray:
    i, j = _arr(i), _arr(j)
    out = np.empty(len(i), dtype=I64)
    _check(lib().cpo_dominancecount_walk(ctypes.byref(_csc(A)), ctypes.c_longlong(len(i)), _ptr(i), _ptr(j), _ptr(out)))
    return out


def _colorcount(which, A, j, jp, hint):
    j, jp = _arr(j), _arr(jp)
    out = np.empty(len(j), dtype=I64)
    _check(lib().cpo_colorcount(which, hint.code, ctypes.byref(_csc(A)), ctypes.c_longlong(len(j)), _ptr(j), _ptr(jp), _ptr(out)))
    return out


def pincount(A, j, jp, hint=T.NoHint()):
    return _colorcount(0, A, j, jp, hint)


def netcount(A, j, jp, hint=T.NoHint()):
    return _colorcount(1, A, j, jp, hint)


def dianetcount(A, j, jp, hint=T.NoHint()):
    return _colorcount(2, A, j, jp, hint)


def selfnetcount(A, j, jp, hint=T.NoHint()):
    return _colorcount(3, A, j, jp, hint)


def selfpincount(A, j, jp, hint=T.NoHint()):
    return _colorcount(4, A, j, jp, hint)


def rowenvelope(A, j, jp):
    j, jp = _arr(j), _arr(jp)
    lo = np.empty(len(j), dtype=I64)
    hi = np.empty(len(j), dtype=I64)
    _check(lib().cpo_rowenvelope(ctypes.byref(_csc(A)), ctypes.c_longlong(len(j)), _ptr(j), _ptr(jp), _ptr(lo), _ptr(hi)))
    return lo, hi


# ----------------------------------------------------------------------------- oracles


def _model(mdl, A, con, Pi):
    cm, keep = mdl.to_c(**T.model_tables(mdl, A, con, Pi))
    return cm, keep


def oracle_query(mdl, A, j, jp, k=None, hint=T.NoHint(), Pi=None) -> np.ndarray:
    """``oracle_stripe(hint, mdl, A[, Pi])(j, j', k)`` for arrays of queries (cost as float64)."""
    f, con = T.split_constrained(mdl)
    j, jp = _arr(j), _arr(jp)
    kk = _arr(k) if k is not None else None
    spl, pK = _pi(Pi)
    cm, keep = _model(f, A, T.CConstraint(), Pi)
    out = np.empty(len(j), dtype=np.float64)
    _check(lib().cpo_oracle_query(ctypes.byref(cm), hint.code, ctypes.byref(_csc(A)), _ptr(spl), ctypes.c_longlong(pK),
                                  ctypes.c_longlong(len(j)), _ptr(j), _ptr(jp), _ptr(kk), ctypes.c_void_p(out.ctypes.data)))
    return out


def bound_stripe(A, K, mdl, via_oracle=True):
    """``bound_stripe(A, K, ocl)`` (the oracle form every Bisect splitter calls; default) or, with ``via_oracle=False``,
    ``bound_stripe(A, K, mdl)`` -- the two differ only for the envelope model (EnvelopeCosts.jl:30-42 vs :44-54)."""
    cm, keep = _model(mdl, A, T.CConstraint(), None)
    out = (ctypes.c_double * 2)()
    _check(lib().cpo_bound_stripe(ctypes.byref(cm), ctypes.byref(_csc(A)), ctypes.c_longlong(K), int(bool(via_oracle)), out))
    return (out[0], out[1])


def permute(A, col_prm=None, row_new=None) -> T.SparseMatrixCSC:
    """A[:, col_prm] with row r renamed row_new[r] (numpy; Costs.jl:34-39, 52-57 "A_prm = A[:, Phi_dom.prm]")."""
    deg = np.diff(A.colptr)
    prm = np.arange(A.n) if col_prm is None else _arr(col_prm) - 1
    pos = np.concatenate(([1], 1 + np.cumsum(deg[prm]))).astype(I64)
    src = np.concatenate([np.arange(A.colptr[c] - 1, A.colptr[c + 1] - 1) for c in prm]).astype(I64) if A.nnz else np.zeros(0, dtype=I64)
    rows = A.rowval[src]
    if row_new is not None:
        rows = _arr(row_new)[rows - 1]
    return T.SparseMatrixCSC(A.m, A.n, pos, rows)


def _contiguous(A, Phi, Pi):
    """compute_objective's reduction of Map / DomainPartitions to SplitPartitions of a permuted matrix."""
    if Pi is not None and not isinstance(Pi, T.SplitPartition):
        dom = T.convert(T.DomainPartition, Pi)
        row_new = np.empty(A.m, dtype=I64)
        row_new[dom.prm - 1] = np.arange(1, A.m + 1, dtype=I64)
        A, Pi = permute(A, None, row_new), T.SplitPartition(dom.K, dom.spl)
    if not isinstance(Phi, T.SplitPartition):
        dom = T.convert(T.DomainPartition, Phi)
        A, Phi = permute(A, dom.prm, None), T.SplitPartition(dom.K, dom.spl)
    return A, Phi, Pi


def _objective(total, A, Phi, mdl, Pi=None, hint=T.StepHint()):
    A, Phi, Pi = _contiguous(A, Phi, Pi)
    f, _ = T.split_constrained(mdl)
    spl, pK = _pi(Pi)
    cm, keep = _model(f, A, T.CConstraint(), Pi)
    out = ctypes.c_double()
    s = _arr(Phi.spl)
    _check(lib().cpo_objective(int(total), ctypes.byref(cm), hint.code, ctypes.byref(_csc(A)), _ptr(spl), ctypes.c_longlong(pK),
                               ctypes.c_longlong(Phi.K), _ptr(s), ctypes.byref(out)))
    return out.value


def bottleneck_value(A, Phi, mdl, Pi=None):
    return _objective(0, A, Phi, mdl, Pi)


def total_value(A, Phi, mdl, Pi=None):
    return _objective(1, A, Phi, mdl, Pi)


# ----------------------------------------------------------------------------- solvers

last_seconds = [0.0, 0.0]  # (oracle build, solve) of the most recent solver call


def partition_stripe(A, K, method, Pi=None, **kwargs) -> T.SplitPartition:
    code, spec, eps = T.split_method_code(method)
    spl = np.empty(K + 1, dtype=I64)
    secs = (ctypes.c_double * 2)()
    pspl, pK = _pi(Pi)
    if spec is None:
        cm, con, keep = T.CModel(), T.CConstraint(), []
    else:
        f, con = T.split_constrained(spec)
        cm, keep = _model(f, A, con, Pi)
    _check(lib().cpo_partition_stripe(code, ctypes.byref(cm), ctypes.byref(con), ctypes.c_double(eps), ctypes.byref(_csc(A)),
                                      _ptr(pspl), ctypes.c_longlong(pK), ctypes.c_longlong(K), _ptr(spl), secs))
    last_seconds[:] = [secs[0], secs[1]]
    return T.SplitPartition(K, spl)


def pack_stripe(A, method, Pi=None, n_nets=None, **kwargs) -> T.SplitPartition:
    code, spec, rho, w_max = T.pack_method_code(method)
    spl = np.empty(A.n + 1, dtype=I64)
    nn = np.zeros(max(A.n, 1), dtype=I64)
    Kout = ctypes.c_longlong()
    secs = (ctypes.c_double * 2)()
    pspl, pK = _pi(Pi)
    if spec is None:
        cm, con, keep = T.CModel(), T.CConstraint(), []
    else:
        f, con = T.split_constrained(spec)
        cm, keep = _model(f, A, con, Pi)
    _check(lib().cpo_pack_stripe(code, ctypes.byref(cm), ctypes.byref(con), ctypes.c_double(rho), ctypes.c_longlong(w_max),
                                 ctypes.byref(_csc(A)), _ptr(pspl), ctypes.c_longlong(pK), _ptr(spl), ctypes.byref(Kout), _ptr(nn), secs))
    last_seconds[:] = [secs[0], secs[1]]
    K = Kout.value
    if n_nets is not None:
        n_nets[:] = [nn[:K].copy()]
    return T.SplitPartition(K, spl[: K + 1].copy())


def partition_plaid(A, K, method, adj_A=None, **kwargs):
    """AlternatingPartitioner.jl:6-88 (host-side orchestration of stripe solves)."""
    if isinstance(method, T.DisjointPartitioner):
        Phi = partition_stripe(A, K, method.mtd)
        Pi = partition_stripe(adjointpattern(A), K, method.mtd2, Phi)
        return Pi, Phi
    if isinstance(method, T.AlternatingPartitioner):
        if adj_A is None:
            adj_A = adjointpattern(A)
        Phi = partition_stripe(A, K, method.mtds[0])
        Pi = partition_stripe(adj_A, K, method.mtds[1], Phi)
        for i, mtd in enumerate(method.mtds[2:], start=1):
            if i % 2 == 1:
                Phi = partition_stripe(A, K, mtd, Pi)
            else:
                Pi = partition_stripe(adj_A, K, mtd, Phi)
        return Pi, Phi
    if isinstance(method, T.SymmetricPartitioner):
        if len(method.mtds) > 1:
            if adj_A is None:
                adj_A = adjointpattern(A)
            Pi = partition_stripe(A, K, method.mtds[0])
            for i, mtd in enumerate(method.mtds[1:], start=1):
                Pi = partition_stripe(A if i % 2 == 1 else adj_A, K, mtd, Pi)
        else:
            Pi = partition_stripe(A, K, method.mtds[0])
        return Pi, Pi
    raise TypeError(f"partition_plaid: unsupported method {type(method).__name__}")


def pack_plaid(A, method, adj_A=None, **kwargs):
    """AlternatingPacker.jl:6-53."""
    if isinstance(method, T.DisjointPacker):
        Phi = pack_stripe(A, method.mtd)
        Pi = pack_stripe(adjointpattern(A), method.mtd2, Phi)
        return Pi, Phi
    if adj_A is None:
        adj_A = adjointpattern(A)
    if isinstance(method, T.AlternatingPacker):
        Phi = pack_stripe(A, method.mtds[0])
        Pi = pack_stripe(adj_A, method.mtds[1], Phi)
        for i, mtd in enumerate(method.mtds[2:], start=1):
            if i % 2 == 1:
                Phi = pack_stripe(A, mtd, Pi)
            else:
                Pi = pack_stripe(adj_A, mtd, Phi)
        return Pi, Phi
    if isinstance(method, T.SymmetricPacker):
        Pi = pack_stripe(A, method.mtds[0])
        for i, mtd in enumerate(method.mtds[1:], start=1):
            Pi = pack_stripe(A if i % 2 == 1 else adj_A, mtd, Pi)
        return Pi, Pi
    raise TypeError(f"pack_plaid: unsupported method {type(method).__name__}")
