"""Shared test helpers (CPU): loads the C restatement of the evaluator oracle."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_c_oracle():
    so = os.path.join(ROOT, "oracle", "_build", "liboracle_eval.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so)
    lib.grl_oracle_evaluate.restype = ctypes.c_int
    return lib


def c_evaluate(lib, d, qp, gp, qc, gc, max_rank=100):
    d = np.ascontiguousarray(d, np.float32)
    nq, ng = d.shape
    mr = min(max_rank, ng)
    cmc = np.zeros(mr, np.float32)
    m = ctypes.c_double()
    ap = np.zeros(nq, np.float64)
    arr = lambda a: np.ascontiguousarray(a, np.int64).ctypes.data_as(ctypes.c_void_p)
    nv = lib.grl_oracle_evaluate(d.ctypes.data_as(ctypes.c_void_p), nq, ng, arr(qp), arr(gp), arr(qc), arr(gc),
                                 max_rank, cmc.ctypes.data_as(ctypes.c_void_p), ctypes.byref(m),
                                 ap.ctypes.data_as(ctypes.c_void_p))
    return cmc, m.value, nv, ap
