"""ctypes binding of oracle/liboracle.so (the plain-C restatement of the reference; TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "relem_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
        L = C.CDLL(LIB)
        vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_model_new.restype = vp
        L.orc_model_new.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
        L.orc_model_set_debug.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
        L.orc_model_set_params.argtypes = [vp, dp, dp, C.c_double]
        for f in ("orc_hmm_M", "orc_hmm_S", "orc_hmm_nparam"):
            getattr(L, f).argtypes = [vp]
        L.orc_hmm_get.argtypes = [vp, C.c_int, ip]
        L.orc_energy_get.argtypes = [vp, C.c_char_p, dp, C.c_int]
        L.orc_bpp.restype = C.c_double
        L.orc_bpp.argtypes = [vp, ip, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, dp]
        L.orc_estep_seq.argtypes = [vp, ip, C.c_int, dp, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp]
        L.orc_scan_seq.argtypes = [vp, ip, C.c_int, dp, dp, dp, dp, ip, C.c_char_p, ip, ip, dp, dp]
        L.orc_debug_eval.argtypes = [vp, ip, C.c_int, dp, C.c_double, dp, dp, dp, dp]
        L.orc_last_counts.argtypes = [vp, dp, dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Oracle(object):
    def __init__(self, pattern, energy="~T2004~", max_span=50, max_iloop=30, min_bpp=1e-4, no_rss=0, no_prf=0, no_ene=0):
        self.L = lib()
        es = {"~T2004~": 0, "~A2007~": 1}[energy]
        self.h = self.L.orc_model_new(pattern.encode(), es, int(min(max_span, 2 ** 31 - 1)), int(min(max_iloop, 2 ** 31 - 1)),
                                      float(min_bpp), int(no_rss), int(no_prf), int(no_ene))
        assert self.h, "bad pattern"
        self.M, self.S = self.L.orc_hmm_M(self.h), self.L.orc_hmm_S(self.h)
        self.n_theta = self.L.orc_hmm_nparam(self.h)
        self.max_span = max_span

    @classmethod
    def from_model(cls, model):
        import rnaelem_b200 as rb
        o = cls(model["pattern"], model["ene-param"], model["max-span"], model["max-internal-loop"], model["min-bpp"],
                model.get("no-rss", 0), model.get("no-profile", 0), model.get("no-energy", 0))
        o.set_params(rb.model_theta_flat(model), model["lambda"], model["tau"])
        return o

    def set_params(self, theta, lam, tau):
        th = np.ascontiguousarray(theta, dtype=np.float64); la = np.ascontiguousarray(lam, dtype=np.float64)
        self.L.orc_model_set_params(self.h, _d(th), _d(la), float(tau))

    def set_debug(self, no_theta, no_turn, fix_rss):
        self.L.orc_model_set_debug(self.h, int(no_theta), int(no_turn), fix_rss.encode() if fix_rss is not None else None)

    def hmm_get(self, kind):
        n = self.L.orc_hmm_get(self.h, kind, None)
        buf = (C.c_int * max(1, n))()
        self.L.orc_hmm_get(self.h, kind, buf)
        return list(buf)[:n]

    def energy_get(self, name):
        buf = np.zeros(40000)
        n = self.L.orc_energy_get(self.h, name.encode(), _d(buf), 40000)
        assert n >= 0, name
        return buf[:n].copy()

    def bpp(self, seq):
        seq = np.ascontiguousarray(seq, dtype=np.int32)
        L = len(seq); W = min(L, self.max_span)
        n = (L + 1) * (W + 1)
        bp = np.zeros(n, np.uint8); lf = np.zeros(n, np.uint8); ln = np.zeros(n); lnz = np.zeros(1)
        eff = self.L.orc_bpp(self.h, _i(seq), L, bp.ctypes.data, lf.ctypes.data, ln.ctypes.data, _d(lnz))
        return bp, lf, ln, eff, lnz[0]

    def estep_seq(self, seq, ws, restricted, is_negative):
        seq = np.ascontiguousarray(seq, dtype=np.int32); ws = np.ascontiguousarray(ws, dtype=np.float64)
        NT = self.n_theta
        Z = np.zeros(3); Zx = np.zeros(1); ENo = np.zeros(NT); ENx = np.zeros(NT); EHo = np.zeros(2); EHx = np.zeros(2)
        eff = np.zeros(1)
        sk = self.L.orc_estep_seq(self.h, _i(seq), len(seq), _d(ws), int(restricted), int(is_negative), _d(Z), _d(Zx),
                                  _d(ENo), _d(ENx), _d(EHo), _d(EHx), _d(eff))
        return dict(skipped=sk, Z=Z, Zx=Zx[0], ENo=ENo, ENx=ENx, EHo=EHo, EHx=EHx, bpp_eff=eff[0])

    def scan_seq(self, seq, ws):
        seq = np.ascontiguousarray(seq, dtype=np.int32); ws = np.ascontiguousarray(ws, dtype=np.float64)
        L = len(seq)
        Pys = np.zeros(L); Pye = np.zeros(L + 1); Pyi = np.zeros(L); psi = np.zeros(L, np.int32)
        rss = C.create_string_buffer(L + 1); Ys = C.c_int(); Ye = C.c_int(); ex = np.zeros(1); EN = np.zeros(self.n_theta)
        self.L.orc_scan_seq(self.h, _i(seq), L, _d(ws), _d(Pys), _d(Pye), _d(Pyi), _i(psi), rss, C.byref(Ys), C.byref(Ye),
                            _d(ex), _d(EN))
        return dict(PysL=Pys, PyeL=Pye, PyiL=Pyi, psihat=psi, rss=rss.raw[:L].decode(), Ys=Ys.value, Ye=Ye.value,
                    exist=ex[0], EN=EN)

    def debug_eval(self, seq, ws, ZL=0.0):
        seq = np.ascontiguousarray(seq, dtype=np.int32); ws = np.ascontiguousarray(ws, dtype=np.float64)
        pf = np.zeros(1); pfo = np.zeros(1); EN = np.zeros(self.n_theta); EH = np.zeros(2)
        self.L.orc_debug_eval(self.h, _i(seq), len(seq), _d(ws), float(ZL), _d(pf), _d(pfo), _d(EN), _d(EH))
        return pf[0], pfo[0], EN, EH

    def counts(self):
        a = np.zeros(1); b = np.zeros(1)
        self.L.orc_last_counts(self.h, _d(a), _d(b))
        return a[0], b[0]
