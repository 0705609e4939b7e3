"""ctypes binding of include/relem.h.  This is the stub a maintainer of a Python host would write; the C++ host
(the reference's own language) binds the same symbols directly."""
import ctypes as C
import os
import numpy as np

POS_WITHOUT, POS_WITH, NEG = 0, 1, 2
LR_WITHOUT, LR_NEG = 3, 4   # --lik-ratio kinds (include/relem.h)

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    # RELEM_LIBRARY: an A/B build of the same CUDA sources with other compile-time switches (tools/build_variant.sh);
    # always a CUDA build of this repository's kernels, never a CPU path
    return os.environ.get("RELEM_LIBRARY") or os.path.join(_HERE, "librelem.so")


class RelemError(RuntimeError):
    pass


class _EstepOut(C.Structure):
    _fields_ = [("fn", C.c_double), ("EN_diff", C.POINTER(C.c_double)), ("EH_diff", C.c_double * 2),
                ("sum_eff", C.c_double), ("n_skipped", C.c_int64), ("Z", C.POINTER(C.c_double)),
                ("ENo", C.POINTER(C.c_double)), ("ENx", C.POINTER(C.c_double)), ("EH", C.POINTER(C.c_double)),
                ("bpp_eff", C.POINTER(C.c_double)), ("skipped", C.POINTER(C.c_uint8))]


class _ScanOut(C.Structure):
    _fields_ = [("PysL", C.POINTER(C.c_double)), ("PyeL", C.POINTER(C.c_double)), ("PyiL", C.POINTER(C.c_double)),
                ("psihat", C.POINTER(C.c_int32)), ("rss", C.c_char_p), ("Ys", C.POINTER(C.c_int32)),
                ("Ye", C.POINTER(C.c_int32)), ("exist_prob", C.POINTER(C.c_double)), ("EN", C.POINTER(C.c_double)),
                ("ZL", C.POINTER(C.c_double))]


_LIBS = {}

SYMBOLS = ["relem_version", "relem_create", "relem_destroy", "relem_last_error", "relem_set_energy",
           "relem_set_pattern", "relem_model_dims", "relem_theta_rows", "relem_hmm_get", "relem_energy_get",
           "relem_set_params", "relem_batch_create", "relem_batch_destroy", "relem_batch_cells", "relem_estep_run",
           "relem_estep", "relem_bpp", "relem_scan_run", "relem_scan", "relem_comm_unique_id", "relem_comm_init",
           "relem_allreduce_sum", "relem_assigned_range", "relem_last_timing", "relem_fp64_peak"]


def load_library(path=None):
    path = path or lib_path()
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        raise RelemError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.relem_version.restype = C.c_char_p
    lib.relem_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.relem_destroy.argtypes = [vp]
    lib.relem_destroy.restype = None
    lib.relem_last_error.argtypes = [vp]
    lib.relem_last_error.restype = C.c_char_p
    lib.relem_set_energy.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_int]
    lib.relem_set_pattern.argtypes = [vp, C.c_char_p, C.c_int, C.c_int]
    lib.relem_model_dims.argtypes = [vp, ip, ip, ip, ip]
    lib.relem_theta_rows.argtypes = [vp, ip]
    lib.relem_hmm_get.argtypes = [vp, C.c_int, ip]
    lib.relem_energy_get.argtypes = [vp, C.c_char_p, dp, C.c_int]
    lib.relem_set_params.argtypes = [vp, dp, C.c_int, dp, C.c_double]
    lib.relem_batch_create.argtypes = [vp, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.POINTER(vp)]
    lib.relem_batch_destroy.argtypes = [vp, vp]
    lib.relem_batch_destroy.restype = None
    lib.relem_batch_cells.argtypes = [vp]
    lib.relem_batch_cells.restype = C.c_int64
    lib.relem_estep_run.argtypes = [vp, vp, C.POINTER(_EstepOut)]
    lib.relem_estep.argtypes = [vp, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.POINTER(_EstepOut)]
    lib.relem_bpp.argtypes = [vp, vp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.relem_scan_run.argtypes = [vp, vp, C.POINTER(_ScanOut)]
    lib.relem_scan.argtypes = [vp, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_ScanOut)]
    lib.relem_comm_unique_id.argtypes = [C.c_void_p]
    lib.relem_comm_init.argtypes = [vp, C.c_void_p, C.c_int, C.c_int]
    lib.relem_allreduce_sum.argtypes = [vp, C.c_void_p, C.c_int]
    lib.relem_assigned_range.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.relem_assigned_range.restype = None
    lib.relem_last_timing.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), ip, C.c_int]
    lib.relem_fp64_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    _LIBS[path] = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class EstepResult(object):
    """fn, EN_diff (theta shaped flat), EH_diff[2], sum_eff, n_skipped (+ per-sequence detail when asked)."""


class ScanResult(object):
    pass


class Batch(object):
    def __init__(self, ctx, handle, off, nseq):
        self.ctx, self.handle, self.off, self.nseq = ctx, handle, off, nseq

    @property
    def cells(self):
        return int(self.ctx.lib.relem_batch_cells(self.handle))

    def close(self):
        if self.handle is not None:
            self.ctx.lib.relem_batch_destroy(self.ctx.h, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context(object):
    """One GPU context (one per rank / host thread)."""

    def __init__(self, device=0, lib=None):
        self.lib = load_library(lib)
        h = C.c_void_p()
        rc = self.lib.relem_create(C.byref(h), int(device))
        if rc != 0:
            raise RelemError("relem_create failed (%d): %s" % (rc, self.lib.relem_last_error(None).decode()))
        self.h = h
        self.n_theta = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.relem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RelemError("%s failed (%d): %s" % (what, rc, self.lib.relem_last_error(self.h).decode()))

    # ---- model
    def set_energy(self, param="~T2004~", max_span=50, max_iloop=30, min_bpp=1e-4, no_ene=False):
        self._check(self.lib.relem_set_energy(self.h, param.encode(), int(max_span), int(max_iloop), float(min_bpp),
                                              int(bool(no_ene))), "relem_set_energy")
        self._ms = int(max_span)

    def set_pattern(self, pattern, no_rss=False, no_prf=False):
        self._check(self.lib.relem_set_pattern(self.h, pattern.encode(), int(bool(no_rss)), int(bool(no_prf))),
                    "relem_set_pattern")
        M, S, R, T = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.relem_model_dims(self.h, C.byref(M), C.byref(S), C.byref(R), C.byref(T)), "relem_model_dims")
        self.M, self.S, self.n_rows, self.n_theta = M.value, S.value, R.value, T.value
        rows = (C.c_int * self.n_rows)()
        self._check(self.lib.relem_theta_rows(self.h, rows), "relem_theta_rows")
        self.row_sizes = list(rows)

    def set_params(self, theta_flat, lam, tau):
        th = np.ascontiguousarray(theta_flat, dtype=np.float64)
        la = np.ascontiguousarray(lam, dtype=np.float64)
        self._check(self.lib.relem_set_params(self.h, _dptr(th), len(th), _dptr(la), float(tau)), "relem_set_params")

    def set_model(self, model):
        """model: dict from hostio.read_model."""
        from .hostio import model_theta_flat
        self.set_energy(model["ene-param"], model["max-span"], model["max-internal-loop"], model["min-bpp"],
                        model.get("no-energy", 0))
        self.set_pattern(model["pattern"], model.get("no-rss", 0), model.get("no-profile", 0))
        self.set_params(model_theta_flat(model), model["lambda"], model["tau"])

    def hmm_get(self, kind):
        n = self.lib.relem_hmm_get(self.h, kind, None)
        if n < 0:
            raise RelemError("relem_hmm_get(%d)" % kind)
        buf = (C.c_int * max(n, 1))()
        self.lib.relem_hmm_get(self.h, kind, buf)
        return list(buf)[:n]

    def energy_get(self, name):
        n = self.lib.relem_energy_get(self.h, name.encode(), None, 0)
        if n < 0:
            raise RelemError("relem_energy_get(%s)" % name)
        buf = np.zeros(max(n, 1))
        self.lib.relem_energy_get(self.h, name.encode(), _dptr(buf), n)
        return buf[:n]

    # ---- batches
    def batch(self, seq_cat, off, ws_cat, kind=None, gate=None):
        seq_cat = np.ascontiguousarray(seq_cat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        ws_cat = np.ascontiguousarray(ws_cat, dtype=np.float64)
        kind = None if kind is None else np.ascontiguousarray(kind, dtype=np.uint8)
        gate = None if gate is None else np.ascontiguousarray(gate, dtype=np.int32)
        h = C.c_void_p()
        self._check(self.lib.relem_batch_create(self.h, len(off) - 1, _ptr(seq_cat), _ptr(off), _ptr(ws_cat), _ptr(kind),
                                                _ptr(gate), C.byref(h)), "relem_batch_create")
        return Batch(self, h, off.copy(), len(off) - 1)

    def _estep_out(self, nseq, detail):
        NT = self.n_theta
        r = EstepResult()
        r.EN_diff = np.zeros(NT)
        o = _EstepOut()
        o.EN_diff = _dptr(r.EN_diff)
        if detail:
            r.Z = np.zeros((nseq, 3)); r.ENo = np.zeros((nseq, NT)); r.ENx = np.zeros((nseq, NT))
            r.EH = np.zeros((nseq, 4)); r.bpp_eff = np.zeros(nseq); r.skipped = np.zeros(nseq, dtype=np.uint8)
            o.Z, o.ENo, o.ENx, o.EH, o.bpp_eff = _dptr(r.Z), _dptr(r.ENo), _dptr(r.ENx), _dptr(r.EH), _dptr(r.bpp_eff)
            o.skipped = r.skipped.ctypes.data_as(C.POINTER(C.c_uint8))
        return r, o

    @staticmethod
    def _estep_fill(r, o):
        r.fn, r.sum_eff, r.n_skipped = o.fn, o.sum_eff, int(o.n_skipped)
        r.EH_diff = np.array([o.EH_diff[0], o.EH_diff[1]])
        return r

    def estep_run(self, batch, detail=False):
        r, o = self._estep_out(batch.nseq, detail)
        self._check(self.lib.relem_estep_run(self.h, batch.handle, C.byref(o)), "relem_estep_run")
        return self._estep_fill(r, o)

    def estep(self, seq_cat, off, ws_cat, kind=None, gate=None, detail=False):
        """host buffers in, host results out (the reference-facing call)."""
        # same coercion as batch(): a wrong dtype or a strided view would be reinterpreted by the C side
        seq_cat = np.ascontiguousarray(seq_cat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        ws_cat = np.ascontiguousarray(ws_cat, dtype=np.float64)
        nseq = len(off) - 1
        r, o = self._estep_out(nseq, detail)
        kind = None if kind is None else np.ascontiguousarray(kind, dtype=np.uint8)
        gate = None if gate is None else np.ascontiguousarray(gate, dtype=np.int32)
        self._check(self.lib.relem_estep(self.h, nseq, _ptr(seq_cat), _ptr(off), _ptr(ws_cat), _ptr(kind), _ptr(gate),
                                         C.byref(o)), "relem_estep")
        return self._estep_fill(r, o)

    def bpp(self, batch, want_lnbpp=True):
        nseq = batch.nseq
        moff = np.zeros(nseq + 1, dtype=np.int64)
        # sizes first
        for n in range(nseq):
            L = int(batch.off[n + 1] - batch.off[n])
            moff[n + 1] = moff[n] + (L + 1) * (min(L, self._max_span()) + 1)
        tot = int(moff[-1])
        bp = np.zeros(tot, dtype=np.uint8); lf = np.zeros(tot, dtype=np.uint8)
        ln = np.zeros(tot) if want_lnbpp else None
        eff = np.zeros(nseq); lnz = np.zeros(nseq)
        self._check(self.lib.relem_bpp(self.h, batch.handle, _ptr(moff), _ptr(bp), _ptr(lf), _ptr(ln), _ptr(eff),
                                       _ptr(lnz)), "relem_bpp")
        return moff, bp, lf, ln, eff, lnz

    def _max_span(self):
        return getattr(self, "_ms", 1 << 30)

    def _scan_out(self, nseq, tl):
        NT = self.n_theta
        r = ScanResult()
        r.PysL = np.zeros(tl); r.PyeL = np.zeros(tl + nseq); r.PyiL = np.zeros(tl)
        r.psihat = np.zeros(tl, dtype=np.int32); r.rss_buf = C.create_string_buffer(tl + 1)
        r.Ys = np.zeros(nseq, dtype=np.int32); r.Ye = np.zeros(nseq, dtype=np.int32)
        r.exist_prob = np.zeros(nseq); r.EN = np.zeros(NT); r.ZL = np.zeros(nseq)
        o = _ScanOut()
        o.PysL, o.PyeL, o.PyiL = _dptr(r.PysL), _dptr(r.PyeL), _dptr(r.PyiL)
        o.psihat = r.psihat.ctypes.data_as(C.POINTER(C.c_int32))
        o.rss = C.cast(r.rss_buf, C.c_char_p)
        o.Ys = r.Ys.ctypes.data_as(C.POINTER(C.c_int32)); o.Ye = r.Ye.ctypes.data_as(C.POINTER(C.c_int32))
        o.exist_prob, o.EN, o.ZL = _dptr(r.exist_prob), _dptr(r.EN), _dptr(r.ZL)
        return r, o

    def scan_run(self, batch):
        nseq, tl = batch.nseq, int(batch.off[-1])
        r, o = self._scan_out(nseq, tl)
        self._check(self.lib.relem_scan_run(self.h, batch.handle, C.byref(o)), "relem_scan_run")
        r.rss = r.rss_buf.raw[:tl].decode("ascii")
        r.off = batch.off
        return r

    def scan(self, seq_cat, off, ws_cat, decode_rss=True):
        """host buffers in, host results out (the reference-facing call: RNAelemScanner::scan, motif_scanner.hpp:938-949)."""
        seq_cat = np.ascontiguousarray(seq_cat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        ws_cat = np.ascontiguousarray(ws_cat, dtype=np.float64)
        nseq, tl = len(off) - 1, int(off[-1])
        r, o = self._scan_out(nseq, tl)
        self._check(self.lib.relem_scan(self.h, nseq, _ptr(seq_cat), _ptr(off), _ptr(ws_cat), C.byref(o)), "relem_scan")
        r.rss = r.rss_buf.raw[:tl].decode("ascii") if decode_rss else None
        r.off = off
        return r

    def timing(self):
        names = (C.c_char_p * 48)(); ms = (C.c_float * 48)(); ln = (C.c_int * 48)()
        n = self.lib.relem_last_timing(self.h, names, ms, ln, 48)
        return [(names[k].decode(), float(ms[k]), int(ln[k])) for k in range(n)]

    def fp64_peak(self):
        """-> (DFMA per second, exp() per second) measured on this GPU"""
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.relem_fp64_peak(self.h, C.byref(a), C.byref(b)), "relem_fp64_peak")
        return a.value, b.value

    # ---- collective
    def comm_init(self, uid_bytes, rank, nranks):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid_bytes)
        self._check(self.lib.relem_comm_init(self.h, buf, rank, nranks), "relem_comm_init")

    def allreduce_sum(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        self._check(self.lib.relem_allreduce_sum(self.h, _ptr(a), a.size), "relem_allreduce_sum")
        return a
