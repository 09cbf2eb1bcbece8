"""ctypes front-end of oracle/libcpu_ref.so (the C restatement of the reference algorithm).

TEST INFRASTRUCTURE ONLY — see the header of cpu_ref.c.  Used by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class CModel(C.Structure):
    _fields_ = [("nw", C.c_int32), ("nR", C.c_int32), ("R", C.c_void_p), ("H", C.c_void_p), ("r", C.c_void_p),
                ("wf_centers", C.c_void_p), ("a_latt", C.c_double * 9)]


class CParams(C.Structure):
    _fields_ = [("mp_grid", C.c_int32 * 3), ("do_gauge_trafo", C.c_int32), ("use_cart_velo", C.c_int32),
                ("do_mep", C.c_int32), ("do_kubo", C.c_int32), ("do_ahc", C.c_int32), ("do_opt", C.c_int32),
                ("do_write_mep_bands", C.c_int32), ("valence_bands", C.c_int32), ("n_hw", C.c_int32), ("n_eF", C.c_int32),
                ("hw_min", C.c_double), ("hw_max", C.c_double), ("eF_min", C.c_double), ("eF_max", C.c_double),
                ("T_kelvin", C.c_double), ("eta", C.c_double), ("kubo_tol", C.c_double), ("do_gyro", C.c_int32), ("posop_mode", C.c_int32)]


class CResult(C.Structure):
    _fields_ = [("n_k", C.c_int64), ("eF_idx", C.c_int32), ("n_degenerate", C.c_int64), ("Ne", C.c_void_p),
                ("mep_ic", C.c_void_p), ("mep_lc", C.c_void_p), ("mep_cs", C.c_void_p), ("mep_bands", C.c_void_p),
                ("kubo_mep_ic", C.c_void_p), ("kubo_mep_lc", C.c_void_p), ("kubo_mep_cs", C.c_void_p),
                ("ahc", C.c_void_p), ("ahc_velo", C.c_void_p), ("ohc_velo", C.c_void_p),
                ("opt_s", C.c_void_p), ("opt_a", C.c_void_p), ("photo2", C.c_void_p),
                ("gyro_c", C.c_void_p), ("gyro_d", C.c_void_p), ("gyro_dw", C.c_void_p)]


def openblas_path() -> str:
    import scipy
    cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so"))
    if not cands:
        raise RuntimeError("scipy's OpenBLAS (zheevd/zgemm for the oracle) not found")
    return sorted(cands)[0]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libcpu_ref.so")
    src = os.path.join(_HERE, "cpu_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libcpu_ref.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.cpu_ref_init.argtypes = [C.c_char_p]
        L.cpu_ref_partials_size.restype = C.c_int64
        L.cpu_ref_partials_size.argtypes = [C.POINTER(CModel), C.POINTER(CParams)]
        L.cpu_ref_accumulate.argtypes = [C.POINTER(CModel), C.POINTER(CParams), C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p]
        L.cpu_ref_accumulate_mt.argtypes = [C.POINTER(CModel), C.POINTER(CParams), C.c_int64, C.c_int64, C.c_int, C.c_void_p]
        L.cpu_ref_accumulate_abs.argtypes = [C.POINTER(CModel), C.POINTER(CParams), C.c_int64, C.c_int64, C.c_int, C.c_void_p]
        L.cpu_ref_finalize.argtypes = [C.POINTER(CModel), C.POINTER(CParams), C.c_void_p, C.POINTER(CResult)]
        L.cpu_ref_interp.argtypes = [C.POINTER(CModel), C.POINTER(CParams), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cpu_ref_band_sums.argtypes = [C.POINTER(CModel), C.POINTER(CParams), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cpu_ref_interp_pos.argtypes = [C.POINTER(CModel), C.POINTER(CParams), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = L.cpu_ref_init(openblas_path().encode())
        if rc != 0:
            raise RuntimeError(f"cpu_ref_init failed rc={rc}")
        _LIB = L
    return _LIB


def _f(a, dtype):
    return np.asfortranarray(np.asarray(a, dtype=dtype))


class Handle:
    """Keeps the numpy buffers alive next to the C structs."""

    def __init__(self, model, params):
        # model / params: objects with the attributes of oracle.mep_oracle.Model / Params
        self.nw, self.nR = int(model.nw), int(model.nR)
        self._R = _f(model.R, np.float64)
        self._H = _f(model.H, np.complex128)
        self._wfc = None if model.wf_centers is None else _f(model.wf_centers, np.float64)
        m = CModel()
        m.nw, m.nR = self.nw, self.nR
        m.R = self._R.ctypes.data
        m.H = self._H.ctypes.data
        self._r = None if getattr(model, "r", None) is None else _f(model.r, np.complex128)      # (3,nw,nw,nR) Bohr
        m.r = None if self._r is None else self._r.ctypes.data
        m.wf_centers = None if self._wfc is None else self._wfc.ctypes.data
        al = _f(model.a_latt, np.float64)
        for j in range(3):
            for i in range(3):
                m.a_latt[i + 3 * j] = al[i, j]
        self.cm = m
        p = CParams()
        for i in range(3):
            p.mp_grid[i] = int(params.mp_grid[i])
        for k in ("do_gauge_trafo", "use_cart_velo", "do_mep", "do_kubo", "do_ahc", "do_opt", "do_write_mep_bands"):
            setattr(p, k, int(bool(getattr(params, k))))
        p.valence_bands, p.n_hw, p.n_eF = int(params.valence_bands), int(params.n_hw), int(params.n_eF)
        for k in ("hw_min", "hw_max", "eF_min", "eF_max", "T_kelvin", "eta", "kubo_tol"):
            setattr(p, k, float(getattr(params, k)))
        p.do_gyro = int(bool(getattr(params, "do_gyro", False)))
        p.posop_mode = {"paper": 0, "coded": 1, 0: 0, 1: 1}[getattr(params, "posop_mode", "paper")]
        self.cp = p
        self.params = params

    def partials_size(self) -> int:
        return int(lib().cpu_ref_partials_size(C.byref(self.cm), C.byref(self.cp)))

    def accumulate(self, lin_begin, lin_end, nthreads=1, partials=None):
        if partials is None:
            partials = np.zeros(self.partials_size())
        rc = lib().cpu_ref_accumulate_mt(C.byref(self.cm), C.byref(self.cp), lin_begin, lin_end, nthreads, partials.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"cpu_ref_accumulate rc={rc}")
        return partials

    def accumulate_abs(self, lin_begin, lin_end, nthreads=1):
        """sum_k |X(k)| for every packed entry (test helper: the scale reassociation errors are measured against)."""
        buf = np.zeros(self.partials_size())
        rc = lib().cpu_ref_accumulate_abs(C.byref(self.cm), C.byref(self.cp), lin_begin, lin_end, nthreads, buf.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"cpu_ref_accumulate_abs rc={rc}")
        return buf

    def band_sums(self, e_k, V_ka, partials=None, A_ka=None, Om_kab=None):
        """The per-k response functions on supplied eigenpairs: e_k (nw,nk), V_ka (3,nw,nw,nk) (+ A_ka (3,nw,nw,nk), Om_kab
        (3,3,nw,nw,nk) with a position operator); partials += ."""
        e = _f(e_k, np.float64)
        V = _f(V_ka, np.complex128)
        A = None if A_ka is None else _f(A_ka, np.complex128)
        Om = None if Om_kab is None else _f(Om_kab, np.complex128)
        if partials is None:
            partials = np.zeros(self.partials_size())
        with np.errstate(all="ignore"):
            rc = lib().cpu_ref_band_sums(C.byref(self.cm), C.byref(self.cp), e.shape[1], e.ctypes.data, V.ctypes.data,
                                         None if A is None else A.ctypes.data, None if Om is None else Om.ctypes.data, partials.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"cpu_ref_band_sums rc={rc}")
        return partials

    def finalize(self, partials) -> dict:
        return finalize_with(lambda res: lib().cpu_ref_finalize(C.byref(self.cm), C.byref(self.cp), partials.ctypes.data, C.byref(res)),
                             self.params)

    def run_mesh(self, nthreads=1) -> dict:
        g = self.params.mp_grid
        return self.finalize(self.accumulate(0, g[0] * g[1] * g[2], nthreads))

    def interp(self, kpts, want_V=True, want_H=False):
        kp = _f(np.asarray(kpts, dtype=np.float64).T, np.float64)       # (3,nk)
        nk = kp.shape[1]
        nw = self.nw
        e = np.zeros((nw, nk), order="F")
        V = np.zeros((3, nw, nw, nk), dtype=np.complex128, order="F") if want_V else None
        Hk = np.zeros((nw, nw, nk), dtype=np.complex128, order="F") if want_H else None
        Hka = np.zeros((3, nw, nw, nk), dtype=np.complex128, order="F") if want_H else None
        rc = lib().cpu_ref_interp(C.byref(self.cm), C.byref(self.cp), nk, kp.ctypes.data, e.ctypes.data,
                                  None if V is None else V.ctypes.data, None if Hk is None else Hk.ctypes.data,
                                  None if Hka is None else Hka.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"cpu_ref_interp rc={rc}")
        return e, V, Hk, Hka


def _interp_pos(self, kpts):
    """get_wann_interp with a position operator on a k list: e (nw,nk), V (3,nw,nw,nk), A (3,nw,nw,nk), Om (3,3,nw,nw,nk)."""
    kp = _f(np.asarray(kpts, dtype=np.float64).T, np.float64)
    nk, nw = kp.shape[1], self.nw
    e = np.zeros((nw, nk), order="F")
    V = np.zeros((3, nw, nw, nk), dtype=np.complex128, order="F")
    A = np.zeros((3, nw, nw, nk), dtype=np.complex128, order="F")
    Om = np.zeros((3, 3, nw, nw, nk), dtype=np.complex128, order="F")
    rc = lib().cpu_ref_interp_pos(C.byref(self.cm), C.byref(self.cp), nk, kp.ctypes.data, e.ctypes.data, V.ctypes.data, A.ctypes.data, Om.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"cpu_ref_interp_pos rc={rc}")
    return e, V, A, Om


Handle.interp_pos = _interp_pos


def finalize_with(call, params) -> dict:
    """Allocate every output the params enable, run `call(CResult)`, return a dict keyed like the out/*.npy files."""
    n_hw, n_eF, nval = int(params.n_hw), int(params.n_eF), int(params.valence_bands)
    bufs = {"Ne": np.zeros(n_eF)}
    if params.do_mep:
        for k in ("mep_ic", "mep_lc", "mep_cs"):
            bufs[k] = np.zeros((3, 3), order="F")
        if params.do_write_mep_bands:
            bufs["mep_bands"] = np.zeros((3, 3, nval), order="F")
    if params.do_kubo:
        for k in ("kubo_mep_ic", "kubo_mep_lc", "kubo_mep_cs"):
            bufs[k] = np.zeros((3, 3), order="F")
    if params.do_ahc:
        bufs["ahc"] = np.zeros((3, 3), order="F")
        bufs["ahc_velo"] = np.zeros((3, 3, n_hw), dtype=np.complex128, order="F")
        bufs["ohc_velo"] = np.zeros((3, 3, n_hw), dtype=np.complex128, order="F")
    if params.do_opt:
        bufs["opt_s"] = np.zeros((3, 3, n_hw), dtype=np.complex128, order="F")
        bufs["opt_a"] = np.zeros((3, 3, n_hw), dtype=np.complex128, order="F")
        bufs["photo2"] = np.zeros((3, 3, 3, n_hw, n_eF), dtype=np.complex128, order="F")
    if getattr(params, "do_gyro", False):
        bufs["gyro_c"] = np.zeros((3, 3, n_eF), dtype=np.complex128, order="F")
        bufs["gyro_d"] = np.zeros((3, 3, n_eF), dtype=np.complex128, order="F")
        bufs["gyro_dw"] = np.zeros((3, 3, n_hw, n_eF), dtype=np.complex128, order="F")
    res = CResult()
    for k, v in bufs.items():
        setattr(res, k, v.ctypes.data)
    rc = call(res)
    if rc != 0:
        raise RuntimeError(f"finalize rc={rc}")
    out = {"n_k": int(res.n_k), "eF_idx": int(res.eF_idx), "n_degenerate": int(res.n_degenerate), "Ne": bufs["Ne"]}
    names = {"ahc": "ahc_tens", "ahc_velo": "ahcVELO", "ohc_velo": "ohcVELO", "opt_s": "opt_Ssymm", "opt_a": "opt_Asymm",
             "photo2": "photoC_2nd", "gyro_c": "gyro_C", "gyro_d": "gyro_D", "gyro_dw": "gyro_Dw"}
    for k, v in bufs.items():
        if k != "Ne":
            out[names.get(k, k)] = v
    if params.do_mep:
        out["mep_tens"] = out["mep_ic"] + out["mep_lc"] + out["mep_cs"]
    if params.do_kubo:
        out["kubo_mep_tens"] = out["kubo_mep_ic"] + out["kubo_mep_lc"] + out["kubo_mep_cs"]
    return out
