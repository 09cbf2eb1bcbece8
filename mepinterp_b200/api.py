"""ctypes binding of libmepb.so — the C ABI declared in include/mepb.h.

This is plumbing only: it marshals numpy arrays into the plain-pointer structs of the ABI.  All compute happens in
the hand-written sm_100a kernels behind `mepb_*`; if the shared library is missing or no GPU is present the calls
fail loudly (there is no CPU path in the product).
"""
from __future__ import annotations

import ctypes as C
import os
from types import SimpleNamespace

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmepb.so")
_LIB = None

OK = 0
ERR_ARG, ERR_CUDA, ERR_NOT_HERMITIAN, ERR_UNSUPPORTED, ERR_EIG, ERR_STATE, ERR_NCCL = -1, -2, -3, -4, -5, -6, -7

EXPORTS = ["mepb_create", "mepb_destroy", "mepb_run_mesh", "mepb_reset", "mepb_accumulate", "mepb_accumulate_strided",
           "mepb_chunk_size", "mepb_balanced_block", "mepb_comm_unique_id", "mepb_comm_init_rank", "mepb_comm_size", "mepb_comm_rank",
           "mepb_allreduce_partials", "mepb_get_timing", "mepb_create_multi", "mepb_multi_run_mesh", "mepb_multi_update_model",
           "mepb_multi_size", "mepb_multi_ctx", "mepb_multi_destroy", "mepb_partials_size",
           "mepb_get_partials", "mepb_set_partials", "mepb_partials_device_ptr", "mepb_finalize", "mepb_interp",
           "mepb_interp_pos", "mepb_fourier", "mepb_fourier_mesh", "mepb_band_sums", "mepb_update_model", "mepb_update_params", "mepb_model_info", "mepb_selftest_math", "mepb_timer_start", "mepb_timer_stop", "mepb_set_profiling",
           "mepb_get_profile", "mepb_work_items", "mepb_last_kernel_ms", "mepb_kernel_launches", "mepb_strerror", "mepb_last_error", "mepb_version"]


class MepbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmepb error {code}: {msg}")
        self.code = code


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


def lib():
    """Load libmepb.so (built in-tree by mepinterp_b200/build.py).  Raises if it is missing — no fallback."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise MepbError(ERR_CUDA, f"{LIB_PATH} not built; run `python -m mepinterp_b200.build` (needs nvcc)")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.mepb_create.argtypes = [C.POINTER(CModel), C.POINTER(CParams), C.c_int, C.POINTER(vp)]
        L.mepb_destroy.argtypes = [vp]
        L.mepb_destroy.restype = None
        L.mepb_run_mesh.argtypes = [vp, C.POINTER(CResult)]
        L.mepb_reset.argtypes = [vp]
        L.mepb_accumulate.argtypes = [vp, C.c_int64, C.c_int64]
        L.mepb_accumulate_strided.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int]
        L.mepb_chunk_size.argtypes = [vp]
        L.mepb_chunk_size.restype = C.c_int64
        L.mepb_balanced_block.argtypes = [vp, C.c_int64, C.c_int]
        L.mepb_balanced_block.restype = C.c_int64
        L.mepb_comm_unique_id.argtypes = [vp, C.c_int]
        L.mepb_comm_init_rank.argtypes = [vp, C.c_int, C.c_int, vp]
        L.mepb_comm_size.argtypes = [vp]
        L.mepb_comm_rank.argtypes = [vp]
        L.mepb_allreduce_partials.argtypes = [vp]
        L.mepb_get_timing.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.mepb_create_multi.argtypes = [C.POINTER(CModel), C.POINTER(CParams), C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]
        L.mepb_multi_run_mesh.argtypes = [vp, C.POINTER(CResult)]
        L.mepb_multi_update_model.argtypes = [vp, C.POINTER(CModel)]
        L.mepb_multi_size.argtypes = [vp]
        L.mepb_multi_ctx.argtypes = [vp, C.c_int]
        L.mepb_multi_ctx.restype = vp
        L.mepb_multi_destroy.argtypes = [vp]
        L.mepb_multi_destroy.restype = None
        L.mepb_partials_size.argtypes = [vp]
        L.mepb_partials_size.restype = C.c_int64
        L.mepb_get_partials.argtypes = [vp, vp]
        L.mepb_set_partials.argtypes = [vp, vp]
        L.mepb_partials_device_ptr.argtypes = [vp, C.POINTER(vp)]
        L.mepb_finalize.argtypes = [vp, C.POINTER(CResult)]
        L.mepb_interp.argtypes = [vp, C.c_int64, vp, vp, vp]
        L.mepb_interp_pos.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp]
        L.mepb_fourier.argtypes = [vp, C.c_int64, vp, vp, vp]
        L.mepb_fourier_mesh.argtypes = [vp, C.c_int64, C.c_int64, vp, vp]
        L.mepb_band_sums.argtypes = [vp, C.c_int64, vp, vp]
        L.mepb_update_model.argtypes = [vp, C.POINTER(CModel)]
        L.mepb_update_params.argtypes = [vp, C.POINTER(CParams)]
        L.mepb_selftest_math.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.mepb_model_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int]
        L.mepb_timer_start.argtypes = [vp]
        L.mepb_timer_stop.argtypes = [vp, C.POINTER(C.c_double)]
        L.mepb_set_profiling.argtypes = [vp, C.c_int]
        L.mepb_get_profile.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.mepb_work_items.argtypes = [vp, C.c_int]
        L.mepb_work_items.restype = C.c_int64
        L.mepb_last_kernel_ms.argtypes = [vp]
        L.mepb_last_kernel_ms.restype = C.c_double
        L.mepb_kernel_launches.argtypes = [vp]
        L.mepb_kernel_launches.restype = C.c_int64
        L.mepb_strerror.argtypes = [C.c_int]
        L.mepb_strerror.restype = C.c_char_p
        L.mepb_last_error.restype = C.c_char_p
        L.mepb_version.restype = C.c_char_p
        _LIB = L
    return _LIB


def _check(rc):
    if rc != OK:
        L = lib()
        raise MepbError(rc, f"{L.mepb_strerror(rc).decode()} — {L.mepb_last_error().decode()}")


def _f(a, dtype):
    return np.asfortranarray(np.asarray(a, dtype=dtype))


def make_params(**kw) -> SimpleNamespace:
    """Parameter bundle with the reference's names (src/input_paras.f90:49-68), energies already in Hartree."""
    d = dict(mp_grid=(4, 4, 4), do_gauge_trafo=True, use_cart_velo=False, do_mep=False, do_kubo=False, do_ahc=False,
             do_opt=False, do_write_mep_bands=True, valence_bands=1, n_hw=1, n_eF=1, hw_min=0.0, hw_max=0.0,
             eF_min=0.0, eF_max=0.0, T_kelvin=0.0, eta=0.0, kubo_tol=1e-5, do_gyro=False, posop_mode="paper")
    d.update(kw)
    return SimpleNamespace(**d)


def make_model(R, H, a_latt, wf_centers=None, r=None) -> SimpleNamespace:
    """R (3,nR) fractional; H (nw,nw,nR) complex Hartree; a_latt (3,3) rows = a_i; wf_centers (3,nw) or None;
    r (3,nw,nw,nR) complex position operator in Bohr or None."""
    H = np.asarray(H)
    return SimpleNamespace(nw=H.shape[0], nR=H.shape[2], R=np.asarray(R, dtype=np.float64), H=H,
                           a_latt=np.asarray(a_latt, dtype=np.float64), wf_centers=wf_centers, r=r)


POSOP_MODES = {"paper": 0, "coded": 1, 0: 0, 1: 1}


def comm_unique_id() -> bytes:
    """128-byte NCCL unique id (rank 0 creates it, the host broadcasts it, every rank passes it to Context.comm_init_rank)."""
    buf = C.create_string_buffer(128)
    _check(lib().mepb_comm_unique_id(buf, 128))
    return buf.raw


class _Marshal:
    """numpy model / parameter bundle -> the plain-pointer structs of the ABI (keeps the arrays alive)."""

    def _marshal(self, model, params):
        self.nw, self.nR = int(model.nw), int(model.nR)
        self._R = _f(model.R, np.float64)
        self._H = _f(model.H, np.complex128)
        self._wfc = None if model.wf_centers is None else _f(model.wf_centers, np.float64)
        m = CModel()
        m.nw, m.nR = self.nw, self.nR
        self._r = None if getattr(model, "r", None) is None else _f(model.r, np.complex128)
        if self._r is not None and self._r.shape != (3, self.nw, self.nw, self.nR):
            raise ValueError("position operator must be (3, nw, nw, nR)")
        m.R, m.H, m.r = self._R.ctypes.data, self._H.ctypes.data, (None if self._r is None else self._r.ctypes.data)
        m.wf_centers = None if self._wfc is None else self._wfc.ctypes.data
        al = np.asarray(model.a_latt, dtype=np.float64)
        for j in range(3):
            for i in range(3):
                m.a_latt[i + 3 * j] = al[i, j]
        p = CParams()
        for i in range(3):
            p.mp_grid[i] = int(params.mp_grid[i])
        for k in ("do_gauge_trafo", "use_cart_velo", "do_mep", "do_kubo", "do_ahc", "do_opt", "do_write_mep_bands"):
            setattr(p, k, int(bool(getattr(params, k))))
        p.valence_bands, p.n_hw, p.n_eF = int(params.valence_bands), int(params.n_hw), int(params.n_eF)
        for k in ("hw_min", "hw_max", "eF_min", "eF_max", "T_kelvin", "eta", "kubo_tol"):
            setattr(p, k, float(getattr(params, k)))
        p.do_gyro = int(bool(getattr(params, "do_gyro", False)))
        p.posop_mode = POSOP_MODES[getattr(params, "posop_mode", "paper")]
        self.cm, self.cp, self.params = m, p, params

    def _result(self, call) -> dict:
        P = self.params
        n_hw, n_eF, nval = int(P.n_hw), int(P.n_eF), int(P.valence_bands)
        bufs = {"Ne": np.zeros(n_eF)}
        if P.do_mep:
            for k in ("mep_ic", "mep_lc", "mep_cs"):
                bufs[k] = np.zeros((3, 3), order="F")
            if P.do_write_mep_bands:
                bufs["mep_bands"] = np.zeros((3, 3, nval), order="F")
        if P.do_kubo:
            for k in ("kubo_mep_ic", "kubo_mep_lc", "kubo_mep_cs"):
                bufs[k] = np.zeros((3, 3), order="F")
        if P.do_ahc:
            bufs["ahc"] = np.zeros((3, 3), order="F")
            bufs["ahc_velo"] = np.zeros((3, 3, n_hw), dtype=np.complex128, order="F")
            bufs["ohc_velo"] = np.zeros((3, 3, n_hw), dtype=np.complex128, order="F")
        if P.do_opt:
            bufs["opt_s"] = np.zeros((3, 3, n_hw), dtype=np.complex128, order="F")
            bufs["opt_a"] = np.zeros((3, 3, n_hw), dtype=np.complex128, order="F")
            bufs["photo2"] = np.zeros((3, 3, 3, n_hw, n_eF), dtype=np.complex128, order="F")
        if getattr(P, "do_gyro", False):
            bufs["gyro_c"] = np.zeros((3, 3, n_eF), dtype=np.complex128, order="F")
            bufs["gyro_d"] = np.zeros((3, 3, n_eF), dtype=np.complex128, order="F")
            bufs["gyro_dw"] = np.zeros((3, 3, n_hw, n_eF), dtype=np.complex128, order="F")
        res = CResult()
        for k, v in bufs.items():
            setattr(res, k, v.ctypes.data)
        _check(call(res))
        out = {"n_k": int(res.n_k), "eF_idx": int(res.eF_idx), "n_degenerate": int(res.n_degenerate), "Ne": bufs["Ne"]}
        names = {"ahc": "ahc_tens", "ahc_velo": "ahcVELO", "ohc_velo": "ohcVELO", "opt_s": "opt_Ssymm", "opt_a": "opt_Asymm",
                 "photo2": "photoC_2nd", "gyro_c": "gyro_C", "gyro_d": "gyro_D", "gyro_dw": "gyro_Dw"}     # file names of file_io.f90:266-398
        for k, v in bufs.items():
            if k != "Ne":
                out[names.get(k, k)] = v
        if P.do_mep:
            out["mep_tens"] = out["mep_ic"] + out["mep_lc"] + out["mep_cs"]
        if P.do_kubo:
            out["kubo_mep_tens"] = out["kubo_mep_ic"] + out["kubo_mep_lc"] + out["kubo_mep_cs"]
        return out


class MultiContext(_Marshal):
    """mepb_multi: one process driving several B200s (a context per device + NCCL all-reduce inside the library)."""

    def __init__(self, model, params, devices):
        self._marshal(model, params)
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        self._h = C.c_void_p()
        _check(lib().mepb_create_multi(C.byref(self.cm), C.byref(self.cp), len(devices), devs, C.byref(self._h)))

    def run_mesh(self) -> dict:
        return self._result(lambda res: lib().mepb_multi_run_mesh(self._h, C.byref(res)))

    def update_model(self, H):
        self._H = _f(H, np.complex128)
        self.cm.H = self._H.ctypes.data
        _check(lib().mepb_multi_update_model(self._h, C.byref(self.cm)))

    def timings(self):
        out = []
        for i in range(int(lib().mepb_multi_size(self._h))):
            a, b, c = C.c_double(), C.c_double(), C.c_double()
            _check(lib().mepb_get_timing(lib().mepb_multi_ctx(self._h, i), C.byref(a), C.byref(b), C.byref(c)))
            out.append({"accumulate_ms": a.value, "allreduce_ms": b.value, "finalize_ms": c.value})
        return out

    def close(self):
        if self._h:
            lib().mepb_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context(_Marshal):
    """One mepb_ctx: a model + parameter set resident on one B200."""

    def __init__(self, model, params, device: int = 0):
        self._marshal(model, params)
        self._h = C.c_void_p()
        _check(lib().mepb_create(C.byref(self.cm), C.byref(self.cp), int(device), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().mepb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- mesh runs
    @property
    def n_kpts(self) -> int:
        g = self.params.mp_grid
        return int(g[0]) * int(g[1]) * int(g[2])

    def reset(self):
        _check(lib().mepb_reset(self._h))

    def accumulate(self, lin_begin: int, lin_end: int):
        _check(lib().mepb_accumulate(self._h, int(lin_begin), int(lin_end)))

    def accumulate_strided(self, lin_begin: int, lin_end: int, block: int, rank: int, nranks: int):
        _check(lib().mepb_accumulate_strided(self._h, int(lin_begin), int(lin_end), int(block), int(rank), int(nranks)))

    @property
    def chunk_size(self) -> int:
        return int(lib().mepb_chunk_size(self._h))

    def balanced_block(self, nk: int, nranks: int) -> int:
        return int(lib().mepb_balanced_block(self._h, int(nk), int(nranks)))

    def comm_init_rank(self, nranks: int, rank: int, unique_id: bytes):
        assert len(unique_id) == 128
        _check(lib().mepb_comm_init_rank(self._h, int(nranks), int(rank), C.c_char_p(unique_id)))

    @property
    def comm_size(self) -> int:
        return int(lib().mepb_comm_size(self._h))

    def allreduce_partials(self):
        """pack + NCCL all-reduce of the packed partial sums inside the library (on its launch stream)."""
        _check(lib().mepb_allreduce_partials(self._h))

    def timings(self) -> dict:
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _check(lib().mepb_get_timing(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"accumulate_ms": a.value, "allreduce_ms": b.value, "finalize_ms": c.value}

    def partials_size(self) -> int:
        return int(lib().mepb_partials_size(self._h))

    def get_partials(self, out=None) -> np.ndarray:
        if out is None:
            out = np.zeros(self.partials_size())
        _check(lib().mepb_get_partials(self._h, out.ctypes.data))
        return out

    def set_partials(self, buf: np.ndarray):
        buf = np.ascontiguousarray(buf, dtype=np.float64)
        assert buf.size == self.partials_size()
        _check(lib().mepb_set_partials(self._h, buf.ctypes.data))

    def partials_device_ptr(self) -> int:
        p = C.c_void_p()
        _check(lib().mepb_partials_device_ptr(self._h, C.byref(p)))
        return int(p.value)

    def finalize(self) -> dict:
        return self._result(lambda res: lib().mepb_finalize(self._h, C.byref(res)))

    def run_mesh(self) -> dict:
        return self._result(lambda res: lib().mepb_run_mesh(self._h, C.byref(res)))

    # ---- explicit k lists
    def interp(self, kpts, want_V: bool = True):
        kp = _f(np.asarray(kpts, dtype=np.float64).T, np.float64)        # (3,nk)
        nk, nw = kp.shape[1], self.nw
        e = np.zeros((nw, nk), order="F")
        V = np.zeros((3, nw, nw, nk), dtype=np.complex128, order="F") if want_V else None
        _check(lib().mepb_interp(self._h, nk, kp.ctypes.data, e.ctypes.data, None if V is None else V.ctypes.data))
        return e, V

    def interp_pos(self, kpts):
        """get_wann_interp with a position operator: e (nw,nk), V (3,nw,nw,nk) connection-corrected, A (3,nw,nw,nk) H gauge,
        om (3,nw,nk) curvature pseudo-vector per band."""
        kp = _f(np.asarray(kpts, dtype=np.float64).T, np.float64)
        nk, nw = kp.shape[1], self.nw
        e = np.zeros((nw, nk), order="F")
        V = np.zeros((3, nw, nw, nk), dtype=np.complex128, order="F")
        A = np.zeros((3, nw, nw, nk), dtype=np.complex128, order="F")
        om = np.zeros((3, nw, nk), dtype=np.complex128, order="F")
        _check(lib().mepb_interp_pos(self._h, nk, kp.ctypes.data, e.ctypes.data, V.ctypes.data, A.ctypes.data, om.ctypes.data))
        return e, V, A, om

    def fourier(self, kpts):
        kp = _f(np.asarray(kpts, dtype=np.float64).T, np.float64)
        nk, nw = kp.shape[1], self.nw
        Hk = np.zeros((nw, nw, nk), dtype=np.complex128, order="F")
        Hka = np.zeros((3, nw, nw, nk), dtype=np.complex128, order="F")
        _check(lib().mepb_fourier(self._h, nk, kp.ctypes.data, Hk.ctypes.data, Hka.ctypes.data))
        return Hk, Hka

    def fourier_mesh(self, lin_begin: int, nk: int):
        """Fourier sums on the mesh points [lin_begin, lin_begin + nk) through the mesh code path (general path only)."""
        nw = self.nw
        Hk = np.zeros((nw, nw, nk), dtype=np.complex128, order="F")
        Hka = np.zeros((3, nw, nw, nk), dtype=np.complex128, order="F")
        _check(lib().mepb_fourier_mesh(self._h, lin_begin, nk, Hk.ctypes.data, Hka.ctypes.data))
        return Hk, Hka

    def band_sums(self, e_k, V_ka):
        """Response functions on supplied eigenpairs e_k (nw,nk), V_ka (3,nw,nw,nk): adds to the accumulators (reset first)."""
        e = _f(e_k, np.float64)
        V = _f(V_ka, np.complex128)
        assert e.shape[0] == self.nw and V.shape == (3, self.nw, self.nw, e.shape[1])
        _check(lib().mepb_band_sums(self._h, e.shape[1], e.ctypes.data, V.ctypes.data))

    def update_model(self, H, r=None):
        """Re-upload H(R) (nw,nw,nR complex, Hartree) from host memory (and r(R) for a position-operator context; kept when None);
        same R set and FT convention."""
        self._H = _f(H, np.complex128)
        assert self._H.shape == (self.nw, self.nw, self.nR)
        self.cm.H = self._H.ctypes.data
        if r is not None:
            self._r = _f(r, np.complex128)
            self.cm.r = self._r.ctypes.data
        _check(lib().mepb_update_model(self._h, C.byref(self.cm)))

    def update_params(self, **kw):
        """hw_min / hw_max / eF_min / eF_max / T_kelvin / eta / kubo_tol (Hartree, K) of a live context."""
        import copy
        P = copy.copy(self.params)
        for k, v in kw.items():
            if not hasattr(P, k):
                raise KeyError(k)
            setattr(P, k, v)
        p = CParams.from_buffer_copy(self.cp)
        for k in ("hw_min", "hw_max", "eF_min", "eF_max", "T_kelvin", "eta", "kubo_tol"):
            setattr(p, k, float(getattr(P, k)))
        p.valence_bands, p.n_hw, p.n_eF = int(P.valence_bands), int(P.n_hw), int(P.n_eF)
        for i in range(3):
            p.mp_grid[i] = int(P.mp_grid[i])
        for k in ("do_gauge_trafo", "use_cart_velo", "do_mep", "do_kubo", "do_ahc", "do_opt", "do_write_mep_bands"):
            setattr(p, k, int(bool(getattr(P, k))))
        p.do_gyro = int(bool(getattr(P, "do_gyro", False)))
        p.posop_mode = POSOP_MODES[getattr(P, "posop_mode", "paper")]
        _check(lib().mepb_update_params(self._h, C.byref(p)))
        self.cp, self.params = p, P

    def model_info(self) -> dict:
        h, s, pth = C.c_int(), C.c_int(), C.c_int()
        why = C.create_string_buffer(256)
        _check(lib().mepb_model_info(self._h, C.byref(h), C.byref(s), C.byref(pth), why, 256))
        return {"herm_closed": bool(h.value), "sparse": bool(s.value), "path": ("sub", "warp", "general")[pth.value],
                "why": why.value.decode()}

    def selftest_math(self, x):
        """(1/x by two Newton steps, 1/x by one cubic step, 1/sqrt(x) by one cubic step) as the kernels compute them."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        a, b, c = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
        _check(lib().mepb_selftest_math(self._h, x.size, x.ctypes.data, a.ctypes.data, b.ctypes.data, c.ctypes.data))
        return a, b, c

    def timer_start(self):
        _check(lib().mepb_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        _check(lib().mepb_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    PROF_KINDS = ("interp", "sweep_ahc", "sweep_photo", "reduce", "eig", "fourier", "eig_rot", "tridiag", "tql", "zrot", "posop")

    def set_profiling(self, on: bool):
        _check(lib().mepb_set_profiling(self._h, int(bool(on))))

    def get_profile(self) -> dict:
        out = {}
        for i, name in enumerate(self.PROF_KINDS):
            ms, n = C.c_double(), C.c_int64()
            _check(lib().mepb_get_profile(self._h, i, C.byref(ms), C.byref(n)))
            out[name] = {"ms": float(ms.value), "launches": int(n.value)}
        return out

    def work_items(self) -> dict:
        return {"sweep_ahc": int(lib().mepb_work_items(self._h, 1)), "sweep_photo": int(lib().mepb_work_items(self._h, 2))}

    @property
    def last_kernel_ms(self) -> float:
        return float(lib().mepb_last_kernel_ms(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(lib().mepb_kernel_launches(self._h))
