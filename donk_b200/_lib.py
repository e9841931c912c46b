"""ctypes binding of the CUDA library ``donk_b200/csrc/libdonk_b200.so`` (C ABI: include/donk_b200.h).

There is no CPU path: :func:`native` raises if the library has not been built or no CUDA
device is present.  Tests that exercise the host-side logic without a GPU inject the
host-emulation build of the very same kernel sources through :func:`install_test_backend`
(tests/emu/); nothing in the package does that on its own.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_float, c_int, c_size_t, c_void_p
from pathlib import Path

import numpy as np
import torch

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "csrc" / "libdonk_b200.so"

DONK_OK, DONK_BAD_SHAPE, DONK_NOT_PD, DONK_CUDA_ERROR, DONK_SMEM_LIMIT = range(5)
ILQR_BACKWARD, ILQR_FORWARD, ILQR_KL, ILQR_COST = 1, 2, 4, 8

P = c_void_p  # every array argument is a raw pointer

# name -> argument ctypes (return type is int unless listed in _RESTYPES)
SIGNATURES = {
    "ilqr_step_f64": [c_int] * 6 + [P] * 15 + [c_double, c_double] + [P] * 10 + [P],
    "ilqr_step_f32": [c_int] * 6 + [P] * 15 + [c_float, c_float] + [P] * 10 + [P],
    "extended_costs_kl_f64": [c_int] * 3 + [P] * 5 + [P],
    "spd_factor_f64": [c_int] * 2 + [P] * 5 + [P],
    "kl_divergence_action_f64": [c_int] * 4 + [P] * 10 + [P],
    "brent_update_f64": [c_int, c_int, c_double, c_double, P, P, P, c_double, c_double, c_double, c_int]
    + [P] * 5 + [P],
    "ilqr_optimize_f64": [c_int] * 4 + [P] * 15 + [c_double] * 7 + [c_int] + [P] * 14 + [c_size_t, P],
    "quadratic_cost_l2_f64": [c_int, c_int] + [P] * 5 + [P],
    "quadratic_cost_l1_f64": [c_int, c_int, P, P, P, c_double, P, P, P, P],
    "dynamics_log_prob_f64": [c_int] * 5 + [P] * 7 + [P],
    "fit_lr_f64": [c_int] * 5 + [P, P, c_int, P, P, P, P, c_double, c_double, c_double, P, P, P, P, P],
    "fit_lr_f32": [c_int] * 5 + [P, P, c_int, P, P, P, P, c_float, c_float, c_float, P, P, P, P, P],
    "fit_lr_niw_f64": [c_int] * 5 + [P] * 6 + [c_double, c_double, P, P, P, P, P],
    "to_transitions_f64": [c_int] * 4 + [P, P, P, P],
    "state_fit_f64": [c_int] * 3 + [P, c_size_t, c_size_t, P, P, P],
    "gmm_em_step_f64": [c_int] * 3 + [P] * 7 + [P, c_size_t, P],
    "gmm_em_step_mode_f64": [c_int] * 3 + [P] * 7 + [P, c_size_t, c_int, P],
    "gmm_m_finalize_f64": [c_double, c_int, c_int, P, P, c_double, P, P, P, P, P, P, P],
    "gmm_tc_prepare_f32": [c_int, c_int, P, P, P, c_size_t, P],
    "gmm_tc_estep_f32": [c_int] * 3 + [P] * 8 + [c_size_t, P],
    "gmm_tc_estep_m_f32": [c_int] * 3 + [P] * 8 + [c_size_t, P, c_size_t, P],
    "gmm_tc_mprepare_f32": [c_int, c_int, P, P, P, c_size_t, P],
    "gmm_tc_mstep_f32": [c_int] * 3 + [P] * 7 + [c_size_t, P],
    "gmm_predict_proba_f64": [c_int] * 3 + [P] * 5 + [P],
    "gmm_prec_chol_from_precisions_f64": [c_int] * 2 + [P] * 3 + [P],
    "fp64_probe": [c_int, c_int, c_int, P, ctypes.POINTER(c_double), P],
    "kmeans_pass_f64": [c_int] * 3 + [P] * 6 + [P, c_size_t, P],
    "kmeanspp_round_f64": [c_int] * 3 + [P] * 5 + [P, c_size_t, P],
}
_SIZE_FUNCS = {
    "gmm_stats_size": [c_int, c_int],
    "gmm_workspace_bytes": [c_int, c_int, c_int],
    "kmeans_workspace_bytes": [c_int, c_int],
    "ilqr_optimize_workspace_bytes": [c_int],
    "gmm_tc_supported": [c_int, c_int],
    "gmm_tc_planes_bytes": [c_int, c_int],
    "gmm_tc_workspace_bytes": [c_int, c_int, c_int],
    "gmm_tc_m_supported": [c_int, c_int],
    "gmm_tc_mplanes_bytes": [c_int, c_int],
    "gmm_tc_mwork_bytes": [c_int, c_int, c_int],
}


class DonkError(RuntimeError):
    pass


class Native:
    """One loaded build of the kernels: the CUDA library, or (tests only) the host emulation."""

    def __init__(self, path, prefix: str, device_type: str):
        self.path = str(path)
        self.prefix = prefix
        self.device_type = device_type
        self.lib = ctypes.CDLL(self.path)
        self._fn = {}
        for name, argtypes in SIGNATURES.items():
            fn = getattr(self.lib, prefix + name, None)
            if fn is None:
                continue
            fn.argtypes = argtypes
            fn.restype = c_int
            self._fn[name] = fn
        for name, argtypes in _SIZE_FUNCS.items():
            fn = getattr(self.lib, prefix + name, None)
            if fn is None:
                continue
            fn.argtypes = argtypes
            fn.restype = c_size_t
            self._fn[name] = fn

    # -- helpers -------------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        if self.device_type == "cuda":
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def stream(self):
        if self.device_type == "cuda":
            return c_void_p(torch.cuda.current_stream().cuda_stream)
        return None

    def _ptr(self, t):
        if t is None:
            return None
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"expected a tensor, got {type(t)}")
        if t.device.type != self.device_type:
            raise DonkError(f"tensor on {t.device}, library runs on {self.device_type}")
        if not t.is_contiguous():
            raise DonkError("tensor must be contiguous")
        return c_void_p(t.data_ptr())

    def call(self, name: str, *args):
        """Call ``<prefix><name>``; tensors become pointers, the stream argument is appended."""
        fn = self._fn.get(name)
        if fn is None:
            raise DonkError(f"{self.prefix}{name} is not exported by {self.path}")
        conv = [self._ptr(a) if (isinstance(a, torch.Tensor) or a is None) else a for a in args]
        rc = fn(*conv, self.stream())
        if rc != DONK_OK:
            self._raise(name, rc)

    def size(self, name: str, *args) -> int:
        return int(self._fn[name](*args))

    def _raise(self, name, rc):
        msg = ""
        if hasattr(self.lib, self.prefix + "last_error"):
            f = getattr(self.lib, self.prefix + "last_error")
            f.restype = ctypes.c_char_p
            msg = (f() or b"").decode()
        kind = {1: "bad shape", 2: "not positive definite", 3: "CUDA error", 4: "shared-memory limit"}.get(rc, rc)
        raise DonkError(f"{self.prefix}{name} failed: {kind} {msg}".strip())

    def launch_count(self) -> int:
        f = getattr(self.lib, self.prefix + "launch_count", None)
        if f is None:
            return 0
        f.restype = ctypes.c_ulonglong
        return int(f())

    def exported(self, name: str) -> bool:
        return hasattr(self.lib, self.prefix + name)


_native: Native | None = None


def native() -> Native:
    """The CUDA library.  Fails loudly if it is missing or there is no GPU (no CPU fallback)."""
    global _native
    if _native is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a). donk_b200 has no CPU fallback."
            )
        if not torch.cuda.is_available():
            raise DonkError("donk_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
        _native = Native(LIB_PATH, "donk_", "cuda")
    return _native


def install_test_backend(backend: Native | None) -> None:
    """TESTS ONLY: route calls to the host-emulation build of the kernels (tests/emu)."""
    global _native
    _native = backend


def load_library_symbols_only() -> ctypes.CDLL:
    """dlopen the CUDA library without touching a device (CPU-side export check)."""
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} not built")
    return ctypes.CDLL(str(LIB_PATH))


def as_device(x, dtype=torch.float64) -> torch.Tensor:
    """numpy / tensor -> contiguous tensor of ``dtype`` on the library's device."""
    nat = native()
    if isinstance(x, torch.Tensor):
        return x.to(device=nat.device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(device=nat.device, dtype=dtype).contiguous()
