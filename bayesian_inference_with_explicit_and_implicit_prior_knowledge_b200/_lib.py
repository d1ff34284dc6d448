"""ctypes binding of libpgas_b200.so (include/pgas_b200.h) and its in-tree build.

The library is the product: there is no CPU fallback.  Importing this module never compiles
anything; `build()` (called by __graft_entry__.build) runs nvcc for sm_100a, and `lib()` loads
the built shared object or raises.
"""
import ctypes as C
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.path.join(_HERE, "libpgas_b200.so")
_SOURCES = ["model.cu", "sweep.cu", "weights.cu", "weights_lat.cu", "sweep_api.cu", "suffstats.cu", "mniw_draw.cu", "gp_posterior.cu", "chains.cu",
            "marginal.cu", "microbench.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]

PGAS_MAX_NX, PGAS_MAX_NY, PGAS_MAX_NU, PGAS_MAX_D, PGAS_MAX_GP = 4, 2, 4, 3, 2
LINK_IDENTITY, LINK_ATAN, LINK_TANH = 0, 1, 2
MAP_AFFINE, MAP_VEHICLE_SLIP, MAP_PROGRAM = 0, 1, 2
PGAS_MAX_PROG, PGAS_PROG_STACK = 64, 8
OPS = dict(PUSH_X=1, PUSH_U=2, PUSH_C=3, ADD=4, SUB=5, MUL=6, DIV=7, NEG=8, SIN=9, COS=10, TAN=11, TANH=12, ATAN=13, EXP=14, LOG=15,
           SQRT=16, ABS=17, POW=18, ATAN2=19, PUSH_Y=20)
FLAG_ANCESTOR_GATHER, FLAG_INPUT_PREV, FLAG_VCHOL_TRANSPOSE = 1, 2, 4


def _sources():
    return [s for s in _SOURCES if os.path.exists(os.path.join(_CSRC, s))]


def _stale():
    if not os.path.exists(_SO):
        return True
    t = os.path.getmtime(_SO)
    deps = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)] + [
        os.path.join(_HERE, "..", "include", "pgas_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def _header_abi_version():
    import re
    with open(os.path.join(_HERE, "..", "include", "pgas_b200.h")) as f:
        m = re.search(r"#define\s+PGAS_ABI_VERSION\s+(\d+)", f.read())
    return int(m.group(1)) if m else None


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a and link libpgas_b200.so in-tree (nvcc cross-compiles
    without a GPU)."""
    if not force and not _stale():
        return _SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    bdir = os.path.join(_HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    hdrs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")] + [
        os.path.join(_HERE, "..", "include", "pgas_b200.h")]
    hdr_t = max(os.path.getmtime(h) for h in hdrs)

    def one(src):
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        sp = os.path.join(_CSRC, src)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(sp)
                and os.path.getmtime(obj) > hdr_t):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("PGAS_NVCC_EXTRA", "").split() + ["-c", sp, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, _sources()))
    cmd = [nvcc, "-shared", "-o", _SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return _SO


# ------------------------------------------------------------------------------- structs
class ModelParams(C.Structure):
    _fields_ = [
        ("n_x", C.c_int32), ("n_y", C.c_int32), ("n_u", C.c_int32), ("D", C.c_int32),
        ("M", C.c_int32), ("T", C.c_int32),
        ("freq", C.POINTER(C.c_int32)),
        ("idx_start", C.c_int32), ("idx_step", C.c_int32),
        ("center", C.c_double * PGAS_MAX_D),
        ("half_width", C.c_double * PGAS_MAX_D),
        ("map_kind", C.c_int32),
        ("Az", (C.c_double * (PGAS_MAX_NX + PGAS_MAX_NU)) * PGAS_MAX_D),
        ("bz", C.c_double * PGAS_MAX_D),
        ("slip_lf", C.c_double), ("slip_lr", C.c_double),
        ("prog_len", C.c_int32), ("prog_op", C.c_int32 * PGAS_MAX_PROG), ("prog_const", C.c_double * PGAS_MAX_PROG),
        ("H", (C.c_double * PGAS_MAX_NX) * PGAS_MAX_NY),
        ("h0", C.c_double * PGAS_MAX_NY),
        ("R", (C.c_double * PGAS_MAX_NY) * PGAS_MAX_NY),
        ("observations", C.POINTER(C.c_double)),
        ("inputs", C.POINTER(C.c_double)),
        ("m0", C.c_double * PGAS_MAX_NX),
        ("P0", (C.c_double * PGAS_MAX_NX) * PGAS_MAX_NX),
        ("flags", C.c_int32),
        ("lik_prog_len", C.c_int32), ("lik_prog_op", C.c_int32 * PGAS_MAX_PROG), ("lik_prog_const", C.c_double * PGAS_MAX_PROG),
    ]


class Rng(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("seed", C.c_uint64), ("chain_base", C.c_uint32),
        ("iteration", C.c_uint32),
        ("Z", C.c_void_p), ("U", C.c_void_p), ("chi2", C.c_void_p), ("G", C.c_void_p),
        ("Nrm", C.c_void_p),
    ]


class MargProgram(C.Structure):
    _fields_ = [("len", C.c_int32), ("n_const", C.c_int32), ("ops", C.POINTER(C.c_int32)), ("consts", C.POINTER(C.c_double))]


class MargGP(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("D", C.c_int32),
        ("sqrt_eig", C.POINTER(C.c_double)),
        ("center", C.c_double * PGAS_MAX_D), ("half_width", C.c_double * PGAS_MAX_D),
        ("link", C.c_int32),
        ("gp_in", C.POINTER(C.c_double)), ("gp_post", C.POINTER(C.c_double)),
        ("eta0", C.POINTER(C.c_double)), ("eta1", C.POINTER(C.c_double)),
        ("eta2", C.c_double), ("eta3", C.c_double),
        ("xi_mean", C.c_double), ("xi_var", C.c_double),
        ("prog", MargProgram),
    ]


class MargParams(C.Structure):
    _fields_ = [
        ("n_x", C.c_int32), ("n_y", C.c_int32), ("n_gp", C.c_int32), ("T", C.c_int32),
        ("gp", MargGP * PGAS_MAX_GP),
        ("trans", C.POINTER(C.c_double)), ("outp", C.POINTER(C.c_double)),
        ("out_link", C.c_int32),
        ("observations", C.POINTER(C.c_double)),
        ("Q", (C.c_double * PGAS_MAX_NX) * PGAS_MAX_NX),
        ("R", (C.c_double * PGAS_MAX_NY) * PGAS_MAX_NY),
        ("m0", C.c_double * PGAS_MAX_NX),
        ("P0", (C.c_double * PGAS_MAX_NX) * PGAS_MAX_NX),
        ("n_u", C.c_int32),
        ("inputs", C.POINTER(C.c_double)),
        ("trans_prog", MargProgram), ("outp_prog", MargProgram),
    ]


class MargRng(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("seed", C.c_uint64), ("chain_base", C.c_uint32), ("iteration", C.c_uint32),
        ("Z", C.c_void_p), ("ZXI0", C.c_void_p), ("U", C.c_void_p), ("TS", C.c_void_p),
    ]


_PP = C.POINTER(C.c_void_p)      # array of device pointers (double* const*)

EXPORTS = {
    # name: (restype, argtypes)
    "pgas_last_error": (C.c_char_p, []),
    "pgas_version": (C.c_int, []),
    "pgas_abi_version": (C.c_int, []),
    "pgas_device_count": (C.c_int, []),
    "pgas_launch_count": (C.c_longlong, []),
    "pgas_model_create": (C.c_int, [C.POINTER(ModelParams), C.POINTER(C.c_void_p)]),
    "pgas_model_destroy": (C.c_int, [C.c_void_p]),
    "pgas_model_jmax": (C.c_int, [C.c_void_p]),
    "pgas_resample_f64": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pgas_hgp_eval_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "pgas_csmc_step_f64": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 10 + [C.c_int32, C.c_void_p]),
    "pgas_csmc_sweep_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32]),
    "pgas_csmc_sweep_f64": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.POINTER(Rng), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pgas_reconstruct_trajectory_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                                  C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "pgas_suffstats_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pgas_mniw_draw_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "pgas_mniw_draw_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_int32,
                                     C.c_int32, C.POINTER(Rng), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
    "pgas_mniw_posterior_batch_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "pgas_mniw_posterior_batch_f64": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pgas_run_chains_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32]),
    "pgas_run_chains_f64": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_double, C.c_void_p, C.POINTER(Rng), C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pgas_philox_sweep_variates_f64": (C.c_int, [C.POINTER(Rng), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "pgas_marg_model_create": (C.c_int, [C.POINTER(MargParams), C.POINTER(C.c_void_p)]),
    "pgas_marg_model_destroy": (C.c_int, [C.c_void_p]),
    "pgas_marg_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32]),
    "pgas_marg_filter_f64": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.POINTER(MargRng), C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, _PP, _PP, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pgas_marg_refstats_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, _PP,
                                         C.c_void_p]),
    "pgas_marg_csmc_f64": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, _PP, C.POINTER(MargRng), C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
    "pgas_marg_run_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32]),
    "pgas_marg_run_f64": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(MargRng),
                                    C.c_void_p, C.c_void_p, _PP, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pgas_marg_outputs_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pgas_mniw_log_base_measure_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                                 C.c_void_p]),
    "pgas_philox_marg_variates_f64": (C.c_int, [C.POINTER(MargRng), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pgas_measure_fp64_peaks": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]),
}

# developer / measurement exports that are not part of include/pgas_b200.h
DEBUG_EXPORTS = {
    "pgas_debug_state_kernel_f64": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Rng),
                                              C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
}

_lib = None


def lib():
    """The loaded library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(
                f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the PGAS hot path)")
        if _stale():
            import warnings
            warnings.warn(f"{_SO} is older than csrc/ or include/pgas_b200.h: run __graft_entry__.build() to rebuild", RuntimeWarning)
        # developer override (A/B measurements of compile-time variants): PGAS_LIB_PATH=<another build of this library>
        L = C.CDLL(os.environ.get("PGAS_LIB_PATH") or _SO)
        want = _header_abi_version()
        try:
            L.pgas_abi_version.restype = C.c_int
            got = L.pgas_abi_version()
        except AttributeError:
            got = None
        if got != want:
            raise RuntimeError(f"{_SO} was built against ABI {got}, include/pgas_b200.h declares {want}: rebuild with "
                               "__graft_entry__.build() (ctypes layouts and the binary must come from the same header)")
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)          # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        for name, (res, args) in DEBUG_EXPORTS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class PgasError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise PgasError(f"libpgas_b200 error {rc}: {lib().pgas_last_error().decode()}")


def require_cuda():
    import torch
    if not torch.cuda.is_available() or lib().pgas_device_count() < 1:
        raise PgasError("no CUDA device: the PGAS hot path has no CPU fallback")
    return torch


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr_array(tensors):
    """ctypes array of device pointers (double* const*) for a list of CUDA tensors; None -> NULL array pointer"""
    if tensors is None:
        return C.cast(None, _PP)
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return C.cast(arr, _PP)


def ptr(t):
    """device pointer of a CUDA tensor (or NULL for None)"""
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return C.c_void_p(t.data_ptr())
