"""ctypes binding of libpnmol_b200.so (C ABI declared in include/pnmol_b200.h).

There is deliberately no fallback: if the shared library is missing, or no CUDA device is
present when a compute entry point is called, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PNMOL_B200_LIB") or os.path.join(_HERE, "libpnmol_b200.so")  # env override: tuning builds

c_int, c_double, c_void_p, c_int64 = ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_int64

# name -> (restype, argtypes); mirrors include/pnmol_b200.h one to one
SIGNATURES = {
    "pnmol_b200_last_error": (ctypes.c_char_p, []),
    "pnmol_b200_version": (c_int, []),
    "pnmol_b200_launch_count": (c_int64, []),
    "pnmol_b200_profile": (c_int, [c_void_p, c_int, c_void_p]),
    "pnmol_b200_path": (c_int, [c_void_p]),
    "pnmol_b200_cluster_size": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pnmol_b200_create": (c_int, [ctypes.POINTER(c_void_p)] + [c_int] * 8),
    "pnmol_b200_destroy": (c_int, [c_void_p]),
    "pnmol_b200_set_operator": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pnmol_b200_set_prior": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "pnmol_b200_set_members": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "pnmol_b200_structure": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "pnmol_b200_initialize": (c_int, [c_void_p, c_void_p, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pnmol_b200_step": (c_int, [c_void_p, c_double, c_double] + [c_void_p] * 10 + [c_int, c_void_p]),
    "pnmol_b200_run": (c_int, [c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_int] + [c_void_p] * 11 + [c_int, c_void_p]),
    "pnmol_b200_run_marginals": (c_int, [c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_int] + [c_void_p] * 9 + [c_int, c_void_p]),
    "pnmol_b200_run_adaptive": (c_int, [c_void_p, c_double, c_double, c_void_p, c_double, c_double, c_double, c_double, c_double,
                                        c_int] + [c_void_p] * 11 + [c_int, c_void_p]),
    "pnmol_b200_run_adaptive_trajectory": (c_int, [c_void_p, c_double, c_double, c_void_p, c_double, c_double, c_double, c_double,
                                                   c_double, c_int] + [c_void_p] * 16 + [c_int, c_int, c_void_p]),
    "pnmol_b200_marginal_std": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "pnmol_b200_rescale": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "pnmol_b200_simulate_final_state_host": (c_int, [c_void_p, c_void_p, c_double, c_double, c_void_p, c_void_p, c_void_p,
                                                     c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pnmol_b200_sqrt_propagate": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pnmol_b200_smoother_step": (c_int, [c_void_p] * 10 + [c_int, c_int, c_int, c_void_p]),
    "pnmol_b200_sqrt_update": (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, c_int, c_void_p]),
    "pnmol_b200_fd_coefficients": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_void_p,
                                           c_void_p, c_int, c_void_p]),
    "pnmol_b200_gram_cholesky": (c_int, [c_int, c_void_p, c_int, c_void_p, c_int, c_double, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_int, c_void_p]),
}

_lib = None


class PnmolB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PnmolB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(pnmol_b200 has no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().pnmol_b200_last_error()
        raise PnmolB200Error(f"libpnmol_b200 error {rc}: {msg.decode() if msg else '?'}")


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return c_void_p(t.data_ptr())
    return c_void_p(t.ctypes.data)


def current_stream(device):
    import torch

    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch_count():
    return int(load().pnmol_b200_launch_count())
