"""ctypes binding of libbarefoot_sm100.so (C ABI: include/barefoot_sm100.h).

This is the whole "extension": torch only supplies device memory and the
stream; every call passes raw device pointers to hand-written sm_100a kernels.
There is no CPU fallback: a missing library or a missing CUDA device raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbarefoot_sm100.so")

KERN = {"SE": 0, "M32": 1, "M52": 2}
ACQ_EI, ACQ_PI, ACQ_UCB, ACQ_GREEDY = 1, 2, 4, 8
SLOT_EI, SLOT_PI, SLOT_UCB, SLOT_GREEDY, SLOT_TOP2 = 0, 1, 2, 3, 4
NSLOT_VAL, NSLOT_IDX = 5, 4
MAX_DIM, MAX_MODELS = 16, 8

c_p, c_i, c_ll, c_d, c_sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_double, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/barefoot_sm100.h declares
SIGNATURES = {
    "bf_version": (c_i, []),
    "bf_last_error": (ctypes.c_char_p, []),
    "bf_padded_n": (c_i, [c_i]),
    "bf_linv_packed_doubles": (c_sz, [c_i]),
    "bf_posterior_tile": (c_i, [c_i]),
    "bf_posterior_parts": (c_i, [c_i, c_ll, c_i]),
    "bf_kernel_matrix": (c_i, [c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_ll, c_p, c_p]),
    "bf_kernel_matrix_lower": (c_i, [c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_ll, c_p, c_p]),
    "bf_cholesky_batched": (c_i, [c_i, c_i, c_p, c_p, c_p]),
    "bf_factor_finish": (c_i, [c_i, c_i, c_p, c_p, c_ll, c_p, c_p, c_p, c_p]),
    "bf_factor_finish_kernels": (c_i, [c_i, c_i, c_i]),
    "bf_cholesky_dinv_supported": (c_i, [c_i]),
    "bf_cholesky_batched_dinv": (c_i, [c_i, c_i, c_p, c_p, c_p, c_p]),
    "bf_posterior_solve_cols": (c_i, [c_i]),
    "bf_posterior_solve_parts": (c_i, [c_i, c_ll]),
    "bf_gp_posterior_solve": (c_i, [c_i, c_ll, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_p, c_p,
                                    c_p, c_p, c_i, c_i, c_d, c_d, c_d, c_p, c_p, c_p]),
    "bf_prescale_train": (c_i, [c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p]),
    "bf_gp_posterior": (c_i, [c_i, c_ll, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                              c_p, c_p, c_i, c_i, c_d, c_d, c_d, c_p, c_p, c_p, c_p]),
    "bf_gp_solve_rows": (c_i, [c_i, c_ll, c_i, c_i, c_p, c_p, c_p, c_d, c_p, c_p, c_p, c_p, c_p, c_p]),
    "bf_gp_cov": (c_i, [c_i, c_ll, c_ll, c_i, c_i, c_p, c_p, c_p, c_d, c_p, c_p, c_p, c_ll, c_p]),
    "bf_acq_finalize": (c_i, [c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    "bf_kg_tiles": (c_i, [c_ll]),
    "bf_top2": (c_i, [c_ll, c_i, c_p, c_p, c_p, c_p]),
    "bf_kg_diag": (c_i, [c_ll, c_ll, c_i, c_p, c_p, c_p, c_p, c_d, c_p, c_p, c_p, c_p, c_p, c_p]),
    "bf_kg_full_max_n": (c_i, []),
    "bf_kg_full_workspace": (c_sz, [c_i, c_ll, c_i]),
    "bf_kg_full": (c_i, [c_i, c_ll, c_i, c_p, c_p, c_d, c_p, c_p, c_p, c_p, c_p]),
    "bf_acq_tiles": (c_i, [c_ll]),
    "bf_acq_eval": (c_i, [c_i, c_ll, c_i, c_p, c_p, c_d, c_d, c_d, c_p, c_p, c_p, c_p, c_p, c_p]),
    "bf_records_pack": (c_i, [c_i, c_p, c_p, c_ll, c_p, c_p]),
    "bf_records_reduce": (c_i, [c_i, c_i, c_p, c_p, c_p, c_p]),
    "bf_reification": (c_i, [c_i, c_ll, c_p, c_p, c_p, c_p, c_p]),
    "bf_kmedoids_rowsum": (c_i, [c_ll, c_i, c_p, c_ll, c_ll, c_p, c_p]),
    "bf_kmedoids_assign": (c_i, [c_ll, c_i, c_i, c_p, c_p, c_p, c_p]),
    "bf_kmedoids_cluster_cost": (c_i, [c_ll, c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    "bf_kmedoids_cluster_cost_part": (c_i, [c_ll, c_i, c_i, c_p, c_p, c_p, c_i, c_i, c_p, c_p]),
    "bf_kmedoids_sort_workspace": (c_sz, [c_ll, c_i]),
    "bf_kmedoids_sort_labels": (c_i, [c_ll, c_i, c_p, c_p, c_p, c_p, c_p]),
    "bf_kmedoids_update": (c_i, [c_ll, c_i, c_p, c_p, c_p, c_p, c_p, c_p]),
    "bf_kmedoids_update_partial": (c_i, [c_ll, c_i, c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_p]),
    "bf_kmedoids_update_final": (c_i, [c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    "bf_kmedoids_iteration": (c_i, [c_ll, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p]),
    "bf_cholesky_fused_supported": (c_i, [c_i, c_i]),
    "bf_cholesky_ll_inplace": (c_i, [c_i, c_i, c_p, c_p, c_p, c_i, c_p]),
    "bf_cholesky_fused": (c_i, [c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_ll, c_p, c_p, c_p, c_i, c_p]),
    "bf_zscore_targets_batched": (c_i, [c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p]),
    "bf_fantasy_rows": (c_i, [c_i, c_i, c_p, c_p, c_p, c_ll, c_p, c_p, c_p, c_d, c_d, c_d, c_d, c_p, c_p, c_p, c_p]),
    "bf_probe_dmma_flop": (c_d, [c_i, c_i]),
    "bf_probe_dmma": (c_i, [c_i, c_i, c_p, c_p]),
    "bf_kmedoids_matrix_assign": (c_i, [c_ll, c_i, c_p, c_ll, c_p, c_p, c_p]),
    "bf_kmedoids_matrix_update": (c_i, [c_ll, c_i, c_p, c_ll, c_p, c_p, c_p, c_p, c_p, c_p]),
    "bf_group_best": (c_i, [c_ll, c_ll, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "bf_affine": (c_i, [c_ll, c_p, c_d, c_d, c_p, c_p]),
    "bf_total_variance": (c_i, [c_ll, c_p, c_d, c_d, c_p, c_d, c_p, c_p]),
    "bf_zscore_targets": (c_i, [c_i, c_p, c_p, c_p, c_p, c_p, c_p]),
}

_lib = None
launches = 0          # kernels launched through this binding (bench.py reports it)


class BarefootCudaError(RuntimeError):
    pass


def load():
    """Load the shared library (no CUDA call is made here)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise BarefootCudaError(
                "libbarefoot_sm100.so is not built (%s); run `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C barefoot_framework_b200/csrc`" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise BarefootCudaError("barefoot_framework_b200 needs a CUDA device (sm_100a); there is no "
                                "CPU fallback")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device-resident contiguous tensor required"
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


# kernels per exported call, for the launch counter
_KERNELS_PER_CALL = {"bf_kmedoids_matrix_update": 2, "bf_kg_diag": 2, "bf_acq_eval": 2, "bf_gp_solve_rows": 3, "bf_kg_full": 2}


def call(name, *args):
    """Invoke an int-returning export and raise on a non-zero status."""
    global launches
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise BarefootCudaError("%s failed (%d): %s" % (name, rc, lib.bf_last_error().decode()))
    if name == "bf_factor_finish":
        launches += lib.bf_factor_finish_kernels(args[0], args[1], 1 if args[7] else 0)
    elif name == "bf_kmedoids_sort_labels":
        launches += 3
    elif name == "bf_kmedoids_iteration":
        launches += 6
    elif name == "bf_group_best":
        launches += 4 if args[0] else 2
    else:
        launches += _KERNELS_PER_CALL.get(name, 1)
    return rc
