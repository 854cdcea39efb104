"""ctypes binding of libmurcl_b200.so (C ABI in include/murcl_b200.h).

There is no CPU fallback: if the shared object is missing or a call fails, a ``MurclError`` is
raised.  Use ``python -m murcl_b200.build`` (or ``__graft_entry__.build()``) to compile it.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libmurcl_b200.so"

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_TANH, ACT_SIGMOID, ACT_TANH_SIGMOID = 0, 1, 2, 3, 4
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05 = 0, 1, 2


class MurclError(RuntimeError):
    pass


_p, _i, _l, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

# name -> (restype, argtypes); must list every symbol declared in include/murcl_b200.h
SIGNATURES = {
    "murcl_version": (_i, []),
    "murcl_last_error": (C.c_char_p, []),
    "murcl_device_info": (_i, [C.POINTER(_i)] * 3),
    "murcl_launch_count": (_l, []),
    "murcl_set_row_order": (_i, [_i]),
    "murcl_csr_rank_patches": (_i, [_p, _p, _i, _i, _p, _p, _p]),
    "murcl_pack_select": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "murcl_pack_gather": (_i, [_p, _i, _i, _p, _i, _i, _p, _p, _p, _i, _p]),
    "murcl_perm_cycle_order": (_i, [_p, _i, _i, _p, _p]),
    "murcl_pack_gather_ordered": (_i, [_p, _i, _i, _p, _i, _i, _p, _p, _p, _p, _i, _p]),
    "murcl_linear_fwd": (_i, [_p, _p, _p, _p, _l, _i, _i, _i, _i, _i, _i, _p, _p]),
    "murcl_linear_bwd_input": (_i, [_p, _p, _p, _l, _i, _i, _p, _p, _p, _p, _p, _f, _p, _i, _i, _p]),
    "murcl_linear_bwd_input_accum_supported": (_i, [_l, _i, _i, _i]),
    "murcl_linear_bwd_input_accum": (_i, [_p, _p, _p, _l, _i, _i, _i, _p]),
    "murcl_linear_bwd_weight_workspace": (_l, [_l, _i, _i]),
    "murcl_linear_bwd_weight": (_i, [_p, _p, _p, _p, _l, _i, _i, _i, _i, _p, _i, _p]),
    "murcl_split_planes": (_i, [_p, _l, _i, _i, _l, _p, _p]),
    "murcl_linear_split_supported": (_i, [_l, _i, _i]),
    "murcl_linear_fwd_split": (_i, [_p, _p, _p, _p, _l, _i, _i, _i, _i, _l, _l, _p, _p]),
    "murcl_linear_bwd_input_split": (_i, [_p, _p, _p, _l, _i, _i, _p, _p, _p, _p, _f, _p, _i, _l, _l, _p]),
    "murcl_linear_bwd_weight_split_workspace": (_l, [_l, _i, _i]),
    "murcl_linear_bwd_weight_split": (_i, [_p, _p, _p, _l, _i, _i, _i, _l, _p, _i, _p]),
    "murcl_attn_score_fwd": (_i, [_p, _p, _p, _p, _l, _i, _i, _i, _p]),
    "murcl_seg_softmax": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "murcl_attnpool_supported": (_i, [_i, _i, _i, _i]),
    "murcl_attnpool_workspace": (_l, [_l, _i, _i]),
    "murcl_attnpool_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _l, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "murcl_attnpool_bwd_supported": (_i, [_i, _i, _i, _i]),
    "murcl_attnpool_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _l, _i, _i, _i, _i, _i, _f, _i, _p, _p, _p, _p, _p]),
    "murcl_seg_wsum_workspace": (_l, [_l, _i, _i, _i]),
    "murcl_seg_wsum": (_i, [_p, _p, _p, _l, _i, _i, _i, _i, _p, _p, _p]),
    "murcl_pool_bwd_scores": (_i, [_p, _p, _p, _p, _p, _p, _l, _i, _i, _i, _i, _i, _p, _p, _p]),
    "murcl_pool_bwd_direct": (_i, [_p, _p, _p, _l, _i, _i, _i, _p, _i, _p]),
    "murcl_attn_score_bwd": (_i, [_p, _p, _p, _p, _p, _p, _l, _i, _i, _f, _i, _p]),
    "murcl_seg_topk_ends": (_i, [_p, _p, _i, _i, _p, _p, _p]),
    "murcl_seg_argmax": (_i, [_p, _p, _i, _i, _p, _p]),
    "murcl_dsmil_scores_fwd": (_i, [_p, _p, _p, _l, _i, _i, _p, _p]),
    "murcl_dsmil_scores_bwd": (_i, [_p, _p, _p, _p, _p, _l, _i, _i, _i, _p, _p]),
    "murcl_gather_rows": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "murcl_scatter_add_rows": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "murcl_clam_inst_ce_fwd": (_i, [_p, _p, _p, _p, _i, _p, _p, _i, _p, _p, _p, _p]),
    "murcl_clam_inst_ce_bwd": (_i, [_p, _p, _p, _p, _p, _i, _p, _i, _p, _p, _p, _p]),
    "murcl_ntxent_workspace": (_l, [_i, _i]),
    "murcl_ntxent_fwd_bwd": (_i, [_p, _i, _i, _f, _p, _p, _p, _p, _p]),
    "murcl_ntxent_fwd_bwd_slab": (_i, [_p, _i, _i, _f, _i, _i, _p, _p, _p, _p, _p]),
    "murcl_ntxent_slab_workspace": (_l, [_i, _i, _i]),
    "murcl_ntxent_lse_slab": (_i, [_p, _i, _i, _f, _i, _i, _p, _p, _p, _p, _p, _p]),
    "murcl_ntxent_grad_slab": (_i, [_p, _i, _i, _f, _i, _i, _p, _p, _p, _p, _p]),
    "murcl_gru_cell_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _p]),
    "murcl_gru_cell_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "murcl_gru_cell_fwd_tape": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "murcl_gru_cell_bwd_tape": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "murcl_actor_head": (_i, [_p, _p, _f, _p, _p, _p, _i, _i, _p]),
    "murcl_cast": (_i, [_p, _i, _p, _i, _l, _p]),
    "murcl_row_segments": (_i, [_p, _i, _p, _p]),
    "murcl_colsum": (_i, [_p, _l, _i, _i, _p, _p]),
    "murcl_relu_bwd": (_i, [_p, _p, _p, _l, _i, _p]),
    "murcl_dropout": (_i, [_p, _l, _f, _p, _i, _p]),
    "murcl_adam_step": (_i, [_p, _p, _p, _p, _p, _l, _d, _d, _d, _d, _d, _d, _p, _p, _p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared object and attach prototypes.  Raises MurclError when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise MurclError(f"{LIB_PATH} not found: build it with `python -m murcl_b200.build` "
                         "(there is no CPU fallback for the MIL hot path)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here means header/library drift
        fn.restype = res
        fn.argtypes = args
    if lib.murcl_version() != 1:
        raise MurclError(f"ABI version mismatch: library reports {lib.murcl_version()}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().murcl_last_error().decode(errors="replace")
        raise MurclError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().murcl_launch_count())
