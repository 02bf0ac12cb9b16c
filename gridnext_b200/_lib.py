"""ctypes binding of the C-ABI shared library (include/gridnext_b200.h).

PyTorch only provides device memory and the CUDA stream; every kernel on the hot path is a plain
``extern "C"`` entry point taking raw device pointers, sizes and a ``cudaStream_t``.  There is NO
fallback: if the library is missing or a call fails, the caller gets an exception.
"""
import ctypes
import os
import torch
import torch.distributed as _dist

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libgridnext_b200.so')

_lib = None

vp, ci, cl, cf, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_float, ctypes.c_double

# name -> argtypes (all functions return int except where noted); mirrors include/gridnext_b200.h
_SIGS = {
    'gn_version': [],
    'gn_device_sm_count': [],
    'gn_set_pdl': [ci],
    'gn_hexconv_n_taps': [ci],
    'gn_hexconv_pack': [vp, vp, vp, vp, ci, ci, ci, ci, vp, vp],
    'gn_hexconv_unpack_grad': [vp, vp, vp, vp, vp, ci, ci, ci, vp],
    'gn_hexconv_fwd': [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp],
    'gn_hexconv_fwd_tc': [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp, vp],
    'gn_hexconv_tc_supported': [ci, ci, ci, ci, ci],
    'gn_hexconv_tc2_supported': [ci, ci, ci, ci, ci],
    'gn_hexconv_tc2_set_trace': [vp],
    'gn_hexconv_fwd_tc2': [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp, vp],
    'gn_hexconv_wgrad_tc': [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp, vp],
    'gn_hexconv_wgrad_tc2': [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp],
    'gn_hexconv_wgrad': [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp],
    'gn_cell_inverse': [vp, ci, vp, ci, vp],
    'gn_grid_gather_rows': [vp, cl, vp, vp, cl, cl, cl, vp],
    'gn_grid_gather_cols': [vp, cl, vp, vp, cl, ci, ci, vp],
    'gn_grid_labels': [vp, vp, vp, ci, vp],
    'gn_mm_fg_consistency': [vp, cl, vp, cl, ci, vp, vp, ci, vp],
    'gn_sqconv_pack': [vp, ci, ci, ci, ci, vp, vp],
    'gn_sqconv_unpack_grad': [vp, vp, ci, ci, ci, vp],
    'gn_sqconv_fwd': [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp],
    'gn_sqconv_wgrad': [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp],
    'gn_bn_finalize': [vp, vp, vp, vp, vp, cf, cf, cd, vp, vp, vp, ci, ci, vp],
    'gn_bn_eval_affine': [vp, vp, vp, vp, cf, vp, vp, vp, ci, vp],
    'gn_bn_stats': [vp, vp, ci, ci, cl, vp],
    'gn_bn_act_fwd': [vp, vp, vp, vp, ci, ci, cl, ci, vp],
    'gn_bn_act_bwd': [vp, vp, vp, vp, vp, vp, cd, ci, vp, vp, vp, ci, ci, cl, ci, vp],
    'gn_bn_act_bwd_reduce': [vp, vp, vp, vp, vp, vp, ci, ci, cl, ci, vp],
    'gn_bn_act_bwd_apply': [vp, vp, vp, vp, vp, vp, cd, ci, vp, vp, vp, ci, ci, cl, ci, vp],
    'gn_corrector_fused_supported': [ci, vp, vp],
    'gn_corrector_fused_fwd': [vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp, vp, vp, vp],
    'gn_corrector_fused_bwd': [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp],
    'gn_fg_predictions': [vp, vp, vp, vp, vp, vp, ci, ci, cl, vp],
    'gn_masked_ce': [vp, vp, vp, vp, vp, vp, cf, ci, ci, cl, vp],
    'gn_spot_table': [vp, vp, vp, vp, vp, ci, ci, ci, vp, vp, vp],
    'gn_patch_gather': [vp, cl, ci, ci, vp, ci, ci, vp, vp, vp, ci, vp],
    'gn_patch_gather_resize': [vp, cl, ci, ci, vp, ci, ci, ci, vp, vp, ci, ci, vp, vp, vp, ci, vp],
    'gn_normalize_u8': [vp, vp, cl, ci, vp, vp, vp, ci, vp],
    'gn_cast_f32_bf16': [vp, vp, cl, vp],
    'gn_rows_affine_bf16': [vp, cl, vp, vp, ci, vp, cl, cl, ci, ci, vp],
    'gn_colstats_bf16': [vp, cl, cl, ci, vp, vp, vp],
    'gn_bn_train_coeffs': [vp, vp, cl, vp, vp, cf, cf, vp, vp, vp, vp, vp, vp, ci, vp],
    'gn_affine_relu_bf16': [vp, cl, vp, cl, cl, ci, vp, vp, ci, vp],
    'gn_bn_train_fix_coeffs': [vp, vp, vp, vp, vp, cl, ci, vp, vp, ci, vp],
    'gn_bn_train_fix_bf16': [vp, cl, vp, cl, cl, ci, vp, vp, vp],
    'gn_gemm_tn_bf16': [vp, cl, vp, cl, ci, ci, ci, vp, cl, vp, vp, vp],
    'gn_conv1x1_bwd_bf16': [vp, cl, vp, cl, ci, ci, vp, cl, vp, cl, vp, vp, vp, vp, vp, ci, ci, vp, cl, vp],
    'gn_prep_job_bytes': [],
    'gn_prepare_weights': [vp, ci, cl, vp, vp],
    'gn_unpack_gradients': [vp, ci, cl, vp, vp],
    'gn_stem_pack_input': [vp, ci, ci, ci, vp, vp],
    'gn_stem_pack_weight': [vp, ci, vp, vp],
    'gn_stem_conv_fwd': [vp, ci, ci, vp, ci, vp, vp, ci, vp, cl, vp],
    'gn_stem_conv_wgrad': [vp, ci, ci, vp, cl, ci, vp, vp],
    'gn_stem_unpack_wgrad': [vp, ci, vp, vp],
    'gn_maxpool3s2_fwd': [vp, cl, ci, ci, ci, ci, vp, cl, vp, vp],
    'gn_maxpool3s2_bnrelu_bwd': [vp, cl, vp, vp, cl, ci, ci, ci, ci, vp, vp, vp, vp, cl, vp, ci, vp],
    'gn_bnrelu_avgpool2_fwd': [vp, cl, ci, ci, ci, ci, vp, vp, vp, cl, vp],
    'gn_pool_bnrelu_bwd': [vp, cl, ci, vp, cl, ci, ci, ci, ci, vp, vp, vp, vp, vp, cl, vp, ci, vp],
    'gn_bnrelu_gap_fwd': [vp, cl, ci, ci, ci, vp, vp, vp, cl, vp],
    'gn_linear_small_fwd': [vp, cl, vp, vp, ci, ci, ci, vp, vp],
    'gn_linear_small_bwd': [vp, vp, cl, vp, ci, ci, ci, vp, cl, vp, vp, vp],
    'gn_bn_eval_consts': [vp, vp, vp, vp, cf, ci, vp, vp, vp, vp, vp],
    'gn_conv3x3_pack': [vp, ci, ci, ci, vp, ci, vp],
    'gn_conv3x3_bf16': [vp, cl, ci, ci, ci, ci, vp, ci, ci, vp, cl, vp, cl, ci, vp, vp, vp, vp, vp, ci, vp],
    'gn_gemm_bf16': [vp, cl, vp, cl, ci, ci, ci, vp, cl, ci, ci, vp, vp, ci, vp, vp, vp, cl, ci, vp, vp, vp, vp, vp, ci, ci, vp],
    'gn_conv3x3_wgrad_bf16': [vp, cl, vp, cl, ci, ci, ci, ci, ci, vp, vp],
    'gn_conv3x3_unpack_grad': [vp, ci, ci, vp, vp],
}


def declared_symbols():
    return sorted(_SIGS) + ['gn_last_error', 'gn_hexconv_tc_workspace_bytes', 'gn_hexconv_tc_wgrad_workspace_bytes', 'gn_hexconv_tc2_workspace_bytes']


def load():
    """Load the library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "gridnext_b200: CUDA library %s is missing. Build it with `python -m gridnext_b200.build` "
            "(there is no CPU fallback)." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.gn_last_error.restype = ctypes.c_char_p
    lib.gn_last_error.argtypes = []
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = args
    for nm in ('gn_hexconv_tc_workspace_bytes', 'gn_hexconv_tc_wgrad_workspace_bytes'):
        getattr(lib, nm).restype = ctypes.c_long
        getattr(lib, nm).argtypes = [ci, ci, ci]
    lib.gn_hexconv_tc2_workspace_bytes.restype = ctypes.c_long
    lib.gn_hexconv_tc2_workspace_bytes.argtypes = []
    _lib = lib
    return lib


def register(sigs):
    """Used by sibling modules to add entry points (keeps one table for the symbol test)."""
    _SIGS.update(sigs)
    if _lib is not None:
        for name, args in sigs.items():
            fn = getattr(_lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = args


def check(rc, what=''):
    if rc == 0:
        return
    msg = load().gn_last_error().decode(errors='replace')
    if rc < 0:
        raise ValueError('gridnext_b200 %s: %s (code %d)' % (what, msg, rc))
    raise RuntimeError('gridnext_b200 %s: CUDA error %d: %s' % (what, rc, msg))


# kernels launched per C-ABI call (for the benchmark's gpu_launches count); default 1
KERNELS_PER_CALL = {'gn_corrector_fused_supported': 0, 'gn_cell_inverse': 2, 'gn_mm_fg_consistency': 2, 'gn_hexconv_fwd_tc': 3, 'gn_hexconv_wgrad_tc': 3, 'gn_hexconv_tc_supported': 0, 'gn_hexconv_tc2_supported': 0, 'gn_hexconv_tc2_set_trace': 0, 'gn_hexconv_fwd_tc2': 2, 'gn_masked_ce': 3, 'gn_bn_act_bwd': 2, 'gn_spot_table': 1, 'gn_linear_small_bwd': 2, 'gn_version': 0, 'gn_prep_job_bytes': 0,
                    'gn_device_sm_count': 0, 'gn_hexconv_n_taps': 0, 'gn_set_pdl': 0}
LAUNCHES = [0]
PROFILE = None      # when a dict: name -> [n_calls, [cuda event pairs]]


_PDL_DECIDED = [False]


def _decide_pdl(lib):
    """Programmatic dependent launch stays on for single-process runs and is switched off once a process group with more than one rank
    exists (the 2-GPU step with NCCL's all-reduce between kernels that trigger their dependents early hung on the B200 pool)."""
    if not _dist.is_available():
        _PDL_DECIDED[0] = True
    elif _dist.is_initialized():
        if _dist.get_world_size() > 1:
            lib.gn_set_pdl(0)
        _PDL_DECIDED[0] = True


def call(name, *args):
    lib = load()
    if not _PDL_DECIDED[0]:
        _decide_pdl(lib)
    LAUNCHES[0] += KERNELS_PER_CALL.get(name, 1)
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(getattr(lib, name)(*args), name)
        e1.record()
        PROFILE.setdefault(name, []).append((e0, e1, args))
        return
    check(getattr(lib, name)(*args), name)


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("gridnext_b200 kernels need CUDA tensors (got a %s tensor); there is no CPU path" % t.device)
