"""ctypes binding of the C-ABI library (include/pdfnet_b200.h).

There is no CPU or eager-PyTorch fallback: if the shared library is missing the
import of any compute module fails loudly, and every op raises when handed a
non-CUDA tensor.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpdfnet_b200.so")

_i64, _i32, _f32, _vp = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p

# name -> argtypes, exactly mirroring include/pdfnet_b200.h
SIGNATURES = {
    "pdf_knn_ball": [_vp, _i64, _i32, _i32, _i32, _f32, _i64, _i64, _i64, _vp, _vp],
    "pdf_fps": [_vp, _i64, _i32, _i32, _vp, _i64, _i64, _i64, _vp, _vp],
    "pdf_gather_nchw": [_vp, _i64, _i32, _i32, _i64, _vp, _i32, _i64, _vp, _vp],
    "pdf_pyramid_gather": [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp,
                           _vp, _vp],
    "pdf_gather_nhwc": [_vp, _i64, _i32, _i32, _i64, _vp, _i32, _i64, _vp, _vp],
    "pdf_pyramid_gather_nhwc": [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _i32, _vp, _vp,
                                _vp, _vp, _vp],
    "pdf_group_gather": [_vp, _i64, _i32, _i32, _i32, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _vp],
    "pdf_linear_f32": [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _i64, _vp],
    "pdf_sa_mlp_max_bf16": [_vp, _i64, _i32, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _i64, _i32,
                            _vp],
    "pdf_sa_pack_weights_host": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "pdf_pack_image_host": [_vp, _i64, _i32, _i64, _i32, _vp],
    "pdf_rows_to_image": [_vp, _i64, _i64, _i32, _i32, _vp, _i32, _i32, _i32, _vp],
    "pdf_gemm_bf16": [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _i64, _i64, _vp, _i64,
                      _vp, _i32, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _i64, _vp],
    "pdf_sft_xyz_f32": [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "pdf_center_im2col": [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp],
    "pdf_backproject": [_vp, _vp, _i64, _i32, _i32, _vp, _vp],
    "pdf_depth2pcl": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp],
    "pdf_depth2pcl_seeded": [_vp, _vp, _i32, _vp, _vp, _vp, _vp, ctypes.c_uint32, _i64, _i32, _i32, _i32, _i32, _vp, _vp,
                             _vp, _vp],
    "pdf_depth2pcl_host_randomness": [ctypes.c_uint32, _i64, _i64, _vp, _vp],
    "pdf_pyramid_gather_bf16": [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _i32, _vp, _vp,
                                _vp, _vp, _vp],
    "pdf_mano_lbs": [_vp] * 11 + [_i64, _vp, _i32, _i32, _vp, _vp, _vp, _vp],
    "pdf_mano_lbs_rootmat": [_vp] * 11 + [_i64, _vp, _i32, _i32, _vp, _vp, _vp, _vp],
    "pdf_rodrigues": [_vp, _i64, _vp, _vp],
    "pdf_joint_regress": [_vp, _i32, _vp, _i64, _vp, _vp],
    "pdf_mano_pose_feature": [_vp, _vp, _i64, _vp, _vp],
    "pdf_split_coeff": [_vp, _i64, _i32, _i32, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp],
    "pdf_rows_to_image_t": [_vp, _i64, _i64, _i32, _i32, _vp, _i64, _i32, _vp],
    "pdf_gemm_bf16_batched": [_vp, _i32, _i32, _i64, _vp, _i32, _i32, _i64, _i32, _i32, _vp, _i64, _i64, _i64, _vp, _vp],
    "pdf_gemm_tn_bf16": [_vp, _i32, _vp, _i32, _i64, _i32, _i32, _vp, _i64, _i64, _vp],
    "pdf_bn_stats": [_vp, _i64, _i64, _i32, _vp, _vp],
    "pdf_bn_finalize": [_vp, _vp, _i64, _i32, _f32, _f32, _vp, _vp, _vp, _vp, _vp],
    "pdf_bn_act_fwd": [_vp, _i64, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _vp, _i64, _vp, _vp],
    "pdf_bn_act_bwd": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _vp, _vp, _i64, _vp, _vp],
    "pdf_col_sum": [_vp, _i64, _i64, _i32, _vp, _vp],
    "pdf_act_bwd": [_vp, _i64, _vp, _i64, _i32, _i64, _i32, _vp, _i64, _vp],
    "pdf_sft_modulate": [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _vp, _i64, _vp],
    "pdf_sft_modulate_bwd": [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _vp, _i64, _vp, _i64, _vp],
    "pdf_linear_tn_f32": [_vp, _i64, _vp, _i64, _i64, _i32, _i32, _vp, _i64, _vp],
    "pdf_linear_smallk_f32": [_i32, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _vp, _i64, _vp],
    "pdf_group_max": [_vp, _i64, _i32, _i64, _i32, _vp, _i64, _vp, _vp],
    "pdf_bn_maxpool_bwd": [_vp, _i64, _vp, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _vp, _vp, _vp],
    "pdf_group_max_bwd": [_vp, _i64, _vp, _i64, _i32, _i64, _i32, _vp, _i64, _vp],
    "pdf_group_scatter_add": [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _i64, _vp],
    "pdf_gather_nchw_bwd": [_vp, _vp, _i64, _i32, _i64, _i32, _vp, _vp],
    "pdf_row_combine": [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i64, _vp, _vp, _f32, _i32, _vp, _i64, _vp,
                        _i64, _vp, _vp, _vp],
    "pdf_graph_cheby_ln": [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _vp, _vp, _f32, _i32,
                           _vp, _i64, _vp, _vp],
    "pdf_gemm_bf16_grouped": [_vp, _i32, _i32, _vp, _i32, _i32, _i64, _i32, _vp, _i64, _i32, _vp, _i64, _i64, _vp, _i32,
                              _vp, _vp],
    "pdf_row_combine_grouped": [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _i64, _i64, _i64, _vp, _vp, _f32,
                                _i32, _vp, _i64, _vp, _i64, _vp, _vp, _vp],
    "pdf_graph_cheby_ln_grouped": [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _i64, _vp, _vp,
                                   _f32, _i32, _vp, _i64, _vp, _vp],
    "pdf_mha": [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _i64, _vp],
    "pdf_decoder_heads": [_vp, _i64, _i64, _i32, _i32] + [_vp] * 12,
    "pdf_mha_tc": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _vp],
    "pdf_decoder_project": [_vp, _i32, _vp, _i32, _vp, _i64, _f32, _vp, _i32, _i64, _vp, _vp, _vp, _vp, _vp],
    "pdf_mano_lbs_pair": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp],
    "pdf_host_device_pointer": [_vp, _vp],
}
EXPORTS = sorted(list(SIGNATURES) + ["pdf_version", "pdf_last_error", "pdf_launch_count", "pdf_sa_pack_size",
                                     "pdf_image_bytes"])

ACT_NONE, ACT_RELU, ACT_LEAKY01 = 0, 1, 2
GEMM_OUT_SPLIT = 256          # OR-ed onto act: pdf_gemm_bf16 writes out_img as a split image [hi | hi | lo]
GEMM_LIGHT = 512              # OR-ed onto act: half-footprint GEMM configuration (two CTAs per SM)
BN_PLAIN_IMAGE = 2            # OR-ed onto relu (pdf_bn_act_fwd / _bwd / pdf_bn_maxpool_bwd): plain bf16 image output
EPI_STORE, EPI_SFT_SCALE, EPI_ACCUM, EPI_GROUP_MAX = 0, 1, 2, 3

_lib = None


def load():
    """Load (once) and return the ctypes handle; raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "pdfnet_b200: %s not found. Build it with `python -m pdfnet_b200.build` "
            "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    lib.pdf_version.restype = ctypes.c_int
    lib.pdf_last_error.restype = ctypes.c_char_p
    lib.pdf_launch_count.restype = ctypes.c_int64
    lib.pdf_sa_pack_size.argtypes = [_i32, _i32, _i32, _i32]
    lib.pdf_sa_pack_size.restype = ctypes.c_int64
    lib.pdf_image_bytes.argtypes = [_i64, _i32]
    lib.pdf_image_bytes.restype = ctypes.c_int64
    _lib = lib
    return lib


def launch_count():
    return int(load().pdf_launch_count())


_tls = threading.local()      # device index of the CUDA tensors whose pointers ptr() handed out for the pending call


def call(name, *args):
    """Run one C-ABI entry point.  The kernel is enqueued on the device that owns the tensors passed through
    ptr() (not on whatever device happens to be current) and on that device's current stream (see stream())."""
    lib = load()
    dev = getattr(_tls, "dev", None)
    _tls.dev = None
    if dev is not None and dev != torch.cuda.current_device():
        with torch.cuda.device(dev):
            rc = getattr(lib, name)(*args)
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.pdf_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (%d): %s" % (name, rc, msg))


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  All CUDA tensors of one call must live on one device."""
    if t is None:
        return None
    if t.is_cuda:
        dev = getattr(_tls, "dev", None)
        if dev is None:
            _tls.dev = t.device.index
        elif dev != t.device.index:
            _tls.dev = None
            raise RuntimeError("pdfnet_b200: tensors of one call live on different devices (cuda:%d and %s)"
                               % (dev, t.device))
    return ctypes.c_void_p(t.data_ptr())


def host_ptr(t):
    """Device alias of a page-locked (pinned) HOST tensor, for the kernels that read a host-resident input in
    place (zero-copy; pdf_host_device_pointer).  Raises when the tensor is not pinned, mapped host memory."""
    if t.is_cuda:
        return ptr(t)
    out = ctypes.c_void_p()
    rc = load().pdf_host_device_pointer(ctypes.c_void_p(t.data_ptr()), ctypes.byref(out))
    if rc != 0 or not out.value:
        raise RuntimeError("pdfnet_b200: a host tensor handed to a kernel must be page-locked (pin_memory() / "
                           "parallel.pinned_like): %s" % load().pdf_last_error().decode("utf-8", "replace"))
    return out


def stream():
    """Current stream of the device of the tensors already passed through ptr() (the stream argument is the
    last one of every entry point), else of the current device."""
    dev = getattr(_tls, "dev", None)
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("pdfnet_b200 ops run on CUDA tensors only (no CPU fallback); got device %s" % t.device)


def f32c(t):
    """Contiguous fp32 view/copy."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
