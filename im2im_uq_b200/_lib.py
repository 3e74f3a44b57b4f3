"""ctypes binding of libim2im_uq.so (the C ABI in include/im2im_uq.h) + the in-tree nvcc build recipe.

This is the stub a maintainer of the reference would add.  No torch types cross the boundary:
tensors are passed as raw device pointers, sizes and element strides; the CUDA stream is torch's current stream.
"""
import ctypes
import glob
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
LIB_DIR = os.path.join(_PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libim2im_uq.so")
HEADER = os.path.join(_ROOT, "include", "im2im_uq.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]

IM2IM_RCPS_ZERO_OUTPUTS = 1
IM2IM_RCPS_FORCE_GENERIC = 2
IM2IM_HEAD_QUANTILES = 0      # (lower, pred, upper): quantiles, quantiles_l1, inn
IM2IM_HEAD_RESIDUAL = 1       # (pred, |residual|): residual_magnitude, residual_magnitude_l1
IM2IM_HEAD_GAUSSIAN = 2       # (mean, variance)
IM2IM_HEAD_SOFTMAX_SETS = 3   # (lower quantile, argmax, upper quantile) from im2im_softmax_sets
LOSS_QUANTILES, LOSS_QUANTILES_L1, LOSS_GAUSSIAN, LOSS_RESIDUAL, LOSS_RESIDUAL_L1, LOSS_INN = range(6)
THREE_PLANE_HEADS = (IM2IM_HEAD_QUANTILES, IM2IM_HEAD_SOFTMAX_SETS)
IM2IM_RCPS_MAX_LAMBDAS = 8192

_lib = None

# Bumped by every code path that writes model parameters / BatchNorm statistics through raw device pointers or a CUDA-graph
# replay (FusedAdam.step, GraphedTrainStep, the native training forward): torch's tensor._version does not see those
# writes, so caches of derived weights (UNetInferenceEngine's folded BatchNorm) include this counter in their stamp.
_weights_generation = 0


def weights_changed() -> None:
    global _weights_generation
    _weights_generation += 1


def weights_generation() -> int:
    return _weights_generation


class Im2ImError(RuntimeError):
    """A C-ABI call returned a negative code."""


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into lib/libim2im_uq.so (nvcc cross-compiles without a GPU)."""
    srcs = sources()
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


def _declare(lib):
    c = ctypes
    vp, i64, i32, u32, f32 = c.c_void_p, c.c_int64, c.c_int32, c.c_uint32, c.c_float
    lib.im2im_abi_version.restype = c.c_int
    lib.im2im_abi_version.argtypes = []
    lib.im2im_last_error.restype = c.c_char_p
    lib.im2im_last_error.argtypes = []
    lib.im2im_launch_count.restype = c.c_ulonglong
    lib.im2im_launch_count.argtypes = []
    lib.im2im_rcps_miss_counts.restype = c.c_int
    lib.im2im_rcps_miss_counts.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, vp, i32, i32, vp, vp, u32, vp]
    lib.im2im_rcps_loss_table.restype = c.c_int
    lib.im2im_rcps_loss_table.argtypes = [vp, i64, i32, i64, i32, vp, vp]
    lib.im2im_quantile_nested_sets.restype = c.c_int
    lib.im2im_quantile_nested_sets.argtypes = [vp, vp, vp, i64, i64, i64, i64, i64, f32, i32, vp, vp, vp]
    lib.im2im_nested_sets.restype = c.c_int
    lib.im2im_nested_sets.argtypes = [i32, vp, vp, vp, i64, i64, i64, i64, i64, f32, i32, vp, vp, vp]
    lib.im2im_softmax_sets.restype = c.c_int
    lib.im2im_softmax_sets.argtypes = [vp, i64, i32, i64, i64, i64, vp, vp]
    lib.im2im_rcps_miss_map.restype = c.c_int
    lib.im2im_rcps_miss_map.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, f32, i32, vp, u32, vp]


def _declare_more(lib):
    c = ctypes
    vp, i64 = c.c_void_p, c.c_int64
    lib.im2im_fraction_missed_counts.restype = c.c_int
    lib.im2im_fraction_missed_counts.argtypes = [vp, vp, vp, i64, i64, i64, i64, i64, vp, vp]
    i32 = c.c_int32
    f64 = c.c_double
    lib.im2im_rcps_loss_table_dev.restype = c.c_int
    lib.im2im_rcps_loss_table_dev.argtypes = [vp, i64, i32, i64, vp, vp, vp]
    lib.im2im_rcps_decide.restype = c.c_int
    lib.im2im_rcps_decide.argtypes = [vp, i32, f64, f64, f64, f64, f64, f64, vp, vp]
    lib.im2im_rcps_decide_p2p.restype = c.c_int
    lib.im2im_rcps_decide_p2p.argtypes = [vp, vp, vp, vp, i32, i32, i32, f64, f64, f64, f64, f64, f64, vp, vp, vp]
    lib.im2im_rcps_fused_workspace_bytes.restype = c.c_size_t
    lib.im2im_rcps_fused_workspace_bytes.argtypes = [i32]
    lib.im2im_rcps_calibrate_fused.restype = c.c_int
    lib.im2im_rcps_calibrate_fused.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, vp, i32, i32, vp, vp, vp,
                                               f64, f64, f64, f64, f64, f64, vp, c.c_size_t, vp, vp, i32, i32, f64, vp,
                                               vp, vp]
    lib.im2im_rcps_calibrate_fused_check.restype = c.c_int
    lib.im2im_rcps_calibrate_fused_check.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i32, i32]
    lib.im2im_host_wait_flag.restype = c.c_int
    lib.im2im_host_wait_flag.argtypes = [vp, i32, i64]
    lib.im2im_conv_igemm_bf16.restype = c.c_int
    lib.im2im_conv_igemm_bf16.argtypes = [vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.im2im_conv_igemm_bf16_stats.restype = c.c_int
    lib.im2im_conv_igemm_bf16_stats.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp,
                                                vp, vp, vp]
    lib.im2im_conv_igemm_tf32.restype = c.c_int
    lib.im2im_conv_igemm_tf32.argtypes = [vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.im2im_conv_first_nhwc_f32.restype = c.c_int
    lib.im2im_conv_first_nhwc_f32.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.im2im_maxpool2x2_nhwc_f32.restype = c.c_int
    lib.im2im_maxpool2x2_nhwc_f32.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    lib.im2im_upsample2x_bilinear_nhwc_f32.restype = c.c_int
    lib.im2im_upsample2x_bilinear_nhwc_f32.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.im2im_head_conv3x3_act_nhwc_f32.restype = c.c_int
    lib.im2im_head_conv3x3_act_nhwc_f32.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.im2im_conv_wgrad_bf16.restype = c.c_int
    lib.im2im_conv_wgrad_bf16.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.im2im_pack_conv_weights.restype = c.c_int
    lib.im2im_pack_conv_weights.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    lib.im2im_conv_first_bf16.restype = c.c_int
    lib.im2im_conv_first_bf16.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.im2im_maxpool2x2_bf16.restype = c.c_int
    lib.im2im_maxpool2x2_bf16.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    lib.im2im_upsample2x_bilinear_bf16.restype = c.c_int
    lib.im2im_upsample2x_bilinear_bf16.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.im2im_head_conv3x3_f32.restype = c.c_int
    lib.im2im_head_conv3x3_f32.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.im2im_head_conv3x3_tc_f32.restype = c.c_int
    lib.im2im_head_conv3x3_tc_f32.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.im2im_conv_igemm_bf16_pool.restype = c.c_int
    lib.im2im_conv_igemm_bf16_pool.argtypes = [vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.im2im_head_conv3x3_tc_hist.restype = c.c_int
    lib.im2im_head_conv3x3_tc_hist.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp]
    lib.im2im_head_conv3x3_tc_folded_f32.restype = c.c_int
    lib.im2im_head_conv3x3_tc_folded_f32.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.im2im_rcps_counts_from_hist.restype = c.c_int
    lib.im2im_rcps_counts_from_hist.argtypes = [vp, i64, i32, vp, vp, vp]
    lib.im2im_planar_to_nhwc64_bf16.restype = c.c_int
    lib.im2im_planar_to_nhwc64_bf16.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    lib.im2im_planar_to_nhwc64_first8_bf16.restype = c.c_int
    lib.im2im_planar_to_nhwc64_first8_bf16.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    lib.im2im_head_conv3x3_act_f32.restype = c.c_int
    lib.im2im_head_conv3x3_act_f32.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]


def _declare_train(lib):
    c = ctypes
    vp, i64, i32, f32 = c.c_void_p, c.c_int64, c.c_int32, c.c_float
    sigs = {
        "im2im_channel_stats_bf16": [vp, i64, i32, vp, vp],
        "im2im_bn_finalize": [vp, i64, vp, vp, vp, f32, f32, i32, vp, vp, vp, vp, vp, vp, vp],
        "im2im_bn_apply_relu_bf16": [vp, vp, vp, i64, i32, vp, vp],
        "im2im_bn_relu_bwd_bf16": [vp, vp, vp, vp, vp, vp, i64, i32, vp, vp, vp],
        "im2im_bn_relu_bwd_apply_bf16": [vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, vp, vp],
        "im2im_bn_apply_relu_pool_bf16": [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp],
        "im2im_bn_relu_pool_bwd_bf16": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp],
        "im2im_maxpool2x2_bwd_bf16": [vp, vp, i32, i32, i32, i32, i32, vp, vp],
        "im2im_upsample2x_bilinear_bwd_bf16": [vp, i32, i32, i32, i32, i32, i32, vp, vp],
        "im2im_quantile_loss_f32": [vp, vp, i64, i64, f32, f32, f32, f32, f32, vp, vp, vp],
        "im2im_adam_step_f32": [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, f32, vp],
        "im2im_adam_step_dev_f32": [vp, vp, vp, vp, i64, f32, f32, f32, f32, vp, f32, vp],
        "im2im_head_loss_f32": [i32, vp, vp, i64, i64, f32, f32, f32, f32, f32, f32, vp, vp, vp],
        "im2im_head_bwd": [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp],
        "im2im_conv_first_wgrad": [vp, vp, i32, i32, i32, i32, i32, vp, vp],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.restype = c.c_int
        fn.argtypes = args


EXPORTS = ["im2im_abi_version", "im2im_last_error", "im2im_launch_count", "im2im_rcps_miss_counts",
           "im2im_rcps_loss_table", "im2im_quantile_nested_sets", "im2im_rcps_miss_map",
           "im2im_fraction_missed_counts", "im2im_rcps_loss_table_dev", "im2im_rcps_decide",
           "im2im_conv_igemm_bf16", "im2im_conv_wgrad_bf16", "im2im_pack_conv_weights", "im2im_conv_first_bf16",
           "im2im_maxpool2x2_bf16", "im2im_upsample2x_bilinear_bf16", "im2im_head_conv3x3_f32",
           "im2im_channel_stats_bf16", "im2im_bn_finalize", "im2im_bn_apply_relu_bf16", "im2im_bn_relu_bwd_bf16",
           "im2im_maxpool2x2_bwd_bf16", "im2im_upsample2x_bilinear_bwd_bf16", "im2im_quantile_loss_f32",
           "im2im_adam_step_f32", "im2im_head_bwd", "im2im_conv_first_wgrad", "im2im_nested_sets",
           "im2im_softmax_sets", "im2im_head_conv3x3_act_f32", "im2im_head_loss_f32",
           "im2im_adam_step_dev_f32", "im2im_head_conv3x3_tc_f32", "im2im_head_conv3x3_tc_hist", "im2im_conv_igemm_bf16_pool", "im2im_head_conv3x3_tc_folded_f32", "im2im_rcps_counts_from_hist", "im2im_planar_to_nhwc64_bf16",
           "im2im_rcps_decide_p2p", "im2im_rcps_fused_workspace_bytes", "im2im_rcps_calibrate_fused",
           "im2im_host_wait_flag", "im2im_rcps_calibrate_fused_check", "im2im_conv_igemm_tf32",
           "im2im_conv_first_nhwc_f32", "im2im_maxpool2x2_nhwc_f32", "im2im_upsample2x_bilinear_nhwc_f32",
           "im2im_head_conv3x3_act_nhwc_f32", "im2im_conv_igemm_bf16_stats", "im2im_bn_relu_bwd_apply_bf16",
           "im2im_planar_to_nhwc64_first8_bf16", "im2im_bn_apply_relu_pool_bf16", "im2im_bn_relu_pool_bwd_bf16"]


def load():
    """Load the CUDA library; fails loudly if it was not built (there is no CPU path behind this package)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Im2ImError(f"{LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(im2im_uq_b200 has no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        _declare(lib)
        _declare_more(lib)
        _declare_train(lib)
        if lib.im2im_abi_version() != 1:
            raise Im2ImError("libim2im_uq.so ABI version mismatch - rebuild")
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().im2im_last_error().decode("utf-8", "replace")
        raise Im2ImError(f"{what} failed with code {rc}: {msg}")


def launch_count() -> int:
    return int(load().im2im_launch_count())
