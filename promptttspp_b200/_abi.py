"""ctypes binding of include/pttspp_b200.h (the C ABI of libpttspp_b200.so).

This is the only place where Python touches the native library.  There is no fallback:
if the shared object is missing or an entry point fails, a RuntimeError is raised.
"""
import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("PTTSPP_B200_LIB", _HERE / "lib" / "libpttspp_b200.so"))

ACT_NONE, ACT_RELU, ACT_GELU, ACT_SWISH, ACT_GATE, ACT_TANH = range(6)

c_f32p = C.POINTER(C.c_float)
c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)


class Conv1dDesc(C.Structure):
    """pttspp_conv1d_desc"""

    _fields_ = [
        ("in_", C.c_void_p), ("in_bs", C.c_int64), ("in_ld", C.c_int32), ("T_in", C.c_int32), ("Cin", C.c_int32),
        ("w", C.c_void_p), ("w_ld", C.c_int32), ("bias", C.c_void_p),
        ("K", C.c_int32), ("dil", C.c_int32), ("pad", C.c_int32), ("in_stride", C.c_int32),
        ("out", C.c_void_p), ("out_bs", C.c_int64), ("out_ld", C.c_int32), ("T_out", C.c_int32), ("Cout", C.c_int32),
        ("m_begin", C.c_int32), ("M", C.c_int32), ("out_mul", C.c_int32), ("out_off", C.c_int32),
        ("in_len", C.c_void_p), ("out_len", C.c_void_p), ("in_add", C.c_void_p),
        ("addend", C.c_void_p), ("addend_bs", C.c_int64), ("addend_ld", C.c_int32),
        ("act", C.c_int32), ("acc_scale", C.c_float),
        ("res", C.c_void_p), ("res_bs", C.c_int64), ("res_ld", C.c_int32), ("res_scale", C.c_float),
        ("alpha", C.c_float), ("beta", C.c_float), ("out_div", C.c_float),
        ("B", C.c_int32), ("impl", C.c_int32),
        ("in_hi", C.c_void_p), ("in_lo", C.c_void_p), ("w_hi", C.c_void_p), ("w_lo", C.c_void_p),
        ("w_scale_inv", C.c_float), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("out_plane_add", C.c_void_p),
        ("out_plane_bs", C.c_int64), ("out_plane_ld", C.c_int32),
        ("res_hi", C.c_void_p), ("res_lo", C.c_void_p), ("res_plane_sub", C.c_void_p), ("res_plane_bs", C.c_int64),
        ("res_plane_ld", C.c_int32),
    ]


class LayerNormDesc(C.Structure):
    """pttspp_layernorm_desc"""

    _fields_ = [
        ("in_", C.c_void_p), ("in2", C.c_void_p), ("row_add", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("out", C.c_void_p), ("bs", C.c_int64), ("ld", C.c_int32), ("B", C.c_int32), ("T", C.c_int32), ("C", C.c_int32),
        ("eps", C.c_float), ("in_scale", C.c_float), ("in_len", C.c_void_p), ("out_len", C.c_void_p),
    ]


class BigVGANConfig(C.Structure):
    """pttspp_bigvgan_config"""

    _fields_ = [
        ("in_channel", C.c_int32), ("upsample_initial_channel", C.c_int32), ("num_upsamples", C.c_int32),
        ("upsample_rates", C.c_int32 * 8), ("upsample_kernel_sizes", C.c_int32 * 8),
        ("num_kernels", C.c_int32), ("resblock_kernel_sizes", C.c_int32 * 8),
        ("num_dilations", C.c_int32), ("resblock_dilations", (C.c_int32 * 8) * 8),
    ]


class AcousticConfig(C.Structure):
    """pttspp_acoustic_config"""

    _fields_ = [
        ("num_vocab", C.c_int32), ("channels", C.c_int32), ("emb_do_scale", C.c_int32),
        ("enc_heads", C.c_int32), ("enc_linear_units", C.c_int32), ("enc_blocks", C.c_int32),
        ("enc_ff_kernel", C.c_int32), ("enc_cnn_kernel", C.c_int32), ("rel_pos_legacy", C.c_int32),
        ("dur_layers", C.c_int32), ("dur_kernel", C.c_int32), ("dur_gaussians", C.c_int32),
        ("pitch_layers", C.c_int32), ("pitch_kernel", C.c_int32),
        ("fp_layers", C.c_int32), ("fp_kernel", C.c_int32),
        ("prompt_in", C.c_int32), ("prompt_mid", C.c_int32), ("style_gaussians", C.c_int32),
        ("norm_style_emb", C.c_int32),
        ("mel_dim", C.c_int32), ("K_step", C.c_int32), ("diff_layers", C.c_int32), ("diff_channels", C.c_int32),
        ("diff_kernel", C.c_int32), ("diff_dilation_cycle", C.c_int32),
        ("diff_scale", C.c_float), ("norm_scale", C.c_float), ("a_min", C.c_float), ("a_max", C.c_float),
    ]


class DiffNetLayer(C.Structure):
    """pttspp_diffnet_layer"""

    _fields_ = [
        ("wd_hi", C.c_void_p), ("wd_lo", C.c_void_p), ("wo_hi", C.c_void_p), ("wo_lo", C.c_void_p),
        ("bias_d", C.c_void_p), ("bias_o", C.c_void_p), ("scale_d", C.c_float), ("scale_o", C.c_float),
        ("dil", C.c_int32),
    ]


class DiffNetRunDesc(C.Structure):
    """pttspp_diffnet_run_desc"""

    _fields_ = [
        ("B", C.c_int32), ("T", C.c_int32), ("layer_begin", C.c_int32), ("layer_end", C.c_int32),
        ("cond", C.c_void_p), ("step_emb", C.c_void_p), ("y_hi", C.c_void_p * 2), ("y_lo", C.c_void_p * 2),
        ("skip", C.c_void_p), ("skip_hi", C.c_void_p), ("skip_lo", C.c_void_p), ("done", C.c_void_p),
        ("epoch", C.c_uint32), ("dbg_z", C.c_void_p), ("dbg_prof", C.c_void_p),
    ]


# name -> (restype, argtypes); every symbol declared in include/pttspp_b200.h
SIGNATURES = {
    "pttspp_last_error": (C.c_char_p, []),
    "pttspp_abi_version": (C.c_int, []),
    "pttspp_device_check": (C.c_int, []),
    "pttspp_launch_count": (C.c_int64, []),
    "pttspp_reset_launch_count": (None, []),
    "pttspp_debug_reload_env": (None, []),
    "pttspp_prof_enable": (None, [C.c_int]),
    "pttspp_prof_report": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "pttspp_conv1d_cl": (C.c_int, [C.POINTER(Conv1dDesc), C.c_void_p]),
    "pttspp_conv1d_dual_cl": (C.c_int, [C.POINTER(Conv1dDesc), C.POINTER(Conv1dDesc), C.c_void_p]),
    "pttspp_pack_conv_weight": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                          C.c_int, C.c_void_p]),
    "pttspp_pack_conv_weight_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                C.c_void_p, C.c_int, C.c_void_p]),
    "pttspp_split_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pttspp_umma_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "pttspp_pack_convtr_weight": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                            C.c_int, C.c_void_p]),
    "pttspp_layernorm_cl": (C.c_int, [C.POINTER(LayerNormDesc), C.c_void_p]),
    "pttspp_aa_snake_cl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "pttspp_aa_snake_pair_cl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "pttspp_aa_conv1d_cl": (C.c_int, [C.POINTER(Conv1dDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pttspp_duration_quantize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p]),
    "pttspp_length_regulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "pttspp_diffnet_create": (C.c_int, [C.POINTER(DiffNetLayer), C.c_int, C.POINTER(C.c_void_p)]),
    "pttspp_diffnet_destroy": (None, [C.c_void_p]),
    "pttspp_diffnet_flags_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "pttspp_diffnet_run": (C.c_int, [C.c_void_p, C.POINTER(DiffNetRunDesc), C.c_void_p]),
    "pttspp_relpos_attention": (C.c_int, [C.c_void_p] * 7 + [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pttspp_iir_filtfilt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p]),
    "pttspp_mel_spectrogram": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pttspp_mel_from_spec": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p,
                                       C.c_void_p]),
    "pttspp_iir_filtfilt_ragged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "pttspp_bert_embed": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_void_p]),
    "pttspp_mha_masked": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "pttspp_conv2d_bn_relu": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 8 + [C.c_void_p]),
    "pttspp_gru_last_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6),
    "pttspp_style_token_attention": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
                                     + [C.c_void_p] * 10),
    "pttspp_bigvgan_create": (C.c_int, [C.POINTER(BigVGANConfig), C.POINTER(C.c_void_p)]),
    "pttspp_bigvgan_destroy": (None, [C.c_void_p]),
    "pttspp_bigvgan_set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, c_i64p, C.c_int, C.c_void_p]),
    "pttspp_bigvgan_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pttspp_bigvgan_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "pttspp_bigvgan_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_size_t, C.c_void_p]),
    "pttspp_bigvgan_forward_f0": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_size_t, C.c_void_p]),
    "pttspp_nsf_source_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "pttspp_nsf_source": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float,
                                    C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_size_t, C.c_void_p]),
    "pttspp_acoustic_create": (C.c_int, [C.POINTER(AcousticConfig), C.POINTER(C.c_void_p)]),
    "pttspp_acoustic_destroy": (None, [C.c_void_p]),
    "pttspp_acoustic_set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, c_i64p, C.c_int, C.c_void_p]),
    "pttspp_acoustic_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pttspp_acoustic_encode_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "pttspp_acoustic_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pttspp_acoustic_encode_sampled": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pttspp_acoustic_encode_ref": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_size_t, C.c_void_p]),
    "pttspp_acoustic_decode_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "pttspp_acoustic_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pttspp_acoustic_decode_rng": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pttspp_philox_normal": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """Load the shared object (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m promptttspp_b200.build` "
                "(there is no CPU / PyTorch fallback for the hot path)"
            )
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if handle.pttspp_abi_version() != 1:
            raise RuntimeError("libpttspp_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(status):
    if status != 0:
        raise RuntimeError("pttspp_b200: " + lib().pttspp_last_error().decode(errors="replace"))


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    import torch

    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: tensor is on {t.device}; promptttspp_b200 runs on CUDA (sm_100a) only -- "
            "move the module and its inputs to a B200 (`.to('cuda')`), there is no CPU fallback"
        )
