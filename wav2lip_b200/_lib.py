"""ctypes binding of libw2l.so — the declarations of include/w2l.h, nothing else.

The library is built in-tree by `python -c "import __graft_entry__ as g; g.build()"`
(nvcc -gencode arch=compute_100a,code=sm_100a).  If it is missing, importing this module still
works (so that error messages are useful) but any use raises W2LError: there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))

W2L_OK, W2L_EINVAL, W2L_ENODEV, W2L_ECUDA, W2L_ENOMEM, W2L_ESTATE = 0, -1, -2, -3, -4, -5
NET_GENERATOR, NET_SYNCNET, NET_DISC, NET_S3FD = 0, 1, 2, 3
BLOCK_CONV_BN_RELU, BLOCK_CONVT_BN_RELU, BLOCK_CONV_LRELU, BLOCK_CONV_PLAIN, BLOCK_CONV_RELU = 0, 1, 2, 3, 4
PREC_F16, PREC_BF16, PREC_F32X = 0, 1, 2
TRAIN_WGRAD, TRAIN_ACCUMULATE, TRAIN_INPUT_GRAD, TRAIN_NO_STAT_UPDATE = 1, 2, 4, 8

# every symbol include/w2l.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = [
    "w2l_abi_version", "w2l_last_error", "w2l_net_num_layers", "w2l_net_layer_info",
    "w2l_create", "w2l_destroy", "w2l_load_weights",
    "w2l_generator_forward", "w2l_generator_forward_host", "w2l_generator_forward_u8", "w2l_generator_forward_u8_host",
    "w2l_generator_submit_host", "w2l_generator_submit_u8_host", "w2l_host_wait",
    "w2l_syncnet_forward", "w2l_syncnet_forward_frames", "w2l_cosine_bce_loss", "w2l_l1_loss", "w2l_disc_forward",
    "w2l_conv_block_forward", "w2l_debug_layer_output",
    "w2l_melspectrogram", "w2l_melspectrogram_host", "w2l_mel_num_frames", "w2l_mel_num_chunks", "w2l_mel_chunks",
    "w2l_set_debug", "w2l_mel_basis_host", "w2l_launch_count", "w2l_device_bytes", "w2l_profile_plan",
    "w2l_f16_overflow",
    "w2l_crop_resize_u8", "w2l_paste_u8", "w2l_lipsync_frames_u8", "w2l_s3fd_out_dims", "w2l_s3fd_forward",
    "w2l_train_bind", "w2l_train_forward", "w2l_train_backward", "w2l_adam_step", "w2l_wav2lip_train_step",
    "w2l_train_last_output", "w2l_train_flops", "w2l_comm_unique_id", "w2l_comm_init", "w2l_conv_block_train", "w2l_train_profile",
]


class W2LError(RuntimeError):
    pass


class LayerInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("kind", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
                ("kh", C.c_int32), ("kw", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
                ("ph", C.c_int32), ("pw", C.c_int32), ("out_pad", C.c_int32), ("residual", C.c_int32), ("cout_real", C.c_int32)]


def lib_path() -> str:
    return os.environ.get("W2L_LIB", os.path.join(HERE, "libw2l.so"))


_lib: Optional[C.CDLL] = None


def get_lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise W2LError(
            f"{path} not found: build it with `python -c \"import __graft_entry__ as g; g.build()\"`. "
            "wav2lip_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(path)
    vp, i32, i64, cp = C.c_void_p, C.c_int, C.c_int64, C.c_char_p
    lib.w2l_abi_version.restype = i32
    lib.w2l_last_error.restype = cp
    lib.w2l_net_num_layers.argtypes = [i32]
    lib.w2l_net_layer_info.argtypes = [i32, i32, C.POINTER(LayerInfo)]
    lib.w2l_create.argtypes = [i32, i32, C.POINTER(vp)]
    lib.w2l_destroy.argtypes = [vp]
    lib.w2l_set_debug.argtypes = [vp, i32]
    lib.w2l_load_weights.argtypes = [vp, i32, i32, C.POINTER(cp), C.POINTER(vp), C.POINTER(i64), vp]
    lib.w2l_generator_forward.argtypes = [vp, vp, vp, vp, i32, i32, vp]
    lib.w2l_generator_forward_host.argtypes = [vp, vp, vp, vp, i32, i32]
    lib.w2l_generator_forward_u8.argtypes = [vp, vp, vp, vp, i32, vp]
    lib.w2l_generator_forward_u8_host.argtypes = [vp, vp, vp, vp, i32]
    lib.w2l_generator_submit_host.argtypes = [vp, vp, vp, vp, i32, i32]
    lib.w2l_generator_submit_u8_host.argtypes = [vp, vp, vp, vp, i32]
    lib.w2l_host_wait.argtypes = [vp, i32]
    lib.w2l_mel_num_chunks.argtypes = [i64, C.c_double]
    lib.w2l_mel_num_chunks.restype = i64
    lib.w2l_mel_chunks.argtypes = [vp, vp, i64, C.c_double, vp, i64, vp]
    lib.w2l_syncnet_forward.argtypes = [vp, vp, vp, vp, vp, i32, vp]
    lib.w2l_disc_forward.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.w2l_syncnet_forward_frames.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp]
    lib.w2l_cosine_bce_loss.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]
    lib.w2l_l1_loss.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.w2l_conv_block_forward.argtypes = [vp, C.POINTER(LayerInfo), vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.w2l_debug_layer_output.argtypes = [vp, i32, i32, vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), vp]
    lib.w2l_melspectrogram.argtypes = [vp, vp, i64, vp, vp]
    lib.w2l_melspectrogram_host.argtypes = [vp, vp, i64, vp]
    lib.w2l_mel_num_frames.argtypes = [i64]
    lib.w2l_mel_num_frames.restype = i64
    lib.w2l_mel_basis_host.argtypes = [vp]
    lib.w2l_launch_count.argtypes = [vp]
    lib.w2l_launch_count.restype = i64
    lib.w2l_device_bytes.argtypes = [vp]
    lib.w2l_device_bytes.restype = i64
    lib.w2l_profile_plan.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp]
    lib.w2l_f16_overflow.argtypes = [vp, i32, C.POINTER(i32), vp]
    f32 = C.c_float
    lib.w2l_s3fd_out_dims.argtypes = [i32, i32, C.POINTER(i32)]
    lib.w2l_s3fd_forward.argtypes = [vp, vp, C.POINTER(vp), i32, i32, i32, vp]
    lib.w2l_crop_resize_u8.argtypes = [vp, vp, i32, i32, i32, C.POINTER(i32), i32, vp, vp]
    lib.w2l_paste_u8.argtypes = [vp, vp, vp, i32, i32, i32, C.POINTER(i32), i32, vp, vp]
    lib.w2l_lipsync_frames_u8.argtypes = [vp, vp, vp, i32, i32, i32, C.POINTER(i32), i32, vp, vp]
    lib.w2l_train_bind.argtypes = [vp, i32, i32, C.POINTER(cp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]
    lib.w2l_train_forward.argtypes = [vp, i32, vp, vp, vp, vp, i32, i32, i32, vp]
    lib.w2l_train_backward.argtypes = [vp, i32, vp, vp, vp, i32, vp]
    lib.w2l_adam_step.argtypes = [vp, i32, f32, f32, f32, f32, vp]
    lib.w2l_wav2lip_train_step.argtypes = [vp, vp, vp, vp, vp, i32, i32, f32, f32, vp, vp]
    lib.w2l_train_last_output.argtypes = [vp, vp, i64, vp]
    lib.w2l_train_flops.argtypes = [vp, i32]
    lib.w2l_train_flops.restype = C.c_double
    lib.w2l_train_profile.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp]
    lib.w2l_comm_unique_id.argtypes = [vp, C.c_char_p]
    lib.w2l_comm_init.argtypes = [vp, C.c_char_p, i32, i32]
    lib.w2l_conv_block_train.argtypes = [vp, C.POINTER(LayerInfo), vp, i32, i32, i32] + [vp] * 14
    for name in EXPORTS:
        getattr(lib, name)  # AttributeError here == header / library mismatch
    if lib.w2l_abi_version() != 1:
        raise W2LError(f"ABI version mismatch: library {lib.w2l_abi_version()}, binding 1")
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != W2L_OK:
        msg = get_lib().w2l_last_error().decode("utf-8", "replace")
        raise W2LError(f"libw2l error {code}: {msg}")


def net_layers(net: int):
    """The architecture table of `net` as a list of dicts (host only, no GPU needed)."""
    lib = get_lib()
    n = lib.w2l_net_num_layers(net)
    if n < 0:
        check(n)
    out = []
    for i in range(n):
        li = LayerInfo()
        check(lib.w2l_net_layer_info(net, i, C.byref(li)))
        out.append({"name": li.name.decode(), "kind": li.kind, "cin": li.cin, "cout": li.cout,
                    "k": (li.kh, li.kw), "stride": (li.sh, li.sw), "pad": (li.ph, li.pw),
                    "out_pad": li.out_pad, "residual": bool(li.residual), "cout_real": li.cout_real})
    return out


class Context:
    """Owns one w2l_ctx (one device, one precision).  Not thread safe."""

    def __init__(self, device: int = 0, precision: int = PREC_F16):
        self.lib = get_lib()
        h = C.c_void_p()
        check(self.lib.w2l_create(int(device), int(precision), C.byref(h)))
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.w2l_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self) -> int:
        return int(self.lib.w2l_launch_count(self.h))

    def device_bytes(self) -> int:
        return int(self.lib.w2l_device_bytes(self.h))

    def f16_overflow(self, clear: bool = True, stream: int = 0) -> bool:
        """True if an fp16 epilogue stored an inf/NaN activation since the flag was last cleared (synchronises)."""
        flag = C.c_int(0)
        check(self.lib.w2l_f16_overflow(self.h, 1 if clear else 0, C.byref(flag), C.c_void_p(stream)))
        return bool(flag.value)

    def set_debug(self, keep_all: bool):
        check(self.lib.w2l_set_debug(self.h, 1 if keep_all else 0))

    def load_weights(self, net: int, tensors: dict, stream: int = 0):
        """tensors: name -> (device_ptr, numel) of fp32 contiguous CUDA tensors."""
        names = list(tensors.keys())
        n = len(names)
        arr_n = (C.c_char_p * n)(*[s.encode() for s in names])
        arr_p = (C.c_void_p * n)(*[tensors[s][0] for s in names])
        arr_c = (C.c_int64 * n)(*[tensors[s][1] for s in names])
        check(self.lib.w2l_load_weights(self.h, net, n, arr_n, arr_p, arr_c, C.c_void_p(stream)))

    def train_profile(self, net: int, iters: int = 3, stream: int = 0, cap: int = 512):
        ms = (C.c_float * cap)()
        fl = (C.c_double * cap)()
        names = ((C.c_char * 64) * cap)()
        k = self.lib.w2l_train_profile(self.h, net, iters, cap, ms, fl, names, C.c_void_p(stream))
        if k < 0:
            check(k)
        return [(names[i].value.decode(), float(ms[i]), float(fl[i])) for i in range(k)]

    def profile_plan(self, net: int, iters: int = 5, stream: int = 0, cap: int = 256):
        ms = (C.c_float * cap)()
        fl = (C.c_double * cap)()
        names = ((C.c_char * 64) * cap)()
        k = self.lib.w2l_profile_plan(self.h, net, iters, cap, ms, fl, names, C.c_void_p(stream))
        if k < 0:
            check(k)
        return [(names[i].value.decode(), float(ms[i]), float(fl[i])) for i in range(k)]
