"""ctypes binding of csrc/libcomemb_b200.so (the C ABI declared in include/comemb_b200.h).

There is no CPU fallback: if the shared library is missing this module raises, and every entry point raises
`ComembError` on a non-zero status.  Device memory, streams and multi-GPU plumbing come from PyTorch.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libcomemb_b200.so")

TOKEN_NONE = 0xFFFFFFFF
MODE_ORDERED, MODE_HOGWILD = 0, 1
F_DOT_FLOAT, F_ATOMIC, F_ALIAS, F_SEED_HASH = 1, 2, 4, 8

# every symbol include/comemb_b200.h declares: (restype, argtypes)
_c = ctypes
_vp, _i32, _i64, _u32, _u64, _f32, _f64 = (_c.c_void_p, _c.c_int, _c.c_int64, _c.c_uint32, _c.c_uint64, _c.c_float,
                                           _c.c_double)
SIGNATURES = {
    "comemb_init": (_i32, []),
    "comemb_get_lut": (_i32, [_vp]),
    "comemb_abi_version": (_i32, []),
    "comemb_set_opts": (_i32, [_vp]),
    "comemb_get_opts": (_i32, [_vp]),
    "comemb_set_tuning": (_i32, [_i32, _i32, _i32]),
    "comemb_set_max_warps": (_i32, [_i64]),
    "comemb_error_string": (_c.c_char_p, [_i32]),
    "comemb_o2_walks": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _u64, _vp, _u64, _vp, _u32, _i32, _i32,
                               _f32, _f32, _i32, _u32, _vp, _vp]),
    "comemb_o2_walks_sharded": (_i32, [_vp, _vp, _i32, _i64, _i32, _vp, _vp, _i64, _vp, _u64, _vp, _u64, _i32, _i32, _f32,
                                       _f32, _u32, _vp, _vp]),
    "comemb_enable_peer_access": (_i32, [_i32]),
    "comemb_ipc_open": (_i32, [_c.c_char_p, _i64, _vp]),
    "comemb_o1_edges": (_i32, [_vp, _i64, _i32, _vp, _i64, _vp, _u64, _vp, _u64, _vp, _u32, _i32, _f32, _i32, _u32,
                               _i64, _vp]),
    "comemb_o3_batch": (_i32, [_vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp, _i32, _f64, _f32, _i32, _vp]),
    "comemb_o3_batch_top1": (_i32, [_vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _f64, _f32, _i32, _vp]),
    "comemb_sg_fused_top1": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp, _vp,
                                    _i32, _i32, _i32, _f32, _f32, _f32, _u32, _vp]),
    "comemb_transpose_blocks": (_i32, [_vp, _vp, _i32, _i32, _vp]),
    "comemb_gmm_estep": (_i32, [_vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp]),
    "comemb_gmm_mstep": (_i32, [_vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp]),
    "comemb_sg_fused": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp, _i32,
                               _i32, _i32, _f32, _f32, _f32, _i32, _i32, _u32, _vp]),
    "comemb_sg_twin": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _i32, _f64, _f64, _f64, _vp, _vp, _vp, _i32, _i32,
                              _vp]),
    "comemb_walks_csr": (_i32, [_vp, _vp, _i64, _i32, _i32, _f64, _u64, _i32, _i64, _i64, _vp, _vp, _vp]),
    "comemb_downsample_walks": (_i32, [_vp, _vp, _i64, _i32, _vp, _u64, _vp]),
    "comemb_make_table": (_i32, [_vp, _i64, _f64, _vp, _i64, _vp]),
    "comemb_build_alias": (_i32, [_vp, _i64, _i64, _vp, _vp]),
    "comemb_scale": (_i32, [_vp, _i64, _f32, _vp]),
    "comemb_row_probe": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "comemb_o2_pos_loss": (_i32, [_vp, _vp, _i32, _vp, _vp, _i64, _i32, _vp, _vp]),
}


class ComembOpts(ctypes.Structure):
    """comemb_opts_t (include/comemb_b200.h): the calling thread's launch options."""
    _fields_ = [("centres_per_unit", _i32), ("max_walk_len", _i32), ("blocks_per_sm", _i32), ("variant", _i32),
                ("max_warps", _i64)]


(VARIANT_DEFAULT, VARIANT_ROUNDSYNC, VARIANT_TENSOR, VARIANT_L2_HINTS, VARIANT_ROUND1, VARIANT_ORDERED_PIPE,
 VARIANT_ORDERED_PLAIN, VARIANT_GENERIC, VARIANT_ORDERED_FLOW, VARIANT_ORDERED_TEAM) = (0, 3, 4, 5, 6, 7, 8, 9, 10, 11)


class ComembError(RuntimeError):
    pass


_lib = None
_inited_devices = set()


def load():
    """Load the shared library (no CUDA call is made)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ComembError(
                "%s is missing: build it with `python -m %s.build` (nvcc, sm_100a). There is no CPU fallback."
                % (LIB_PATH, "comemb_b200"))
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(status):
    if status != 0:
        msg = load().comemb_error_string(status)
        raise ComembError("libcomemb_b200: status %d: %s" % (status, msg.decode() if msg else "?"))


def ensure_init():
    """comemb_init() once per device (the `FAST_VERSION = init()` of pyx:549)."""
    import torch
    if not torch.cuda.is_available():
        raise ComembError("no CUDA device: the ComEmb B200 path has no CPU fallback")
    dev = torch.cuda.current_device()
    if dev not in _inited_devices:
        torch.cuda.init()
        torch.zeros(1, device="cuda")  # make sure the primary context is current
        check(load().comemb_init())
        tune = os.environ.get("COMEMB_TUNING")  # "centres_per_unit,max_walk_len,blocks_per_sm" (experiments)
        if tune:
            check(load().comemb_set_tuning(*[int(v) for v in tune.split(",")]))
        _inited_devices.add(dev)
    return 0


def get_opts():
    o = ComembOpts()
    check(load().comemb_get_opts(ctypes.byref(o)))
    return o


class opts(object):
    """with opts(max_warps=8, variant=VARIANT_GENERIC): ...  -- edits the calling thread's launch options for the block
    and restores the previous ones afterwards."""

    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.prev = get_opts()
        new = ComembOpts.from_buffer_copy(self.prev)
        for k, v in self.kw.items():
            setattr(new, k, int(v))
        check(load().comemb_set_opts(ctypes.byref(new)))
        return new

    def __exit__(self, *a):
        check(load().comemb_set_opts(ctypes.byref(self.prev)))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
