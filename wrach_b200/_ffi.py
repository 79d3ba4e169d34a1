"""ctypes binding of include/wrach_cuda.h (the C ABI of the CUDA worker).

Fails loudly when the compiled library is missing: there is no CPU or eager fallback anywhere in
this package.  Build it with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C wrach_b200/csrc`.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WRACH_CUDA_LIB", os.path.join(_HERE, "lib", "libwrach_cuda.so"))  # override: tuning sweeps


class WorldSettings(ctypes.Structure):
    """runners/bevy/src/config_shader.rs:15-29 — the 32-byte uniform."""
    _fields_ = [
        ("view_dimensions", ctypes.c_float * 2),
        ("view_anchor", ctypes.c_float * 2),
        ("grid_dimensions", ctypes.c_uint32 * 2),
        ("cell_size", ctypes.c_uint32),
        ("particles_in_frame_count", ctypes.c_uint32),
    ]

    def copy(self):
        c = WorldSettings()
        ctypes.memmove(ctypes.byref(c), ctypes.byref(self), ctypes.sizeof(self))
        return c

    def __repr__(self):
        return ("WorldSettings(view_dimensions=(%g, %g), view_anchor=(%g, %g), grid_dimensions=(%d, %d), "
                "cell_size=%d, particles_in_frame_count=%d)") % (
            self.view_dimensions[0], self.view_dimensions[1], self.view_anchor[0], self.view_anchor[1],
            self.grid_dimensions[0], self.grid_dimensions[1], self.cell_size, self.particles_in_frame_count)


class Stats(ctypes.Structure):
    _fields_ = [
        ("steps_completed", ctypes.c_uint64),
        ("kernel_launches", ctypes.c_uint64),
        ("slow_path_steps", ctypes.c_uint64),
        ("halo_bytes_sent", ctypes.c_uint64),
        ("last_phys_ms", ctypes.c_float),
        ("last_rebin_ms", ctypes.c_float),
        ("phys_launches_last", ctypes.c_uint32),
        ("rebin_launches_last", ctypes.c_uint32),
        ("tile_frames", ctypes.c_uint64),
        ("tile_fallbacks", ctypes.c_uint64),
        ("tile_packs", ctypes.c_uint64),
        ("tile_unpacks", ctypes.c_uint64),
    ]


# enum wrach_buffer
WORLD_SETTINGS_UNIFORM, INDICES_MAIN, INDICES_BLOCK_SUMS, POSITIONS_IN, POSITIONS_OUT, VELOCITIES_IN, \
    VELOCITIES_OUT = range(7)
# enum wrach_status
OK, ERR_BAD_ARG, ERR_CAPACITY, ERR_CUDA, ERR_NCCL, ERR_STATE, ERR_FAR_MIGRATION = 0, -1, -2, -3, -4, -5, -6
# enum wrach_arith
ARITH_UNFUSED, ARITH_SPV = 0, 1

# every symbol include/wrach_cuda.h declares: name -> (restype, argtypes)
_P = ctypes.c_void_p
_SP = ctypes.POINTER(WorldSettings)
SYMBOLS = {
    "wrach_cuda_create": (ctypes.c_int, [_SP, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                         ctypes.POINTER(_P)]),
    "wrach_cuda_create_strip": (ctypes.c_int, [_SP, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, _P, ctypes.POINTER(_P)]),
    "wrach_cuda_nccl_unique_id": (ctypes.c_int, [_P]),
    "wrach_cuda_strip_info": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32),
                                             ctypes.POINTER(ctypes.c_uint32)]),
    "wrach_cuda_strip_group_step": (ctypes.c_int, [ctypes.POINTER(_P), ctypes.c_int, ctypes.c_uint32]),
    "wrach_cuda_strip_columns": (None, [ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]),
    "wrach_cuda_destroy": (None, [_P]),
    "wrach_cuda_write_slice": (ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.c_size_t]),
    "wrach_cuda_write_settings": (ctypes.c_int, [_P, _SP]),
    "wrach_cuda_step": (ctypes.c_int, [_P, ctypes.c_uint32]),
    "wrach_cuda_ready": (ctypes.c_int, [_P]),
    "wrach_cuda_sync": (ctypes.c_int, [_P]),
    "wrach_cuda_read": (ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.c_size_t]),
    "wrach_cuda_read_async": (ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.c_size_t]),
    "wrach_cuda_host_register": (ctypes.c_int, [_P, ctypes.c_size_t]),
    "wrach_cuda_host_unregister": (ctypes.c_int, [_P]),
    "wrach_cuda_buffer_bytes": (ctypes.c_size_t, [_P, ctypes.c_int]),
    "wrach_cuda_device_pointer": (_P, [_P, ctypes.c_int]),
    "wrach_cuda_export_buffer_fd": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_size_t)]),
    "wrach_cuda_settle": (ctypes.c_int, [_P]),
    "wrach_cuda_selftest_import_fd": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, _P, ctypes.c_size_t]),
    "wrach_cuda_last_error": (ctypes.c_char_p, [_P]),
    "wrach_cuda_alloc_host": (_P, [ctypes.c_size_t]),
    "wrach_cuda_free_host": (None, [_P]),
    "wrach_cuda_set_neighbour_mode": (ctypes.c_int, [_P, ctypes.c_int]),
    "wrach_cuda_step_timed": (ctypes.c_int, [_P, ctypes.c_uint32, ctypes.POINTER(ctypes.c_float)]),
    "wrach_cuda_step_profiled": (ctypes.c_int, [_P, ctypes.c_uint32, ctypes.POINTER(ctypes.c_float),
                                                ctypes.POINTER(ctypes.c_float)]),
    "wrach_cuda_get_stats": (ctypes.c_int, [_P, ctypes.POINTER(Stats)]),
    "wrach_cuda_selftest_push_division": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]),
    "wrach_cuda_selftest_push_sqrt": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]),
    "wrach_cuda_version": (ctypes.c_char_p, []),
}

_lib = None


def lib():
    """Load libwrach_cuda.so; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "wrach_b200: %s is missing. The CUDA extension is the product and there is no fallback; "
                "build it with `make -C wrach_b200/csrc` (needs nvcc, sm_100a)." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class WrachCudaError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("wrach_cuda status %d: %s" % (status, message))
        self.status = status


def check(status, handle=None):
    if status < 0:
        msg = lib().wrach_cuda_last_error(handle)
        raise WrachCudaError(status, msg.decode() if msg else "")
    return status
