"""ctypes binding of ``libpmwd_b200.so`` (the C ABI declared in ``include/pmwd_b200.h``).

There is no CPU fallback: if the shared library is missing or a call fails, an
exception is raised.  torch is used only to own device memory and streams.
"""
import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libpmwd_b200.so')

SCATTER_ATOMIC = 0
SCATTER_DETERMINISTIC = 1


class PmwdError(RuntimeError):
    pass


class CicDesc(C.Structure):
    """``pmwd_cic_desc`` (include/pmwd_b200.h)."""
    _fields_ = [
        ('dim', C.c_int32),
        ('pmid_bytes', C.c_int32),
        ('ptcl_num', C.c_int64),
        ('wrap_shape', C.c_int32 * 3),
        ('mesh_shape', C.c_int32 * 3),
        ('nchan', C.c_int32),
        ('general', C.c_int32),
        ('cell_size', C.c_double),
        ('cell_size2', C.c_double),
        ('offset', C.c_double * 3),
    ]


class Sweep(C.Structure):
    """``pmwd_sweep`` of include/pmwd_b200.h."""
    _fields_ = [
        ('table', C.c_void_p),
        ('ty', C.c_int32),
        ('bw', C.c_int32),
        ('lx', C.c_int32),
        ('nx_ext', C.c_int32),
        ('xoff', C.c_int32),
        ('reserved', C.c_int32),
        ('scratch', C.c_void_p),
        ('scratch_bytes', C.c_size_t),
        ('det_halo', C.c_void_p),
        ('det_halo_bytes', C.c_size_t),
    ]


_lib = None
_lock = threading.Lock()

_vp, _i, _i64, _f, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t
_descp = C.POINTER(CicDesc)
_i32p = C.POINTER(C.c_int32)

_SIGNATURES = {
    'pmwd_abi_version': (_i, []),
    'pmwd_last_error': (_i, [C.c_char_p, _sz]),
    'pmwd_launch_count': (C.c_longlong, []),
    'pmwd_cic_fast_path': (_i, [C.c_void_p]),
    'pmwd_profile_enable': (_i, [_i]),
    'pmwd_profile_stage_count': (_i, []),
    'pmwd_profile_stage_name': (C.c_char_p, [_i]),
    'pmwd_profile_read': (_i, [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    'pmwd_ctx_create': (_i, [C.POINTER(_vp), _i]),
    'pmwd_ctx_destroy': (_i, [_vp]),
    'pmwd_ctx_reserve': (_i, [_vp, _i, _i32p]),
    'pmwd_ctx_set_fft2d_chunk': (_i, [_vp, _i]),
    'pmwd_ctx_set_fft2d_pad': (_i, [_vp, _i]),
    'pmwd_fft_r2c': (_i, [_vp, _vp, _i, _i32p, _vp, _vp]),
    'pmwd_fft_c2r': (_i, [_vp, _vp, _i, _i32p, _vp, _vp, _f]),
    'pmwd_fft2d_r2c': (_i, [_vp, _vp, _i32p, _vp, _vp]),
    'pmwd_fft2d_c2r': (_i, [_vp, _vp, _i32p, _vp, _vp]),
    'pmwd_fft_c2c_lead': (_i, [_vp, _vp, _i, C.c_longlong, _vp, _i]),
    'pmwd_scatter_scratch_bytes': (_sz, [_descp, _i]),
    'pmwd_scatter': (_i, [_vp, _descp, _vp, _vp, _vp, _f, _vp, _i, _vp, _sz]),
    'pmwd_gather': (_i, [_vp, _descp, _vp, _vp, _vp, _vp, _f, _vp]),
    'pmwd_scatter_adj': (_i, [_vp, _descp, _vp, _vp, _vp, _vp, _f, _vp, _vp]),
    'pmwd_gather_adj': (_i, [_vp, _descp, _vp, _vp, _vp, _vp, _f, _vp, _vp]),
    'pmwd_laplace': (_i, [_vp, _i, _i32p, _d, _vp, _vp]),
    'pmwd_neg_grad': (_i, [_vp, _i, _i32p, _d, _i, _vp, _vp]),
    'pmwd_kspace_force': (_i, [_vp, _i, _i32p, _d, _f, _vp, C.POINTER(_vp)]),
    'pmwd_kspace_force_adj': (_i, [_vp, _i, _i32p, _d, _f, C.POINTER(_vp), _vp]),
    'pmwd_strain': (_i, [_vp, _i, _i32p, _d, _i, _i, _vp, _vp]),
    'pmwd_powspec_bin': (_i, [_vp, _i32p, _vp, _vp, _i, _d, _vp, _i, _i, _vp]),
    'pmwd_powspec_weight': (_i, [_vp, _i32p, _vp, _i, _d, _vp, _i, _i, _vp, _vp]),
    'pmwd_lpt_source2': (_i, [_vp, _i64, C.POINTER(_vp), _vp]),
    'pmwd_lpt_source2_vjp': (_i, [_vp, _i64, C.POINTER(_vp), _vp, C.POINTER(_vp)]),
    'pmwd_lpt_displace': (_i, [_vp, _i64, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp), _f, _f, _f, _f, _vp, _vp]),
    'pmwd_lpt_displace_vjp': (_i, [_vp, _i64, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp), _f, _f, _f, _f,
                                   C.POINTER(_vp), C.POINTER(_vp), _vp]),
    'pmwd_scatter_soa': (_i, [_vp, _descp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp]),
    'pmwd_gather3': (_i, [_vp, _descp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f]),
    'pmwd_gather3_kdk': (_i, [_vp, _descp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f]),
    'pmwd_force_adj_gather': (_i, [_vp, _descp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp]),
    'pmwd_kspace_force_slab': (_i, [_vp, _i32p, _i, _i, _d, _f, _vp, C.POINTER(_vp)]),
    'pmwd_kspace_force_adj_slab': (_i, [_vp, _i32p, _i, _i, _d, _f, C.POINTER(_vp), _vp]),
    'pmwd_xpass_supported': (_i, [_i]),
    'pmwd_xpass_last_variant': (_i, []),
    'pmwd_xpass_force': (_i, [_vp, _i32p, _i, _i, _d, _f, _vp, C.POINTER(_vp)]),
    'pmwd_xpass_force_adj': (_i, [_vp, _i32p, _i, _i, _d, _f, C.POINTER(_vp), _vp]),
    'pmwd_force_workspace_bytes': (_sz, [_descp, _i, _i]),
    'pmwd_force': (_i, [_vp, _vp, _descp, _vp, _vp, _d, _vp, _vp, _f, _i, _vp, _sz, _vp]),
    'pmwd_force_kdk': (_i, [_vp, _vp, _descp, _vp, _vp, _d, _vp, _vp, _f, _f, _f, _i, _vp, _sz, _vp]),
    'pmwd_force_adj': (_i, [_vp, _vp, _descp, _vp, _vp, _d, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    'pmwd_scatter_sweep': (_i, [_vp, _descp, _vp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp]),
    'pmwd_sweep_pick': (_i, [_descp, _i32p, _i32p]),
    'pmwd_sweep_table_bytes': (_sz, [_descp, _i, _i]),
    'pmwd_sweep_scratch_bytes': (_sz, [_descp]),
    'pmwd_sweep_table': (_i, [_vp, _descp, _i, _i, _vp, _vp, _vp]),
    'pmwd_sweep_last_stragglers': (C.c_longlong, [_vp, _vp]),
    'pmwd_sweep_usable': (_i, [_descp, _vp]),
    'pmwd_scatter_sweep_det': (_i, [_vp, _descp, _vp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp]),
    'pmwd_sweep_det_halo_bytes': (_sz, [_descp, _i, _i]),
    'pmwd_sweep_det_violations': (C.c_longlong, [_vp, _vp]),
    'pmwd_sweep_det_reset': (_i, [_vp, _vp]),
    'pmwd_cell_sort_scratch_bytes': (_sz, [_descp]),
    'pmwd_cell_sort_perm': (_i, [_vp, _descp, _vp, _vp, _vp, _vp, _sz, _i, _i]),
    'pmwd_cell_sort_sorted_keys': (_vp, [_descp, _vp]),
    'pmwd_cell_sort_perm2': (_i, [_vp, _descp, _vp, _vp, _i64, _vp, _i, _vp, _vp, _vp, _vp, _sz, _i, _i, _vp, _vp, _f]),
    'pmwd_permute_rows': (_i, [_vp, _i64, _vp, _i, C.POINTER(_vp), C.POINTER(_vp), _i32p, _i]),
    'pmwd_permute_rows2': (_i, [_vp, _i64, _vp, _i, C.POINTER(_vp), _i64, C.POINTER(_vp), C.POINTER(_vp), _i32p]),
    'pmwd_slab_owner': (_i, [_vp, _i64, _vp, _vp, _d, _i, _i, _i, _i, _vp, _vp]),
    'pmwd_transpose_p2p': (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, C.POINTER(C.c_uint64)]),
    'pmwd_transpose_ce': (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, C.POINTER(C.c_uint64), _i]),
    'pmwd_peer_copy2d': (_i, [_vp, _i, _i, _sz, _sz, _vp, _sz, _sz, C.POINTER(C.c_uint64), _sz, _sz, _i]),
    'pmwd_kick_drift': (_i, [_vp, _i64, _vp, _vp, _vp, _f, _f, _i, _i]),
    'pmwd_kick_drift_adj': (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _i, _vp]),
    'pmwd_kick_kick_drift_adj': (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _vp, _vp]),
}

EXPORTS = tuple(sorted(_SIGNATURES))


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise PmwdError(
                    f'{LIB_PATH} not found: build it with `python -m pmwd_b200.build` '
                    '(there is no CPU fallback)')
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            if handle.pmwd_abi_version() != 2:
                raise PmwdError('libpmwd_b200.so ABI version mismatch')
            _lib = handle
    return _lib


def last_error():
    buf = C.create_string_buffer(512)
    lib().pmwd_last_error(buf, 512)
    return buf.value.decode(errors='replace')


def check(rc, what):
    if rc != 0:
        raise PmwdError(f'{what} failed with status {rc}: {last_error()}')


def profile_enable(on):
    check(lib().pmwd_profile_enable(int(bool(on))), 'pmwd_profile_enable')


def profile_read():
    """{stage name: (milliseconds, calls)} accumulated since the last read (synchronises)."""
    n = lib().pmwd_profile_stage_count()
    ms = (C.c_double * n)()
    calls = (C.c_longlong * n)()
    check(lib().pmwd_profile_read(ms, calls), 'pmwd_profile_read')
    return {lib().pmwd_profile_stage_name(i).decode(): (ms[i], calls[i]) for i in range(n)}


def launch_count():
    return int(lib().pmwd_launch_count())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def shape_arr(shape):
    return (C.c_int32 * len(shape))(*[int(s) for s in shape])


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise PmwdError('pmwd_b200 kernels need CUDA tensors (there is no CPU fallback)')


class Context:
    """Owns a ``pmwd_ctx`` (cuFFT plans + work area) for one device."""

    _cache = {}

    def __init__(self, device):
        self.device = torch.device(device)
        h = _vp()
        check(lib().pmwd_ctx_create(C.byref(h), self.device.index or 0), 'pmwd_ctx_create')
        self.handle = h
        self._reserved = set()

    @classmethod
    def get(cls, device):
        """The context of (device, current CUDA stream): cuFFT plans share one work area per context, so
        two streams of one device must not share a context (``include/pmwd_b200.h``)."""
        device = torch.device(device)
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        key = (device, torch.cuda.current_stream(device).cuda_stream)
        ctx = cls._cache.get(key)
        if ctx is None:
            ctx = cls._cache[key] = cls(device)
        return ctx

    def reserve(self, shape):
        shape = tuple(int(s) for s in shape)
        if shape not in self._reserved:
            with torch.cuda.device(self.device):
                check(lib().pmwd_ctx_reserve(self.handle, len(shape), shape_arr(shape)),
                      'pmwd_ctx_reserve')
            self._reserved.add(shape)
        return self

    def __del__(self):
        try:
            if _lib is not None and self.handle:
                _lib.pmwd_ctx_destroy(self.handle)
        except Exception:
            pass
