"""ctypes binding of libdyncore.so (include/dyncore.h).

The product path has NO CPU fallback: if the CUDA library is missing or cannot be loaded
this module raises.  (The CPU test-suite injects tests/emu/libdyncore_emu.so, a host
emulation built from the same kernel bodies, through `use_library`; nothing in this package
does that by itself.)
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIBRARY = os.path.join(_HERE, 'libdyncore.so')

_dp = ctypes.POINTER(ctypes.c_double)

GRID_FIELDS_2D = ['A', 'dxjs', 'dyis', 'corf', 'corf_is', 'lat_rad', 'lat_is_rad', 'dlon_rad',
                  'dlat_rad']
GRID_FIELDS_1D = ['sigma_vb', 'dsigma', 'UVFLX_dif_coef', 'POTT_dif_coef', 'moist_dif_coef']

DC_NK_2D, DC_NK_NZ, DC_NK_NZS = 0, 1, 2
# registry fields that only exist for the physics coupling terms (dc_grid_desc.i_coupling)
COUPLING_ONLY_FIELDS = ['KMOM', 'KHEAT', 'SMOMXFLX', 'SMOMYFLX', 'SSHFLX', 'SLHFLX',
                        'KMOM_dUWINDdz', 'KMOM_dVWINDdz', 'dUFLXdt_TURB', 'dVFLXdt_TURB',
                        'dPOTTdt_TURB', 'dQVdt_TURB', 'dPOTTdt_RAD']
DC_MODE_FUSED, DC_MODE_KERNELS = 0, 1
DC_COMM_ID_BYTES = 128
DC_P2P_HANDLE_BYTES = 256
DC_PART_ALL, DC_PART_CONT, DC_PART_BOUNDARY, DC_PART_INTERIOR, DC_PART_COLP = 0, 1, 2, 3, 4


class GridDesc(ctypes.Structure):
    """dc_grid_desc (include/dyncore.h)"""
    _fields_ = ([('nx', ctypes.c_int), ('ny', ctypes.c_int), ('nz', ctypes.c_int),
                 ('j0', ctypes.c_int), ('j1', ctypes.c_int), ('i_moist', ctypes.c_int),
                 ('dt', ctypes.c_double), ('pair_top', ctypes.c_double)] +
                [(n, _dp) for n in GRID_FIELDS_2D + GRID_FIELDS_1D] +
                [('i_coupling', ctypes.c_int)])


class DyncoreError(RuntimeError):
    def __init__(self, code, message):
        super().__init__('libdyncore error %d: %s' % (code, message))
        self.code = code


_lib = None
_lib_path = None

_ENTRIES = ['dc_continuity', 'dc_momentum', 'dc_temperature', 'dc_moisture',
            'dc_compute_tendencies', 'dc_euler_forward', 'dc_primary_diag', 'dc_secondary_diag',
            'dc_compute_turbulence']


def _declare(lib):
    vp = ctypes.c_void_p
    lib.dc_last_error.restype = ctypes.c_char_p
    lib.dc_is_cuda.restype = ctypes.c_int
    lib.dc_create.argtypes = [ctypes.POINTER(GridDesc), ctypes.POINTER(vp)]
    lib.dc_destroy.argtypes = [vp]
    lib.dc_get_layout.argtypes = [vp] + [ctypes.POINTER(ctypes.c_int)] * 3
    lib.dc_num_fields.restype = ctypes.c_int
    lib.dc_field_name.argtypes = [ctypes.c_int]
    lib.dc_field_name.restype = ctypes.c_char_p
    lib.dc_field_id.argtypes = [ctypes.c_char_p]
    lib.dc_field_info.argtypes = [ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 3
    lib.dc_bind_field.argtypes = [vp, ctypes.c_int, vp, ctypes.c_size_t]
    for e in _ENTRIES:
        getattr(lib, e).argtypes = [vp, vp]
    lib.dc_exchange_bc.argtypes = [vp, ctypes.c_int, vp]
    lib.dc_step_matsuno.argtypes = [vp, ctypes.c_int, vp]
    lib.dc_set_mode.argtypes = [vp, ctypes.c_int]
    lib.dc_step_begin.argtypes = [vp, vp]
    lib.dc_stage_compute.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp]
    lib.dc_stage_diag.argtypes = [vp, ctypes.c_int, vp]
    lib.dc_halo_bytes.argtypes = [vp, ctypes.POINTER(ctypes.c_size_t)]
    lib.dc_halo_pack.argtypes = [vp, ctypes.c_int, vp, vp, vp]
    lib.dc_halo_unpack.argtypes = [vp, ctypes.c_int, vp, vp, vp]
    lib.dc_comm_unique_id.argtypes = [vp, ctypes.c_size_t]
    lib.dc_set_comm.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int]
    lib.dc_has_comm.argtypes = [vp]
    lib.dc_comm_p2p_handles.argtypes = [vp, vp, ctypes.c_size_t]
    lib.dc_comm_p2p_connect.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    lib.dc_comm_p2p_enable.argtypes = [vp, ctypes.c_int]
    lib.dc_halo_exchange.argtypes = [vp, ctypes.c_int, vp]
    lib.dc_run_diag_bytes.argtypes = [vp, ctypes.POINTER(ctypes.c_size_t)]
    lib.dc_run_diag.argtypes = [vp, vp, ctypes.c_size_t, vp]
    lib.dc_import_field.argtypes = [vp, ctypes.c_int, vp, ctypes.c_size_t, vp]
    lib.dc_export_field.argtypes = [vp, ctypes.c_int, vp, ctypes.c_size_t, vp]
    lib.dc_import_rows.argtypes = [vp, ctypes.c_int, vp, ctypes.c_size_t, ctypes.c_int,
                                   ctypes.c_int, vp]
    lib.dc_export_rows.argtypes = [vp, ctypes.c_int, vp, ctypes.c_size_t, ctypes.c_int,
                                   ctypes.c_int, vp]
    lib.dc_profile_enable.argtypes = [vp, ctypes.c_int]
    lib.dc_profile_read.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p),
                                    ctypes.POINTER(ctypes.c_double),
                                    ctypes.POINTER(ctypes.c_longlong)]
    lib.dc_launch_count.argtypes = [vp]
    lib.dc_launch_count.restype = ctypes.c_longlong
    return lib


def use_library(path):
    """load a specific build of the library (tests: the host emulation)"""
    global _lib, _lib_path
    _lib = _declare(ctypes.CDLL(path))
    _lib_path = path
    return _lib


def lib():
    """the loaded library; loads climate_model_b200/libdyncore.so on first use"""
    if _lib is None:
        if not os.path.exists(DEFAULT_LIBRARY):
            raise ImportError(
                'climate_model_b200/libdyncore.so is missing: build it with '
                '`python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a). '
                'There is no CPU fallback.')
        use_library(DEFAULT_LIBRARY)
    return _lib


def library_path():
    return _lib_path


def is_cuda():
    return bool(lib().dc_is_cuda())


def check(code):
    if code != 0:
        raise DyncoreError(code, lib().dc_last_error().decode())


def field_table():
    """{name: (id, stgx, stgy, nk_kind)} from the library's registry"""
    L = lib()
    out = {}
    for i in range(L.dc_num_fields()):
        sx, sy, nk = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(L.dc_field_info(i, ctypes.byref(sx), ctypes.byref(sy), ctypes.byref(nk)))
        out[L.dc_field_name(i).decode()] = (i, sx.value, sy.value, nk.value)
    return out


def timeline_read(handle, max_entries=256):
    """[(mark name, ms since the first mark)] of the banded steps enqueued while
    dc_profile_enable(h, 2) was on (include/dyncore.h)"""
    names = (ctypes.c_char_p * max_entries)()
    ms = (ctypes.c_double * max_entries)()
    n = (ctypes.c_longlong * max_entries)()
    cnt = lib().dc_profile_read(handle, max_entries, names, ms, n)
    return [(names[i].decode(), ms[i]) for i in range(cnt)]


def profile_read(handle, max_entries=64):
    """{kernel name: (total ms, launches)} since the last read (dc_profile_read)"""
    names = (ctypes.c_char_p * max_entries)()
    ms = (ctypes.c_double * max_entries)()
    n = (ctypes.c_longlong * max_entries)()
    cnt = lib().dc_profile_read(handle, max_entries, names, ms, n)
    return {names[i].decode(): (ms[i], n[i]) for i in range(cnt)}
