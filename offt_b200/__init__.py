"""B200-native distributed 3-D complex FFT behind the C API of rchyena/offt.

The product is the C-ABI shared library `offt_b200/lib/libofft_b200.so` (sources in
`offt_b200/csrc`, headers in `include/`).  This package is only its ctypes binding,
used by the tests and by bench.py; there is no Python or CPU implementation of the
transform, and importing `offt_b200.binding` fails loudly if the library is not built.
"""
from .binding import (  # noqa: F401
    LIB_PATH, OfftError, Plan, PARAM_NAMES, P, lib, world_init, world_init_local, world_fin, world_size,
    world_rank, get_unique_id, set_default_precision, set_force_generic, execute_group, fft_launch_raw, fft_rows,
    params_default, params_range, is_infeasible_point, params_adjust, comm_box, alloc_elems, check_supported,
)
