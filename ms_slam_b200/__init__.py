"""ms_slam_b200: B200-native sliding-window map sparsification (the MapSparsification hot path of MS-SLAM).

Only the hot path lives here: the flattened window view, the synthetic window generator, the ctypes binding to the
CUDA engine (``libmss.so``, C-ABI in ``include/mss.h``) and the multi-GPU window sharding. There is no CPU fallback:
``ms_slam_b200.engine`` raises if the CUDA library is missing.
"""
from .window import WindowView, PackedView, make_view, pack_view, merge_views, split_components, GRID_COLS, GRID_ROWS, N_CELLS, CELL_NONE  # noqa: F401

__version__ = "0.1.0"
