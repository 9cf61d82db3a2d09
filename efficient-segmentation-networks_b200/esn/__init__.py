"""esn: host-side binding of libesn_sm100.so (B200 kernels for the segmentation hot path)."""
from . import _lib, ops  # noqa: F401  (raises ImportError when the CUDA library is missing)
from .ops import launch_count, launch_count_reset  # noqa: F401
