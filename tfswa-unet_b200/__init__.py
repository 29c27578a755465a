"""tfswa_unet_b200 - B200-native (sm_100a) implementation of the TFSWA-UNet hot path.

Host-side mirror of the reference's ``src/models`` module API over hand-written CUDA kernels reached
through the C ABI in ``include/tfswa_b200.h`` (``lib/libtfswa_b200.so``).  There is no CPU fallback and
no alternative backend: every compute call raises if the CUDA library or a B200 is missing.
"""
from .config import get_precision, set_precision  # noqa: F401
from .attention import (FrequencySequenceAttention, MultiHeadAttention, ScaledDotProductAttention,  # noqa: F401
                        ShiftedWindowAttention, TemporalSequenceAttention, window_partition, window_reverse)
from .blocks import DownsampleBlock, TFSWABlock, UpsampleBlock  # noqa: F401
from .tfswa_unet import TFSWAUNet  # noqa: F401
from .compat import convert, install_as_reference  # noqa: F401
from .pipeline import HostPipeline  # noqa: F401

__version__ = "0.1.0"
