"""dsrc_b200 -- B200-native DSRC block codec (C ABI in include/dsrc_b200.h, kernels in dsrc_b200/csrc)."""
from .block_compressor import BlockCompressor, DsrcGpuError  # noqa: F401
from .operators import DsrcCompressorMT, DsrcDecompressorMT, DsrcModule, InputParameters  # noqa: F401,E402
