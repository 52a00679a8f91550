"""makisu_b200 -- B200-native snapshot+hash engine for uber/makisu's build-context
fingerprint (CRC-32 cacheID) and layer digest path.  See DESIGN.md.

The product is libmksnap.so (CUDA, sm_100a) behind include/mksnap.h; this package is
the Python face of that C-ABI plus the host-side mirror of the reference interface
used by the tests and the bench.  There is no CPU fallback.
"""
from .abi import Engine, Extent, Range, Result, MksnapError, MKSNAP_X_CRC, MKSNAP_X_CDC  # noqa: F401
