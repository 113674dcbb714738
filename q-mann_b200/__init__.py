"""qmann_b200 -- B200-native quantized MemN2N inference forward (host-side Python mirror).

The product is libqmann_b200.so (csrc/, C ABI in include/qmann_abi.h); this package only loads it
and mirrors the reference's layer interface for tests and benchmarks."""
from . import synth  # noqa: F401
from . import lib  # noqa: F401
from . import babi  # noqa: F401
from . import weights_io  # noqa: F401
