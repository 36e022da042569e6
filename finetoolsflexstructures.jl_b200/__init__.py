"""finetoolsflexstructures.jl_b200 -- B200-native element-level hot path of
FinEtoolsFlexStructures.jl behind the reference's operator API.

Contents: `csrc/` (sm_100a CUDA kernels + the C ABI of include/fsgpu.h, built into
`libfsgpu.so`), `julia/` (the `ccall` glue a Julia host loads), and this Python mirror of
the reference interface (`femm`), used by tests/ and bench.py because the build image has
no Julia.  Import as `import fsb200` (see /fsb200.py: the directory name contains a dot).
"""
from . import _lib
from ._lib import FsgpuError, BeamParams, ShellParams, EXPORTED_SYMBOLS, LIB_PATH
from .context import Context, Explicit, SparseMatrixCSC
from . import femm
from . import partition
