"""sprs_b200 -- B200-native (sm_100a) implementation of the sprs sparse-product hot
path (SpMV / SpMM / SpGEMM) behind sprs's operator API.  See DESIGN.md.

Nothing here computes on the CPU: every product is a call into libsprs_b200.so
(hand-written CUDA).  Importing works without a GPU (so the ABI can be inspected);
creating a Context without one raises ThirdPartyError.
"""
from . import _lib, io, linalg
from .sparse import (CSC, CSR, Context, CsMat, CsVec, DeviceCsMat, SprsPanic, ThirdPartyError,
                     csmat_mul_csmat, prod, smmp)

__all__ = ["CSC", "CSR", "Context", "CsMat", "CsVec", "DeviceCsMat", "SprsPanic",
           "ThirdPartyError", "csmat_mul_csmat", "prod", "smmp", "_lib", "io", "linalg"]
__version__ = "0.1.0"
SPMV_TILE = 1024  # nnz per SpMV warp tile (default variant in csrc/spmv.cu); tests use it to find rows cut by a tile
