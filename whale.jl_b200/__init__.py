"""whale.jl_b200 — host-side mirror of the Whale.jl API over libwhalecuda (B200, sm_100a).

Scope: the one hot path of arzwa/Whale.jl — `logpdf`/`logpdf!` of the ALE/DLWGD model with its gradient,
and stochastic backtracking — behind the reference's own surface (`WhaleModel`, `DLWGD`, `ConstantDLWGD`,
`read_ale`, `logpdf`, `logpdf!`→`logpdf_`, `backtrack`; src/Whale.jl:38-39).  All arithmetic of the path
runs in CUDA kernels (csrc/whalecuda.cu) reached through the C ABI in include/whalecuda.h.

The directory name contains a dot, so import it through the shim at the repo root: `import whale_jl_b200`.
"""
from .newick import Node, readnw, getlca, getleaves, postwalk, insertnode, nwstr, extree
from .rates import ConstantDLWGD, DLWGD
from .model import WhaleModel
from .ccd import CCD, CCDVector, NativeCCDVector, read_ale, read_ale_native, save_arena, load_arena
from .core import (logpdf, logpdf_, loglikelihood, logpdf_and_gradient, logpdf_per_family, ell, slices, backtrack,
                   BacktrackFailed, logpdf_mixture, logpdf_mixture_and_gradient, logpdf_modelarray, condition, set_probe, track,
                   treekey, sumtrees, sumtrees_device)

__all__ = ["Node", "readnw", "getlca", "getleaves", "postwalk", "insertnode", "nwstr", "extree", "ConstantDLWGD",
           "DLWGD", "WhaleModel", "CCD", "CCDVector", "read_ale", "read_ale_native", "NativeCCDVector", "save_arena", "load_arena", "logpdf", "logpdf_", "loglikelihood",
           "logpdf_and_gradient", "logpdf_per_family", "ell", "slices", "backtrack", "BacktrackFailed", "logpdf_mixture", "logpdf_mixture_and_gradient", "logpdf_modelarray", "condition",
           "set_probe", "track", "treekey", "sumtrees", "sumtrees_device"]
