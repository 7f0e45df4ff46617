"""Import shim: the package directory is named `whale.jl_b200` (not an importable identifier), so this module
exposes it as `whale_jl_b200` (with submodules `whale_jl_b200.lib`, `.core`, ...)."""
import os as _os

__package__ = __name__
__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "whale.jl_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
