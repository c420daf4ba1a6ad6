"""Importable alias of the ``rna-msm_b200/`` package directory (a hyphen is not a legal Python
module name).  ``import rnamsm_b200`` resolves every submodule from ``../rna-msm_b200``."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "rna-msm_b200")
__path__.insert(0, _impl)
with open(_os.path.join(_impl, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_impl, "__init__.py"), "exec"))
del _f
