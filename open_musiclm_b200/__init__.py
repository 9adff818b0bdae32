"""Importable alias of the `open-musiclm_b200/` package directory (a hyphen is not a legal module name).

`import open_musiclm_b200` resolves every submodule from `open-musiclm_b200/`.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "open-musiclm_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
