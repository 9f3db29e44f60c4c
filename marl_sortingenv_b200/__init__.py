"""Importable alias of the on-disk package directory `marl-sortingenv_b200/`.

The build contract names the package directory with a hyphen, which Python cannot import.
This shim makes `import marl_sortingenv_b200` resolve to that directory: it points the
package search path there and executes the real `__init__.py` in this namespace, so
`marl_sortingenv_b200.batched`, `._abi`, `.config`, ... are the files under
`marl-sortingenv_b200/`.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "marl-sortingenv_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py"), "r") as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
