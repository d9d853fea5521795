"""Importable alias of the package directory `nodeembedding-to-communityembedding_b200/` (a hyphenated directory name
cannot be imported directly).  `import comemb_b200` == that package."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "nodeembedding-to-communityembedding_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
