"""Import shim: ``import mcgm_b200`` loads the package that lives in the directory
``multimodal-controller-for-generative-models_b200/`` (a name Python cannot import
directly because of the hyphens)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "multimodal-controller-for-generative-models_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
