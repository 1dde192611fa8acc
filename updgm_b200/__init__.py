"""Import alias: ``import updgm_b200`` loads the package that lives in
``unsupervised-probing-using-generative-diffusion-models_b200/`` (not an importable name)."""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "unsupervised-probing-using-generative-diffusion-models_b200")
__path__ = [_REAL]
__file__ = _os.path.join(_REAL, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
