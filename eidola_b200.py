"""Import shim: the package directory is named `cis-565-final-vr-raytracer_b200/` (hyphens, as the
task layout requires), which `import` cannot spell.  `import eidola_b200` loads it under the module
name `cis_565_final_vr_raytracer_b200` and re-exports it."""
import importlib.util
import os
import sys

_NAME = "cis_565_final_vr_raytracer_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cis-565-final-vr-raytracer_b200")


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                  submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


pkg = load()
sys.modules[__name__].__dict__.update({k: v for k, v in pkg.__dict__.items() if not k.startswith("__")})
