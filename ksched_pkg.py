"""Import shim: the package directory is named after the reference repo
(`kube-scheduler-rs-reference_b200/`, not a valid Python identifier), so it is loaded under the
module name `ksched_b200`.  Usage: `import ksched_pkg; ks = ksched_pkg.load()`."""
import importlib.util
import os
import sys

_NAME = "ksched_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kube-scheduler-rs-reference_b200")


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR]
    )
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
