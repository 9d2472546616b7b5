"""Import shim: ``import lzb200`` gives the package that lives in ``lzma-sdk-zliblike_b200/``
(the directory name the build contract fixes is not a valid Python identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lzma-sdk-zliblike_b200")
_name = "lzma_sdk_zliblike_b200"
if _name in sys.modules:
    _mod = sys.modules[_name]
else:
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(_dir, "__init__.py"),
                                                   submodule_search_locations=[_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
sys.modules[__name__] = _mod
