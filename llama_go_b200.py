"""Import shim: the product package lives in the directory `llama.go_b200/` (the name the
project brief prescribes), which is not a valid Python identifier.  `import llama_go_b200`
loads that directory as a regular package under this importable alias."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "llama.go_b200")
_spec = importlib.util.spec_from_file_location(
    "llama_go_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["llama_go_b200"] = _mod
_spec.loader.exec_module(_mod)
