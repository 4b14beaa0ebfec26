"""Importable alias for the package directory whose name (it carries the reference's
name, hyphens and all) is not a valid Python identifier.  `import flowk` anywhere with
the repo root on sys.path gives the package; submodules import as `flowk.<name>`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "gaussian-processes-after-pre-processing-with-normalising-flows-2_b200")
_spec = importlib.util.spec_from_file_location(
    "flowk", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_module = importlib.util.module_from_spec(_spec)
sys.modules["flowk"] = _module
_spec.loader.exec_module(_module)
