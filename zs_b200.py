"""Import alias for the package directory `zerospeech-tts-without-t_b200/`.

The directory name the project layout prescribes is not a valid Python
identifier, so `import zs_b200` loads that directory as the package `zs_b200`
(sub-modules resolve inside it: `zs_b200.model`, `zs_b200.frontend`, ...).
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'zerospeech-tts-without-t_b200')
_spec = importlib.util.spec_from_file_location(
    'zs_b200', os.path.join(_PKG_DIR, '__init__.py'), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules['zs_b200'] = _mod
_spec.loader.exec_module(_mod)
