"""Import shim: the package lives in the directory ``led-net_b200/`` (the layout the
project brief fixes), which is not a valid Python identifier.  ``import lednet_b200``
loads that directory as the package ``lednet_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'led-net_b200')
_spec = importlib.util.spec_from_file_location(
    'lednet_b200', os.path.join(_dir, '__init__.py'), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules['lednet_b200'] = _mod
_spec.loader.exec_module(_mod)
