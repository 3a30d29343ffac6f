"""Loader: makes `import gagan_b200` resolve to the package in `ga-gan_b200/` (a directory name Python cannot import).

After this module has run, `sys.modules['gagan_b200']` IS that package (`gagan_b200.install`, `gagan_b200.torch_utils`,
`gagan_b200.training`, ...); this file only bootstraps it.
"""
import os
import sys
import importlib.util

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ga-gan_b200')
_spec = importlib.util.spec_from_file_location('gagan_b200', os.path.join(_dir, '__init__.py'), submodule_search_locations=[_dir])
_pkg = importlib.util.module_from_spec(_spec)
sys.modules['gagan_b200'] = _pkg
_spec.loader.exec_module(_pkg)
