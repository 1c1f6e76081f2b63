"""Import alias for the product package.

The package directory is named ``movie-recommender-system-with-gnns_b200`` (not a valid Python
identifier), so ``import lgcn_b200`` loads it from that directory under this short name:
``lgcn_b200.models.light_gcn``, ``lgcn_b200.utils.train_test`` ... mirror the reference's
``models/light_gcn.py``, ``utils/train_test.py`` ... module paths.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                     "movie-recommender-system-with-gnns_b200")
_spec = _ilu.spec_from_file_location("lgcn_b200", _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["lgcn_b200"] = _mod
_spec.loader.exec_module(_mod)
