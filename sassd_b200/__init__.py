"""Import alias: the product package lives in the directory ``sa-ssd_b200/`` (the
name the build brief fixes), which is not a legal Python identifier.  This
two-line package points its ``__path__`` there so that ``import sassd_b200``
and ``import sassd_b200.necks`` resolve to files under ``sa-ssd_b200/``."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "sa-ssd_b200")
__path__ = [_impl]
with open(_os.path.join(_impl, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_impl, "__init__.py"), "exec"))
del _f
