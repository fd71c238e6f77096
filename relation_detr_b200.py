"""Import alias for the package directory ``relation-detr_b200/``.

The directory name the project layout prescribes contains a hyphen and is therefore not an
importable identifier; this one-file shim makes ``import relation_detr_b200`` (and its
submodules) resolve to that directory.  No code lives here.
"""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "relation-detr_b200")
__path__ = [_pkg_dir]
__package__ = "relation_detr_b200"
__spec__.submodule_search_locations = __path__  # makes __spec__.parent == __package__
with open(_os.path.join(_pkg_dir, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg_dir, "__init__.py"), "exec"))
del _f
