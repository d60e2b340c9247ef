"""Importable alias of the package directory `advanced-video-event-detection-extraction_b200/` (whose name, fixed
by the project layout, is not a valid Python identifier).  `import b200clip` resolves sub-modules from there."""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "advanced-video-event-detection-extraction_b200")
__path__ = [_pkg_dir]
__version__ = "0.1.0"
