"""Importable alias of the ``tauv-vision_b200/`` source directory.

The product directory carries the upstream project's hyphenated name, which Python cannot
import directly; this shim points the package path at it and runs its ``__init__``.
"""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "tauv-vision_b200"
__path__ = [str(_real)]
__file__ = str(_real / "__init__.py")
exec(compile((_real / "__init__.py").read_text(), __file__, "exec"))
