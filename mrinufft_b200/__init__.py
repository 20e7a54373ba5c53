"""mrinufft_b200 -- B200-native NUFFT backend for mri-nufft.

Importing this package registers ``get_operator("b200")`` in mri-nufft's own backend registry
(``FourierOperatorBase.__init_subclass__``, ``src/mrinufft/operators/base.py:211-218``); there is no
entry-point discovery for out-of-tree backends, so ``import mrinufft_b200`` must run before
``mrinufft.get_operator("b200")``.

The package needs ``mrinufft`` itself (the reference's registry and base classes).  If it is not
installed in the environment, the unmodified reference install under ``baseline/_ref`` (created by
``pip install --no-deps --target baseline/_ref <reference>``) is put on ``sys.path``.
"""

from __future__ import annotations

import sys
from pathlib import Path

try:  # pragma: no cover - depends on the environment
    import mrinufft  # noqa: F401
except ModuleNotFoundError:  # pragma: no cover
    _ref = Path(__file__).resolve().parent.parent / "baseline" / "_ref"
    if (_ref / "mrinufft").is_dir():
        sys.path.insert(0, str(_ref))
        import mrinufft  # noqa: F401
    else:
        raise ModuleNotFoundError(
            "mrinufft_b200 is a backend plug-in for mri-nufft: install `mri-nufft` or provide "
            f"the reference install at {_ref}"
        ) from None

from . import _lib  # noqa: E402
from .operator import MRIB200NUFFT, RawB200Plan  # noqa: E402
from .stacked import MRIB200StackedNUFFT  # noqa: E402

__all__ = ["MRIB200NUFFT", "MRIB200StackedNUFFT", "RawB200Plan", "_lib"]
__version__ = "0.1.0"
