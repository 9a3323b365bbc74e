"""Import-path shim (see dropin/README.md): hot-path modules come from hanabizero_b200; anything else of
this package keeps resolving to the reference checkout named by $HANABIZERO_REFERENCE."""
import os as _os

_ref = _os.environ.get("HANABIZERO_REFERENCE")
if _ref and _os.path.isdir(_os.path.join(_ref, 'core')):
    __path__.append(_os.path.join(_ref, 'core'))
