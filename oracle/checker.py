"""oracle/checker.py -- TEST INFRASTRUCTURE ONLY.

Picks the strongest checker available: the reference's own C++ compiled into
oracle/_ref (``oracle.ref``) when the prebuilt .so files are present, else our
plain-C restatement (``oracle.port``).  Both expose the same function names.
"""
from . import ref as _ref

if _ref.available():
    from .ref import *  # noqa: F401,F403
    KIND = "reference"
else:  # pragma: no cover - exercised on boxes without the prebuilt reference
    from .port import *  # noqa: F401,F403
    KIND = "port"
