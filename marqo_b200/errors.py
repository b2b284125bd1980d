"""Error classes with the reference's names and hierarchy (src/marqo/s2_inference/errors.py:4-73,
src/marqo/api/exceptions.py, src/marqo/vespa/exceptions.py) so that `except` clauses written against Marqo keep
working.  When the real `marqo` package is importable its own classes are re-exported instead, so exceptions raised
here are caught by Marqo's handlers unchanged."""
from __future__ import annotations

from typing import Optional

try:  # pragma: no cover - only where Marqo's own environment exists
    from marqo.s2_inference.errors import (  # type: ignore
        S2InferenceError, VectoriseError, InvalidModelPropertiesError, UnknownModelError, ModelLoadError,
        ModelDownloadError, ModelNotInCacheError, IncompatibleModelDeviceError, UnsupportedModalityError)
    from marqo.api.exceptions import InternalError, ModelCacheManagementError, ConfigurationError  # type: ignore
    from marqo.vespa.exceptions import VespaError, VespaStatusError  # type: ignore
    MARQO_AVAILABLE = True
except Exception:  # the build / GPU containers: Marqo itself is not importable
    MARQO_AVAILABLE = False

    class S2InferenceError(Exception):
        def __init__(self, message: Optional[str] = None) -> None:
            self.message = message
            super().__init__(self.message)

    class VectoriseError(S2InferenceError):
        pass

    class InvalidModelPropertiesError(S2InferenceError):
        pass

    class UnknownModelError(S2InferenceError):
        pass

    class ModelLoadError(S2InferenceError):
        pass

    class ModelDownloadError(S2InferenceError):
        pass

    class ModelNotInCacheError(S2InferenceError):
        pass

    class IncompatibleModelDeviceError(S2InferenceError):
        pass

    class UnsupportedModalityError(S2InferenceError):
        pass

    class MarqoError(Exception):
        def __init__(self, message: Optional[str] = None) -> None:
            self.message = message
            super().__init__(self.message)

    class InternalError(MarqoError):
        pass

    class ModelCacheManagementError(MarqoError):
        pass

    class ConfigurationError(MarqoError):
        pass

    class VespaError(Exception):
        pass

    class VespaStatusError(VespaError):
        @property
        def status_code(self) -> int:
            return self.args[0] if self.args else 500
