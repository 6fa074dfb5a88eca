"""Small numerical helpers (host side)."""
from . import linalg, special  # noqa: F401
