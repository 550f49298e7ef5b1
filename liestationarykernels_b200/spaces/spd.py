from .noncompact import SymmetricPositiveDefiniteMatrices  # noqa: F401
