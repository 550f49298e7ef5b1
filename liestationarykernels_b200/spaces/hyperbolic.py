from .noncompact import HyperbolicSpace  # noqa: F401
