from .homogeneous import Stiefel  # noqa: F401
