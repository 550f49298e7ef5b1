from .homogeneous import Grassmannian, OrientedGrassmannian, _SOxSO, _S_OxO  # noqa: F401
