from .so import SO
from .su import SU
from .homogeneous import Stiefel, Grassmannian, OrientedGrassmannian
from .noncompact import HyperbolicSpace, SymmetricPositiveDefiniteMatrices
from .torus import Torus
from .sphere import Sphere, ProjectiveSpace
