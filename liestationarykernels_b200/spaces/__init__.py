from .so import SO
from .su import SU
