from .Dot_mb import Dot_mb
from .RBF_mb import RBF_mb
