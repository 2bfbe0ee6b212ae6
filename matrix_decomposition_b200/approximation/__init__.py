from . import positive_semidefinite
