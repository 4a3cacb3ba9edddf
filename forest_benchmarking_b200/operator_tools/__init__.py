from .superoperator_transformations import *  # noqa: F401,F403
from .project_superoperators import *  # noqa: F401,F403
from .project_state_matrix import *  # noqa: F401,F403
