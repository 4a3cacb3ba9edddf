from .superoperator_transformations import *  # noqa: F401,F403
