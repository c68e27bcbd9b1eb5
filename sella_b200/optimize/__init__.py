from .restricted_step import get_restricted_step  # noqa: F401
