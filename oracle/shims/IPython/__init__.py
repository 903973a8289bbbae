from . import display  # noqa: F401
