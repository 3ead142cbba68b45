from . import svd  # noqa: F401
