from . import registration, slam  # noqa: F401
