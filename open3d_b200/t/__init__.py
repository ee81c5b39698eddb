from . import geometry, pipelines  # noqa: F401
