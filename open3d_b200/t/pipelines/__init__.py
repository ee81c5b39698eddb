from . import odometry, registration, slam  # noqa: F401
