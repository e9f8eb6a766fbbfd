from ramannoodle.dynamics._trajectory import Trajectory  # noqa: F401
