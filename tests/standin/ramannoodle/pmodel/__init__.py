from ramannoodle.pmodel._art import ARTModel  # noqa: F401
from ramannoodle.pmodel._interpolation import InterpolationModel  # noqa: F401
