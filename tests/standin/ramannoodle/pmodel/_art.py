from ramannoodle.pmodel._interpolation import InterpolationModel


class ARTModel(InterpolationModel):
    """ramannoodle/pmodel/_art.py:48 — inherits calc_polarizabilities."""
