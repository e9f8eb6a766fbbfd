class UserError(Exception):
    """ramannoodle/exceptions.py"""
