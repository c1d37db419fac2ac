"""parla/utils/misc.py:3-9."""


def set_docstring(docstr):
    def decorator(func):
        func.__doc__ = docstr
        return func
    return decorator
