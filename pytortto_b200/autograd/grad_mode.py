"""Global gradient mode: `no_grad`, `enable_grad`, `set_grad_enabled`, `is_grad_enabled`.
Same contract as the reference's autograd/grad_mode.py:6-109 (a process-wide flag honoured by build_links,
helper.py:45-50); usable as context managers and as decorators."""
import functools

_grad_enabled = True


def is_grad_enabled():
    return _grad_enabled


class set_grad_enabled:
    def __init__(self, mode):
        global _grad_enabled
        self.prev = _grad_enabled
        _grad_enabled = bool(mode)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        global _grad_enabled
        _grad_enabled = self.prev
        return False


class _ModeContext:
    mode = True

    def __enter__(self):
        global _grad_enabled
        self.prev = _grad_enabled
        _grad_enabled = self.mode
        return self

    def __exit__(self, *exc):
        global _grad_enabled
        _grad_enabled = self.prev
        return False

    def __call__(self, fn):
        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            with self.__class__():
                return fn(*args, **kwargs)
        return wrapped


class no_grad(_ModeContext):
    mode = False


class enable_grad(_ModeContext):
    mode = True
