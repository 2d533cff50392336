"""Exception types of the reference API (src/bayesbay/exceptions/_exceptions.py:1-57).

The device path cannot raise from inside a forward evaluation, so ``ForwardException`` /
``UserFunctionException`` are kept for API compatibility only; non-finite likelihood ratios are
rejected on the device instead."""


class UserFunctionException(Exception):
    """An exception raised from a user-provided function."""

    def __init__(self, original_exc=None):
        self.original_exc = original_exc
        super().__init__(str(original_exc))


class ForwardException(UserFunctionException):
    """An exception raised from a user-provided forward function."""


class InitException(Exception):
    def __init__(self, name=None, **kwargs):
        self.name = name
        super().__init__(f"could not initialise {name}" if name else "initialisation failed")


class OutOfDomainException(Exception):
    def __init__(self, name=None, position=None):
        super().__init__(f"position {position} is outside the domain of position-dependent hyper-parameter {name}")


class DeviceUnsupported(TypeError):
    """The requested feature has no device implementation (and there is no CPU fallback)."""
