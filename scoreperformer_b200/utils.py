"""Small helpers the hot path imports (reference: scoreperformer/utils/functions.py:12-34, 74-88)."""
from enum import Enum
from inspect import isfunction


def exists(val):
    return val is not None


def default(val, d):
    if exists(val):
        return val
    return d() if isfunction(d) else d


class equals:
    def __init__(self, val):
        self.val = val

    def __call__(self, x, *args, **kwargs):
        return x == self.val


def or_reduce(masks):
    head, *body = masks
    for rest in body:
        head = head | rest
    return head


class ExplicitEnum(str, Enum):
    """Enum with an explicit error message for missing values."""

    @classmethod
    def _missing_(cls, value):
        raise ValueError(f"{value} is not a valid {cls.__name__}, please select one of {list(cls._value2member_map_.keys())}")

    @classmethod
    def has_value(cls, value):
        return value in cls._value2member_map_

    @classmethod
    def list(cls):
        return list(map(lambda c: c.value, cls))
