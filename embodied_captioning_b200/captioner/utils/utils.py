"""Captioner configuration holders.

Drop-in for the two classes of experimenting_env/captioner/utils/utils.py:2-12 (same constructor keywords, same
attribute names: ``Configuration(...).captioner.arch_name`` etc.), written as dataclasses so that they also compare,
print and copy sensibly.
"""
from dataclasses import dataclass, field
from typing import Optional

_FIELDS = ("arch_name", "model_name", "checkpoint_name", "height", "width")


@dataclass
class CaptionerField:
    arch_name: Optional[str] = None
    model_name: Optional[str] = None
    checkpoint_name: Optional[str] = None
    height: Optional[int] = None
    width: Optional[int] = None


@dataclass(init=False)
class Configuration:
    captioner: CaptionerField = field(default_factory=CaptionerField)

    def __init__(self, *args, **kwargs):
        # positional order and keywords follow the reference signature
        if len(args) > len(_FIELDS):
            raise TypeError(f"Configuration takes at most {len(_FIELDS)} positional arguments")
        values = dict(zip(_FIELDS, args))
        for name, value in kwargs.items():
            if name not in _FIELDS:
                raise TypeError(f"Configuration got an unexpected keyword argument {name!r}")
            if name in values:
                raise TypeError(f"Configuration got multiple values for argument {name!r}")
            values[name] = value
        self.captioner = CaptionerField(**values)
