"""Class-attribute configs with the reference's semantics (src/config.py:45-160): a ``Config`` nested class lists the
fields, ``Configurable.__init__`` copies them onto ``self``; nested configs and ``update(dict)`` are supported so the
reference's JSON files and ``-s a.b value`` overrides keep working."""
import copy

SIMPLE_TYPES = (bool, int, float, str)


class Optional:
    def __init__(self, dtype):
        self.dtype = dtype


class BaseConfig:
    def vars(self):
        out = {}
        for key in dir(self):
            if key.startswith("_"):
                continue
            val = getattr(self, key)
            if callable(val):
                continue
            out[key] = val
        return out

    def __init__(self, **kwargs):
        v = self.vars()
        v.update(kwargs)
        for key, val in v.items():
            setattr(self, key, copy.deepcopy(val) if isinstance(val, BaseConfig) else val)

    def update(self, d):
        for key, val in d.items():
            assert hasattr(self, key), f"Cannot set non-existent key {key} in {type(self).__name__}"
            cur = getattr(self, key)
            if isinstance(val, dict) and isinstance(cur, BaseConfig):
                cur.update(val)
            else:
                if isinstance(cur, float) and isinstance(val, int) and not isinstance(val, bool):
                    val = float(val)
                setattr(self, key, val)
        return self

    def nested_set(self, path, value):
        obj = self
        for p in path[:-1]:
            obj = getattr(obj, p)
        assert hasattr(obj, path[-1]), f"Cannot override non-existent key {'.'.join(path)}"
        setattr(obj, path[-1], value)

    def verify(self):
        for key, val in self.vars().items():
            if isinstance(val, BaseConfig):
                val.verify()
            elif isinstance(val, Optional):
                setattr(self, key, None)


def adopt_config(own_cls, foreign):
    """An own ``Config`` carrying the field values of a foreign config object that follows the same protocol (the reference's
    src/config.py classes: ``vars()``, nested configs, ``Optional`` placeholders) - how the reference's driver hands its
    ``model_cfg`` / ``sac_cfg`` to the drop-in classes.  Unknown fields are an error, exactly like ``update``."""
    own = own_cls()
    for key, val in foreign.vars().items():
        assert hasattr(own, key), f"{own_cls.__qualname__} has no field {key!r} (got it from {type(foreign).__qualname__})"
        cur = getattr(own, key)
        if isinstance(cur, BaseConfig) and hasattr(val, "vars"):
            setattr(own, key, adopt_config(type(cur), val))
        elif type(val).__name__ == "Optional" and hasattr(val, "dtype"):
            setattr(own, key, Optional(val.dtype))
        else:
            setattr(own, key, copy.deepcopy(val))
    return own


class Configurable:
    """Subclasses define a nested ``Config``; its fields become attributes of the instance."""

    def __init__(self, config):
        if type(config) is not self.__class__.Config:
            assert hasattr(config, "vars") and type(config).__name__ == "Config", f"expected {self.__class__.Config}, got {type(config)}"
            config = adopt_config(self.__class__.Config, config)
        self.config = copy.deepcopy(config)
        self.config.verify()
        for key, val in self.config.vars().items():
            setattr(self, key, val)
