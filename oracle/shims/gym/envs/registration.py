import importlib

registry = {}


def register(id, entry_point=None, **kwargs):
    registry[id] = (entry_point, kwargs)


def make(id, **kwargs):
    entry_point, kw = registry[id]
    if callable(entry_point):
        return entry_point(**kw, **kwargs)
    mod_name, attr = entry_point.split(":")
    mod = importlib.import_module(mod_name)
    return getattr(mod, attr)(**kw, **kwargs)
