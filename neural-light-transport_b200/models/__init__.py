"""Model registry -- mirror of nlt/models/__init__.py:15-20."""
from importlib import import_module


def get_model_class(name):
    mod = import_module('models.' + name)
    return mod.Model
