"""Model registry: `models.get_model_class('nlt')` -> the `Model` class of module `models.nlt`
(lookup protocol of nlt/models/__init__.py:15-20, used by trainvali.py:116-118 and nlt_test.py:62-64)."""
import importlib


def get_model_class(name):
    return getattr(importlib.import_module('%s.%s' % (__name__, name)), 'Model')
