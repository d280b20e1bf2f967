"""Dataset registry (reference: nlt/datasets/__init__.py:15-20): `datasets.get_dataset_class(name)` imports
`datasets.<name>` and returns its `Dataset` class -- same lookup protocol as `models.get_model_class`."""
from importlib import import_module


def get_dataset_class(name):
    return import_module('datasets.' + name).Dataset
