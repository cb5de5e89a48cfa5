"""Operator plugin base class and loader -- same contract as reference ``localization/base_model.py``:
``BaseModel(conf)`` merges ``default_conf``, calls ``_init``; ``forward`` checks ``required_data_keys``
and calls ``_forward``; ``dynamic_load(root, name)`` returns the single BaseModel subclass defined in
``root.<name>`` (reference base_model.py:7-44)."""
import inspect
from abc import ABCMeta, abstractmethod
from copy import copy

from torch import nn


class BaseModel(nn.Module, metaclass=ABCMeta):
    default_conf = {}
    required_data_keys = []

    def __init__(self, conf):
        super().__init__()
        self.conf = conf = {**self.default_conf, **conf}
        self.required_data_keys = copy(self.required_data_keys)
        self._init(conf)

    def forward(self, data):
        for key in self.required_data_keys:
            assert key in data, 'Missing key {} in data'.format(key)
        return self._forward(data)

    @abstractmethod
    def _init(self, conf):
        raise NotImplementedError

    @abstractmethod
    def _forward(self, data):
        raise NotImplementedError


def dynamic_load(root, model):
    module_path = f'{root.__name__}.{model}'
    module = __import__(module_path, fromlist=[''])
    found = [c for _, c in inspect.getmembers(module, inspect.isclass)
             if c.__module__ == module_path and issubclass(c, BaseModel)]
    assert len(found) == 1, found
    return found[0]
