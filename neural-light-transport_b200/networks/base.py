"""Mirror of nlt/networks/base.py:26-40."""


class Network:
    def __init__(self):
        self.layers = []

    def __call__(self, x):
        raise NotImplementedError

    @staticmethod
    def str2none(str_):
        """Mostly to overcome there being no `config.getnone()` method
        (reference: nlt/networks/base.py:33-40)."""
        assert isinstance(str_, str), "Call this only on strings"
        if str_.lower() == 'none':
            return None
        return str_
