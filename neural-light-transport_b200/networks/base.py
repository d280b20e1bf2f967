"""Root of the network classes.  Interface of nlt/networks/base.py:26-40: a `.layers` list, call-ability, and the
`str2none` helper the INI-driven constructors use (ConfigParser has no notion of None, so 'None' arrives as text)."""


class Network:
    def __init__(self):
        self.layers = []

    def __call__(self, x):
        raise NotImplementedError

    @staticmethod
    def str2none(str_):
        """'none' in any capitalisation -> None, every other string unchanged; non-strings are a caller bug."""
        assert isinstance(str_, str), "Call this only on strings"
        return None if str_.lower() == 'none' else str_
