"""Exception / warning types raised at the API boundary.

Same names and meaning as the reference's bayesml/_exceptions.py (:3-25) so that user code catching
`ParameterFormatError`, `DataFormatError`, `CriteriaError` or filtering `ResultWarning` keeps working.
"""


class _MessageError(Exception):
    def __init__(self, value):
        super().__init__(value)
        self.value = value

    def __str__(self):
        return repr(self.value)


class ParameterFormatError(_MessageError):
    """A constant or hyperparameter has the wrong type, shape or range."""


class DataFormatError(_MessageError):
    """A data array has the wrong type or shape."""


class CriteriaError(_MessageError):
    """An unsupported loss / criterion was requested."""


class ResultWarning(UserWarning):
    """The result may not be what the caller expects (e.g. VB did not converge)."""


class ParameterFormatWarning(UserWarning):
    pass
