"""Abstract contracts shared by the learn models (the subset of the reference's bayesml/base.py that the
gaussianmixture hot path touches: `Posterior` :139-288 and `PredictiveMixin` :290-356)."""
import pickle
from abc import ABCMeta, abstractmethod

from ._exceptions import ParameterFormatError

_PICKLE_HINT = (" must be a pickled python dictionary obtained by ``GenModel.save_h_params()``, "
                "``LearnModel.save_h0_params()`` or ``LearnModel.save_hn_params()``.")


def _dump(obj, filename):
    with open(filename, "wb") as f:
        pickle.dump(obj, f)


def _load_dict(filename):
    with open(filename, "rb") as f:
        obj = pickle.load(f)
    if type(obj) is not dict:
        raise ParameterFormatError(filename + _PICKLE_HINT)
    return obj


class Generative(metaclass=ABCMeta):
    """Data generative model: parameters, their prior hyperparameters h_*, sampling (reference base.py:8-137)."""

    @abstractmethod
    def set_h_params(self): ...

    @abstractmethod
    def get_h_params(self): ...

    @abstractmethod
    def gen_params(self): ...

    @abstractmethod
    def set_params(self): ...

    @abstractmethod
    def get_params(self): ...

    @abstractmethod
    def gen_sample(self): ...

    @abstractmethod
    def save_sample(self): ...

    def save_h_params(self, filename):
        """Pickle the dict returned by get_h_params() (reference base.py:17-31)."""
        _dump(self.get_h_params(), filename)

    def load_h_params(self, filename):
        """Positional load of a pickled hyperparameter dict into set_h_params (reference base.py:33-57)."""
        self.set_h_params(*_load_dict(filename).values())
        return self

    def save_params(self, filename):
        """Pickle the dict returned by get_params() (reference base.py:68-83)."""
        _dump(self.get_params(), filename)

    def load_params(self, filename):
        """Positional load of a pickled parameter dict into set_params (reference base.py:85-103)."""
        with open(filename, "rb") as f:
            obj = pickle.load(f)
        if type(obj) is not dict:
            raise ParameterFormatError(filename + " must be a pickled python dictionary obtained by ``GenModel.save_params()``")
        self.set_params(*obj.values())
        return self


class Posterior(metaclass=ABCMeta):
    """Posterior over the parameters: h0_* (initial) and hn_* (updated) hyperparameters."""

    @abstractmethod
    def set_h0_params(self): ...

    @abstractmethod
    def get_h0_params(self): ...

    @abstractmethod
    def set_hn_params(self): ...

    @abstractmethod
    def get_hn_params(self): ...

    @abstractmethod
    def update_posterior(self): ...

    @abstractmethod
    def estimate_params(self): ...

    @abstractmethod
    def visualize_posterior(self): ...

    def save_h0_params(self, filename):
        """Pickle the dict returned by get_h0_params() (reference base.py:148-167)."""
        _dump(self.get_h0_params(), filename)

    def load_h0_params(self, filename):
        """Positional load of a pickled hyperparameter dict into set_h0_params (reference base.py:169-198)."""
        self.set_h0_params(*_load_dict(filename).values())
        return self

    def save_hn_params(self, filename):
        """Pickle the dict returned by get_hn_params() (reference base.py:208-227)."""
        _dump(self.get_hn_params(), filename)

    def load_hn_params(self, filename):
        """Positional load of a pickled hyperparameter dict into set_hn_params (reference base.py:229-258)."""
        self.set_hn_params(*_load_dict(filename).values())
        return self

    def reset_hn_params(self):
        """hn_* <- h0_* (and the predictive parameters that follow from them; reference base.py:260-267)."""
        self.set_hn_params(*self.get_h0_params().values())
        return self

    def overwrite_h0_params(self):
        """h0_* <- hn_* (reference base.py:269-276)."""
        self.set_h0_params(*self.get_hn_params().values())
        return self


class PredictiveMixin(metaclass=ABCMeta):
    """Predictive distribution interface (reference base.py:290-356)."""

    @abstractmethod
    def get_p_params(self): ...

    @abstractmethod
    def calc_pred_dist(self): ...

    @abstractmethod
    def make_prediction(self): ...

    @abstractmethod
    def pred_and_update(self): ...
