"""Theoretical variogram model container accepted by the kriging calls.

The kriging functions take any object with ``nugget``, ``sill`` (partial sill), ``rang``,
``model_type`` (or ``name``) and ``direction`` attributes — i.e. the reference's own
``TheoreticalVariogram`` (semivariogram/theoretical/classes/theoretical_variogram.py) works
unchanged.  This class is the minimal stand-in for when the reference is not installed:
``from_dict`` / ``to_dict`` use the reference's keys (:661-702) and ``predict`` evaluates gamma(h)
with the device implementation of variogram_models/models.py:41-440.  Model fitting
(``fit`` / ``autofit``) is outside the hot path and stays in the reference.
"""
import numpy as np

from . import _lib

ALL_MODELS = ['circular', 'cubic', 'exponential', 'gaussian', 'linear', 'power', 'spherical']


def model_id(name):
    """variogram_models/functions.py:130-133: unknown names raise KeyError."""
    if name not in _lib.MODEL_IDS:
        raise KeyError(f'Defined model name {name} is not implemented. '
                       f'You may choose one from {ALL_MODELS} instead.')
    return _lib.MODEL_IDS[name]


def model_parameters(theoretical_model):
    """-> (model id, nugget, sill, rang, direction) from a duck-typed model object or dict."""
    if isinstance(theoretical_model, dict):
        d = theoretical_model
        name = d.get("variogram_model_type", d.get("model_type", d.get("name")))
        return (model_id(name), float(d["nugget"]), float(d["sill"]), float(d["rang"]), d.get("direction"))
    name = getattr(theoretical_model, "model_type", None) or getattr(theoretical_model, "name", None)
    return (model_id(name), float(theoretical_model.nugget), float(theoretical_model.sill),
            float(theoretical_model.rang), getattr(theoretical_model, "direction", None))


class TheoreticalVariogram:
    def __init__(self, model_params=None):
        self.model_type = None
        self.nugget = 0.0
        self.sill = 0.0
        self.rang = 0.0
        self.direction = None
        if model_params is not None:
            self.from_dict(model_params)

    @property
    def name(self):
        return self.model_type

    def from_dict(self, parameters):
        name = parameters.get("variogram_model_type", parameters.get("model_type"))
        model_id(name)
        self.model_type = name
        self.nugget = parameters["nugget"]
        self.sill = parameters["sill"]
        self.rang = parameters["rang"]
        self.direction = parameters.get("direction")

    def to_dict(self):
        if self.model_type is None:
            raise AttributeError('Model is not set yet, cannot export model parameters to dict')
        return dict(nugget=self.nugget, sill=self.sill, rang=self.rang,
                    variogram_model_type=self.model_type, direction=self.direction)

    def predict(self, distances):
        """gamma(h) (classes/theoretical_variogram.py:497-524), evaluated on the GPU."""
        distances = np.asarray(distances, dtype=np.float64)
        out = _lib.gamma_host(model_id(self.model_type), self.nugget, self.sill, self.rang, distances.ravel())
        return out.reshape(distances.shape)
