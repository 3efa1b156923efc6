"""Drop-in for /root/reference/modules/data_preprocess.py (B200 engine overlay).

`data_scaler` (data_preprocess.py:65-165) is replaced by the GPU scan of simulgen_vae_b200.preprocess (same
signature, same return values, same ./model_save/scaler.pkl, bit-identical numbers); every other name of the
reference module (`reduce_dataset`, `latent_conditioner_scaler`, ...) is re-exported unchanged from the reference
file found further down the `modules` namespace-package path, so `from modules.data_preprocess import reduce_dataset,
data_scaler, latent_conditioner_scaler` (SimulGen-VAE.py:71) keeps working."""
import importlib.util
import os

import modules as _pkg

_here = os.path.dirname(os.path.abspath(__file__))
for _d in list(getattr(_pkg, "__path__", [])):
    _f = os.path.join(_d, "data_preprocess.py")
    if os.path.abspath(_d) != _here and os.path.isfile(_f):
        _spec = importlib.util.spec_from_file_location("modules._reference_data_preprocess", _f)
        _ref = importlib.util.module_from_spec(_spec)
        _spec.loader.exec_module(_ref)
        globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})
        reference_data_scaler = _ref.data_scaler
        break

from simulgen_vae_b200.preprocess import data_scaler, data_scaler_to_device  # noqa: E402,F401
