"""Drop-in for /root/reference/modules/utils.py (B200 engine overlay, opt-in with `install_overlay(train=True)`).

`evaluate_vae_reconstruction` (utils.py:428-561) is replaced by the batched sweep of
simulgen_vae_b200.export; every other name of the reference module (Dataset, LatentConditionerDataset,
parse_condition_file, setup_distributed_training, ...) is re-exported unchanged from the reference file found further
down the `modules` namespace-package path."""
import importlib.util
import os
import sys

import modules as _pkg

_here = os.path.dirname(os.path.abspath(__file__))
for _d in list(getattr(_pkg, "__path__", [])):
    _f = os.path.join(_d, "utils.py")
    if os.path.abspath(_d) != _here and os.path.isfile(_f):
        _spec = importlib.util.spec_from_file_location("modules._reference_utils", _f)
        _ref = importlib.util.module_from_spec(_spec)
        sys.modules["modules._reference_utils"] = _ref        # classes defined there stay picklable
        _spec.loader.exec_module(_ref)
        globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})
        reference_evaluate_vae_reconstruction = _ref.evaluate_vae_reconstruction
        break

from simulgen_vae_b200.export import evaluate_vae_reconstruction, export_latents  # noqa: E402,F401
