"""irspack_b200: the iALS training + scoring hot path of tohtsky/irspack,
rebuilt for NVIDIA B200 (sm_100a) behind the reference's own interfaces.

    from irspack_b200 import IALSRecommender, Evaluator
    rec = IALSRecommender(X, n_components=128, alpha0=0.1, reg=1e-3).learn()
    Evaluator(X_test, cutoff=10).get_score(rec)

The first attribute access loads ``lib/libials_b200.so`` (build it with
``python -m irspack_b200.build``); there is no CPU fallback, a missing library
is a hard error.  (Attributes are resolved lazily only so that the build module
itself can be imported before the library exists.)
"""
import importlib
from typing import Any

_EXPORTS = {
    "IALSRecommender": "ials", "IALSConfig": "ials", "IALSTrainer": "ials",
    "Evaluator": "evaluation", "EvaluatorWithColdUser": "evaluation", "Metrics": "evaluation",
    "topk_scores": "evaluation", "select_topk": "evaluation",
    "IALSModelConfig": "_ials_core", "IALSModelConfigBuilder": "_ials_core",
    "IALSSolverConfig": "_ials_core", "IALSSolverConfigBuilder": "_ials_core",
    "LossType": "_ials_core", "SolverType": "_ials_core",
    "device_count": "_lib", "version": "_lib",
    "IDMapper": "id_mapping", "ItemIDMapper": "id_mapping",
    "retrieve_recommend_from_score": "id_mapping",
}
_SUBMODULES = {"_ials_core", "_lib", "ials", "evaluation", "synth", "dist", "build", "_threading", "id_mapping", "ops"}

__all__ = sorted(_EXPORTS) + ["_ials_core"]


def __getattr__(name: str) -> Any:
    if name in _EXPORTS:
        return getattr(importlib.import_module(f".{_EXPORTS[name]}", __name__), name)
    if name in _SUBMODULES:
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
