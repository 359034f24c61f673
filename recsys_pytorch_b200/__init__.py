"""recsys_pytorch_b200 - B200 (sm_100a) BPR-MF training + scoring engine behind the
plugin surface of yoongi0428/RecSys_PyTorch (models/BaseModel.py, models/MF.py,
data/generators.py::PairwiseGenerator, evaluation/).  See DESIGN.md.

The compute lives in libb200rec.so (hand-written CUDA, C ABI in include/b200rec.h);
this package is the host-side mirror of the reference's Python interface.
"""
from . import _lib  # noqa: F401
from ._lib import B200RecError, LIB_PATH  # noqa: F401

__all__ = ["MF", "B200MF", "BaseModel", "LightGCN", "NGCF", "PairwiseGenerator", "Evaluator", "UIRTDataset", "engine", "B200RecError"]


def __getattr__(name):  # lazy: importing the package must not require torch.cuda
    if name in ("MF", "B200MF", "BaseModel", "EmbeddingTable"):
        from . import mf
        return mf.MF if name == "B200MF" else getattr(mf, name)
    if name == "LightGCN":
        from . import lightgcn
        return lightgcn.LightGCN
    if name == "NGCF":
        from . import ngcf
        return ngcf.NGCF
    if name == "PairwiseGenerator":
        from .generators import PairwiseGenerator
        return PairwiseGenerator
    if name in ("Evaluator", "predict_topk_func", "eval_func_router", "Statistics"):
        from . import evaluation
        return getattr(evaluation, name)
    if name == "UIRTDataset":
        from .dataset import UIRTDataset
        return UIRTDataset
    if name in ("engine", "synthetic", "dist", "evaluation", "generators", "mf", "lightgcn", "ngcf", "p2p", "dataset"):
        import importlib
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)
