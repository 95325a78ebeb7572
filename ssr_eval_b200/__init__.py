"""ssr_eval_b200 -- B200-native (sm_100a CUDA) implementation of haoheliu/ssr_eval's DSP hot path
behind the reference's own Python API (ssr_eval/__init__.py:1-2)."""
from .eval import SSR_Eval_Helper, BasicTestee  # noqa: F401
from .test import test  # noqa: F401
from .metrics import AudioMetrics  # noqa: F401
from .lowpass import lowpass  # noqa: F401

__all__ = ["SSR_Eval_Helper", "BasicTestee", "test", "AudioMetrics", "lowpass"]
