"""Drop-in `import i8ie` for scripts written against t0037799/INT8InferenceEngine:
the same surface (i8ie/__init__.py:6-10 of the reference), served by the B200 backend."""
from int8inferenceengine_b200.api import (Conv2d, Layer, Linear, Module, Tensor, argmax, dequantize,  # noqa: F401
                                         max_pool2d, quantize, relu, tensor)

__all__ = ["tensor", "argmax", "relu", "max_pool2d", "Linear", "Conv2d", "Tensor", "quantize", "dequantize"]
