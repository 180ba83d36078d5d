"""Front-ends of the hot path: ONNX weights -> nn.Module without the `onnx`
package, VNNLIB -> (input box, [(C, rhs)]) objectives."""
from .onnx_reader import load_onnx, parse_onnx          # noqa: F401
from .vnnlib import read_vnnlib                          # noqa: F401
