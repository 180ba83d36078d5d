"""Front-ends: the VNNLIB reader against outputs of the reference's own reader
(tests/golden/frontend/vnnlib_expected.pt, made by oracle/gen_frontend_golden.py), and the protobuf-level ONNX reader on
the ACAS Xu 1_1 network (BASELINE.json configs[0])."""
import glob
import os

import numpy as np
import torch

from neuralsat_b200.frontend import onnx_reader, vnnlib

D = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'frontend')


def test_vnnlib_matches_the_reference_reader():
    exp = torch.load(os.path.join(D, 'vnnlib_expected.pt'), weights_only=False)
    files = sorted(glob.glob(os.path.join(D, '*.vnnlib')))
    assert len(files) >= 5
    for f in files:
        got = vnnlib.read_vnnlib(f)
        ref = exp[os.path.basename(f)]
        assert len(got) == len(ref), f
        for (box, specs), (rbox, rspecs) in zip(got, ref):
            assert np.array_equal(np.array(box), np.array(rbox)), f
            assert len(specs) == len(rspecs)
            for (m, r), (rm, rr) in zip(specs, rspecs):
                assert np.array_equal(m, np.array(rm, dtype=float)) and np.array_equal(r, np.array(rr, dtype=float)), f


def test_vnnlib_is_layout_independent(tmp_path):
    """Statements split over lines / extra blanks / comments parse to the same objectives."""
    src = open(os.path.join(D, 'prop_6.vnnlib')).read()
    mangled = src.replace('(assert', '(assert\n   ').replace(' (and', '\n (and').replace('(<=', '( <=  ') + '\n; trailing comment\n'
    p = tmp_path / 'm.vnnlib'
    p.write_text(mangled)
    a, b = vnnlib.read_vnnlib(os.path.join(D, 'prop_6.vnnlib')), vnnlib.read_vnnlib(str(p))
    assert repr(a) == repr(b)


def test_objectives_layout():
    obj = vnnlib.objectives(vnnlib.read_vnnlib(os.path.join(D, 'prop_1.vnnlib')))
    assert obj.lower_bounds.shape == (1, 5) and obj.cs.shape == (1, 1, 5) and obj.rhs.shape == (1, 1)
    assert torch.all(obj.lower_bounds <= obj.upper_bounds)
    # ACAS Xu property 1: unsafe iff Y_0 >= 3.9911 -> row -e_0, rhs -3.9911
    assert obj.cs[0, 0].tolist() == [-1.0, 0.0, 0.0, 0.0, 0.0] and abs(float(obj.rhs[0, 0]) + 3.991125645861615) < 1e-6


def test_onnx_reader_acasxu():
    """13 310 parameters, Sub(mean) -> Flatten -> 6 x (MatMul 50 + Add + Relu) -> MatMul 5 + Add."""
    path = os.path.join(D, 'ACASXU_run2a_1_1_batch_2000.onnx')
    g = onnx_reader.load_onnx(path)
    ops = [n['op'] for n in g['nodes']]
    assert ops.count('MatMul') == 7 and ops.count('Relu') == 6 and ops.count('Sub') == 1
    model, in_shape, out_shape, is_nhwc = onnx_reader.parse_onnx(path)
    assert in_shape == (1, 1, 1, 5) and out_shape == (1, 5) and not is_nhwc
    assert sum(p.numel() for p in model.parameters()) == 13305         # + the 5 input means held as a buffer = 13 310
    # the converted module computes what the ONNX graph says: evaluate the graph by hand in float64
    x = torch.rand(3, 1, 1, 5)
    init = {k: torch.from_numpy(np.array(v, dtype=np.float64)) for k, v in g['init'].items()}
    env = {g['inputs'][0][0]: x.double()}
    for n in g['nodes']:
        a = [env.get(i, init.get(i)) for i in n['input']]
        if n['op'] == 'Sub':
            v = a[0] - a[1]
        elif n['op'] == 'Flatten':
            v = a[0].flatten(1)
        elif n['op'] == 'MatMul':
            v = a[0] @ a[1]
        elif n['op'] == 'Add':
            v = a[0] + a[1]
        elif n['op'] == 'Relu':
            v = a[0].clamp(min=0)
        env[n['output'][0]] = v
    ref = env[g['outputs'][0][0]]
    assert torch.allclose(model(x).double(), ref, rtol=1e-5, atol=1e-5)
    # and it traces into the node list of the bounding path: 7 Linear, 6 ReLU, the mean as an unperturbed operand
    from neuralsat_b200.graph import trace_module
    nodes = trace_module(model, in_shape)
    kinds = [nd['op'] for nd in nodes]
    assert kinds.count('linear') == 7 and kinds.count('relu') == 6 and kinds.count('addconst') == 1
