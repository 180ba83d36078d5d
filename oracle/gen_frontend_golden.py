"""Golden vectors for the front-ends (tests/golden/frontend/): the UNMODIFIED reference's VNNLIB reader
(NS/util/spec/read_vnnlib.py) run on the copied benchmark specifications, results stored next to them.

TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference).
    python oracle/gen_frontend_golden.py
The .vnnlib / .onnx files in that directory are benchmark DATA copied from the reference's neuralbench submodule
(ACAS Xu properties 1, 2, 3, 6, 7 and network 1_1) and NS/example/vnnlib/motivation_example.vnnlib."""
import glob
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_bootstrap  # noqa: E402


def main():
    ref_bootstrap.bootstrap()
    from util.spec.read_vnnlib import read_vnnlib
    d = os.path.join(ROOT, 'tests', 'golden', 'frontend')
    out = {}
    for f in sorted(glob.glob(os.path.join(d, '*.vnnlib'))):
        res = read_vnnlib(f)
        out[os.path.basename(f)] = [([list(map(float, b)) for b in box], [(m.tolist(), r.tolist()) for m, r in specs])
                                    for box, specs in res]
        print(os.path.basename(f), len(res), [len(s) for _, s in res])
    torch.save(out, os.path.join(d, 'vnnlib_expected.pt'))


if __name__ == '__main__':
    main()
