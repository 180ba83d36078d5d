"""VNNLIB -> objectives for the bounding path.

Same result convention as the reference's reader (NS/util/spec/read_vnnlib.py:134-313):
    read_vnnlib(path) -> [(box, [(mat, rhs), ...]), ...]
`box` = [[lo, hi] per input], every `(mat, rhs)` one conjunct of the (single) output disjunction, describing the
COUNTER-EXAMPLE region `mat @ y <= rhs`; entries with the same input box are merged.  A conjunct is refuted - the
property holds on the box - as soon as the lower bound of ONE of its rows exceeds its rhs
(`stop_criterion_batch_any`, AL/utils.py:87-93), which is why `objectives()` hands `mat` to the bounding path as the
spec matrix C and `rhs` as the decision threshold (NS/verifier/objective.py:31-48, :99-129).

Unlike the reference (regular expressions over normalised lines) this is a small s-expression reader: statements are
tokenised into nested lists and interpreted, so line breaks and spacing do not matter.  Supported, as in the reference:
`declare-const`, `(assert (<=|>= a b))` on inputs (box constraints) or outputs (added to every conjunct), and
`(assert (or (and c...) ...))` whose comparisons may mix input and output constraints.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import List

import numpy as np
import torch


def _tokens(text: str):
    for line in text.splitlines():
        line = line.split(';', 1)[0]
        for tok in line.replace('(', ' ( ').replace(')', ' ) ').split():
            yield tok


def _parse(text: str) -> list:
    """All top-level s-expressions of the file as nested lists."""
    stack, out = [], []
    for tok in _tokens(text):
        if tok == '(':
            stack.append([])
        elif tok == ')':
            if not stack:
                raise ValueError('mismatched parenthesis in vnnlib file')
            done = stack.pop()
            (stack[-1] if stack else out).append(done)
        else:
            if not stack:
                raise ValueError(f'stray token {tok!r} outside any statement')
            stack[-1].append(tok)
    if stack:
        raise ValueError('mismatched parenthesis in vnnlib file')
    return out


class _Case:
    """One (input box, conjunct under construction)."""

    def __init__(self, n_in):
        self.box = [[-math.inf, math.inf] for _ in range(n_in)]
        self.mat, self.rhs = [], []

    def copy(self):
        c = _Case(0)
        c.box = [list(b) for b in self.box]
        c.mat = [list(r) for r in self.mat]
        c.rhs = list(self.rhs)
        return c

    def constrain(self, op, first, second, n_out):
        """`(op first second)`, NS/util/spec/read_vnnlib.py:74-122."""
        if first.startswith('X_'):
            if second[:1] in ('X', 'Y'):
                raise ValueError(f'input constraints must be box ({op} {first} {second})')
            lim = self.box[int(first[2:])]
            if op == '<=':
                lim[1] = min(float(second), lim[1])
            else:
                lim[0] = max(float(second), lim[0])
            if lim[0] > lim[1]:
                raise ValueError(f'{first} range is empty: {lim}')
            return
        if op == '>=':                                  # a >= b  ==  b <= a
            first, second = second, first
        row, rhs = [0.0] * n_out, 0.0
        if first.startswith('Y_') and second.startswith('Y_'):
            row[int(first[2:])] = 1.0
            row[int(second[2:])] = -1.0
        elif first.startswith('Y_'):
            row[int(first[2:])] = 1.0
            rhs = float(second)
        elif second.startswith('Y_'):
            row[int(second[2:])] = -1.0
            rhs = -float(first)
        else:
            raise ValueError(f'constraint without a variable: ({op} {first} {second})')
        self.mat.append(row)
        self.rhs.append(rhs)


def _is_cmp(e):
    return isinstance(e, list) and len(e) == 3 and e[0] in ('<=', '>=') and all(isinstance(t, str) for t in e[1:])


def read_vnnlib(path: str) -> list:
    with open(path, 'r') as f:
        stmts = _parse(f.read())
    n_in = n_out = 0

    def scan(e):
        nonlocal n_in, n_out
        if isinstance(e, list):
            for t in e:
                scan(t)
        elif e.startswith('X_') and e[2:].isdigit():
            n_in = max(n_in, int(e[2:]) + 1)
        elif e.startswith('Y_') and e[2:].isdigit():
            n_out = max(n_out, int(e[2:]) + 1)

    scan(stmts)
    cases = [_Case(n_in)]
    for st in stmts:
        if not st or st[0] == 'declare-const':
            continue
        if st[0] != 'assert' or len(st) != 2:
            continue                                    # the reference skips what it cannot read, too
        body = st[1]
        if _is_cmp(body):
            for c in cases:
                c.constrain(body[0], body[1], body[2], n_out)
        elif isinstance(body, list) and body and body[0] == 'or':
            new = []
            for c in cases:
                for conj in body[1:]:
                    cmps = [conj] if _is_cmp(conj) else (conj[1:] if conj and conj[0] == 'and' else None)
                    if cmps is None or not all(_is_cmp(t) for t in cmps):
                        raise ValueError(f'unsupported disjunct: {conj}')
                    cc = c.copy()
                    for t in cmps:
                        cc.constrain(t[0], t[1], t[2], n_out)
                    new.append(cc)
            cases = new
        elif isinstance(body, list) and body and body[0] == 'and' and all(_is_cmp(t) for t in body[1:]):
            for c in cases:
                for t in body[1:]:
                    c.constrain(t[0], t[1], t[2], n_out)
    merged = {}
    for c in cases:                                     # same input box -> one entry, its conjuncts as a list
        key = repr(c.box)
        merged.setdefault(key, (c.box, []))[1].append((np.array(c.mat, dtype=float), np.array(c.rhs, dtype=float)))
    out = []
    for box, specs in merged.values():
        for d, r in enumerate(box):
            if r[0] == -math.inf or r[1] == math.inf:
                raise ValueError(f'input X_{d} was unbounded: {r}')
        out.append((box, specs))
    return out


def objectives(vnnlib: list, dtype=torch.float32) -> SimpleNamespace:
    """What `DnfObjectives.pop(all)` hands to `NetworkAbstractor.initialize/forward` (NS/verifier/objective.py:77-129):
    one entry per (input box, conjunct): lower_bounds / upper_bounds [N, n_in], cs [N, S, n_out], rhs [N, S], ids [N].
    Conjuncts must have the same number of rows S (true for every benchmark of BASELINE.json)."""
    lo, hi, cs, rhs = [], [], [], []
    for box, specs in vnnlib:
        b = torch.tensor(box, dtype=dtype)
        for mat, r in specs:
            lo.append(b[:, 0])
            hi.append(b[:, 1])
            cs.append(torch.tensor(mat, dtype=dtype))
            rhs.append(torch.tensor(r, dtype=dtype))
    if len({c.shape[0] for c in cs}) != 1:
        raise NotImplementedError('conjuncts with different numbers of rows')
    n = len(cs)
    return SimpleNamespace(lower_bounds=torch.stack(lo), upper_bounds=torch.stack(hi), cs=torch.stack(cs),
                           rhs=torch.stack(rhs), ids=torch.arange(n) + 3)
