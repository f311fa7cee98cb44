"""Equation-string front end for ``Model.set_equations(equations="...")``.

Mirrors the conventions of the reference parser (``hilo_mpc/util/parsing.py:246-577``): one equation per line,
``dx/dt = ...`` / ``d/dt(x(t)) = ...`` (continuous) or ``x(k+1) = ...`` (discrete) define states, ``y(k) = ...``
defines a measurement, ``name = <number>`` a constant, ``name = <expr>`` an auxiliary expression that is inlined,
``v(k)`` on a right-hand side is an input, any other free name becomes a parameter (in order of appearance), ``^``
means power, ``|`` lines carry descriptions and are ignored here.  Built on Python's ``ast`` like the reference.
"""
from __future__ import annotations

import ast
import math
import re
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import symbolic as sx

FUNCTIONS = {   # the reference's table (parsing.py:36-58) plus a few aliases
    'sqrt': sx.sqrt, 'exp': sx.exp, 'log': sx.log, 'log10': lambda a: sx.log(a) / math.log(10.), 'sign': sx.sign,
    'abs': sx.fabs, 'min': sx.fmin, 'max': sx.fmax, 'sin': sx.sin, 'cos': sx.cos, 'tan': sx.tan,
    'arcsin': sx.asin, 'arccos': sx.acos, 'arctan': sx.atan, 'arctan2': sx.atan2,
    'sinh': sx.sinh, 'cosh': sx.cosh, 'tanh': sx.tanh,
    'arsinh': lambda a: sx.log(a + sx.sqrt(a * a + 1.)), 'arcosh': lambda a: sx.log(a + sx.sqrt(a * a - 1.)),
    'artanh': lambda a: .5 * sx.log((1. + a) / (1. - a)),
    'atan': sx.atan, 'asin': sx.asin, 'acos': sx.acos, 'fabs': sx.fabs, 'atan2': sx.atan2,
}
CONSTANTS = {'pi': math.pi, 'e': math.e, 'inf': math.inf}


class _Evaluator:
    def __init__(self, known: Dict[str, object], params: List[sx.Expr], inputs: List[sx.Expr], discrete: bool,
                 aux: Dict[str, ast.AST]):
        self.known = known
        self.params = params
        self.inputs = inputs
        self.discrete = discrete
        self.aux = aux

    def _name(self, name: str, time_arg: Optional[str] = None):
        if name in self.known:
            return self.known[name]
        if name in self.aux:
            val = self.eval(self.aux[name])
            return val
        if name in CONSTANTS:
            return sx.Expr.const(CONSTANTS[name])
        s = sx.sym(name)
        self.known[name] = s
        if time_arg == 'k' and not self.discrete:
            self.inputs.append(s)  # parsing.py:507-515: v(k) on a continuous right-hand side is an input
        elif time_arg == 't':
            raise NotImplementedError(f"Algebraic variable '{name}(t)' found: DAE models are not supported by the "
                                      f"B200 engine yet")
        else:
            self.params.append(s)
        return s

    def eval(self, node):
        if isinstance(node, ast.Expression):
            return self.eval(node.body)
        if isinstance(node, ast.Constant):
            return sx.Expr.const(float(node.value))
        if isinstance(node, ast.Name):
            return self._name(node.id)
        if isinstance(node, ast.BinOp):
            a, b = self.eval(node.left), self.eval(node.right)
            if isinstance(node.op, ast.Add):
                return a + b
            if isinstance(node.op, ast.Sub):
                return a - b
            if isinstance(node.op, ast.Mult):
                return a * b
            if isinstance(node.op, ast.Div):
                return a / b
            if isinstance(node.op, ast.Pow):
                return a ** b
            raise NotImplementedError(f"Operator {type(node.op).__name__} is not supported")
        if isinstance(node, ast.UnaryOp):
            v = self.eval(node.operand)
            if isinstance(node.op, ast.USub):
                return -v
            if isinstance(node.op, ast.UAdd):
                return v
            raise NotImplementedError(f"Operator {type(node.op).__name__} is not supported")
        if isinstance(node, ast.Call):
            fname = node.func.id if isinstance(node.func, ast.Name) else None
            if fname is None:
                raise NotImplementedError("Only plain function calls are supported")
            if len(node.args) == 1 and isinstance(node.args[0], ast.Name) and node.args[0].id in ('t', 'k') \
                    and fname not in FUNCTIONS:
                return self._name(fname, node.args[0].id)
            if fname in FUNCTIONS:
                return FUNCTIONS[fname](*[self.eval(a) for a in node.args])
            raise NotImplementedError(f"Unknown function '{fname}'")
        raise NotImplementedError(f"Unsupported syntax element {type(node).__name__}")


def parse_dynamic_equations(equations, discrete: bool = False, **kwargs) -> dict:
    """Returns ``{'x','y','u','p','ode','meas','const'}`` with symbol / expression lists."""
    if isinstance(equations, str):
        equations = equations.split('\n')
    known: Dict[str, object] = {}
    model = {'x': [], 'y': [], 'z': [], 'u': [], 'p': [], 'ode': [], 'alg': [], 'meas': [], 'const': {}}
    for key in ('x', 'y', 'u', 'p'):
        given = kwargs.get(key)
        if given is not None:
            for s in given:
                known[s.name] = s
                model[key].append(s)
    for key in ('dt', 't'):
        s = kwargs.get(key)
        if s is not None:
            known[key] = s
    odes, meas, aux = [], [], {}
    derivative_of = {}   # 'dv(t)/dt' / 'd/dt(v(t))' / 'v(k+1)' -> its right-hand side, for nested use (parsing.py:374, :420)
    pending = '...'
    for raw in equations:
        if '|' in raw:
            continue
        # a line ending in '...' is continued on the next one (parsing.py:344-347)
        pending = pending.replace('...', raw.replace(' ', '').replace('\t', ''))
        if pending.endswith('...'):
            continue
        eq, pending = pending, '...'
        if not eq or eq.startswith('#') or '=' not in eq:
            continue
        eq = eq.replace('^', '**')
        lhs, rhs = eq.split('=', 1)
        state = None
        if not discrete:
            m = re.fullmatch(r'd(.+)/dt', lhs)
            if m is not None:
                state = re.sub(r'\(.*?\)', '', m.group(1))
            else:
                m = re.fullmatch(r'd/dt\((.+)\)', lhs)
                if m is not None:
                    state = re.sub(r'\(.*?\)', '', m.group(1))
        else:
            m = re.fullmatch(r'(.*)\(k\+1\)', lhs)
            if m is not None:
                state = m.group(1)
        if state is not None:
            if state not in known:
                s = sx.sym(state)
                known[state] = s
                model['x'].append(s)
            odes.append(rhs)
            derivative_of[lhs] = f'({rhs})'
            continue
        if lhs == '0':
            raise NotImplementedError("Algebraic equations (DAE models) are not supported by the B200 engine yet")
        try:
            model['const'][lhs] = float(rhs)
            known[lhs] = sx.Expr.const(float(rhs))
            continue
        except ValueError:
            pass
        m = re.fullmatch(r'(.*)\(k\)', lhs)
        if m is not None:
            name = m.group(1)
            if name not in known:
                s = sx.sym(name)
                known[name] = s
                model['y'].append(s)
            meas.append(rhs)
            continue
        if re.fullmatch(r'(.*)\(t\)', lhs):
            raise NotImplementedError("Explicit algebraic variables z(t) = ... are not supported by the B200 engine")
        aux[lhs] = ast.parse(rhs, mode='eval')
    ev = _Evaluator(known, model['p'], model['u'], discrete, aux)
    # A left-hand side such as dv(t)/dt may be used inside other equations: the reference substitutes its right-hand
    # side textually, longest keys first, where the match is followed by an operator, ')' or the end (parsing.py:446-467)
    keys = sorted(derivative_of, reverse=True)
    pattern = re.compile('|'.join(re.escape(k) for k in keys)) if keys else None

    def nested(m):
        end = m.end()
        nxt = '+' if end == len(m.string) else m.string[end]
        return pattern.sub(nested, derivative_of[m.group(0)]) if nxt in '+-*/)' else m.group(0)

    def expand(rhs):
        return pattern.sub(nested, rhs) if pattern is not None else rhs
    for rhs in odes:
        model['ode'].append(ev.eval(ast.parse(expand(rhs), mode='eval')))
    for rhs in meas:
        model['meas'].append(ev.eval(ast.parse(expand(rhs), mode='eval')))
    return model


def parse_expression(text: str, known: Dict[str, object], free: Optional[List[sx.Expr]] = None):
    """Parse one right-hand side such as ``'x/2 + 25*dt*x/(1 + x^2)'`` with ``known`` name -> symbol bindings; unknown
    names are appended to ``free`` (new parameters)."""
    free = [] if free is None else free
    ev = _Evaluator(dict(known), free, [], True, {})
    out = ev.eval(ast.parse(text.replace(' ', '').replace('^', '**'), mode='eval'))
    return out, free
