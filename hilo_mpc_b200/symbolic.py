"""Tiny symbolic layer: hash-consed scalar expression DAG with forward-mode AD and code generation.

The reference delegates every expression to CasADi ``SX`` (e.g. ``hilo_mpc/modules/dynamic_model/
dynamic_model.py:2882-2994`` builds the ODE/measurement ``ca.Function`` objects and IPOPT gets its
Jacobians/Hessians from CasADi AD, ``hilo_mpc/modules/controller/mpc.py:1778-1787``).  CasADi is not
available here, and the B200 engine needs the residuals/Jacobians as CUDA device functions anyway,
so this module provides the three things the hot path needs and nothing more:

* scalar expression nodes with operator overloading (vectors/matrices are ``numpy`` object arrays),
* symbolic differentiation (``jacobian``/``gradient``/``hessian``) with common-subexpression sharing,
* emitters: CUDA/C source for device functions and NumPy-vectorised Python callables.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

Number = Union[int, float]

_UNARY = ('neg', 'sin', 'cos', 'tan', 'exp', 'log', 'sqrt', 'tanh', 'atan', 'asin', 'acos', 'sinh', 'cosh', 'abs',
          'sign', 'sigmoid')
_BINARY = ('add', 'sub', 'mul', 'div', 'pow', 'atan2', 'fmin', 'fmax')


class Expr:
    """One node of the expression DAG. Nodes are interned, so ``a is b`` iff structurally equal."""
    __slots__ = ('op', 'args', 'value', 'name', 'id', '__weakref__')
    __array_priority__ = 1000.  # let numpy defer to our reflected operators elementwise

    _table: Dict[tuple, 'Expr'] = {}
    _count = 0

    def __new__(cls, op: str, args: tuple = (), value: Optional[float] = None, name: Optional[str] = None):
        if op == 'sym':
            key = None  # symbols are unique objects even when names collide
        elif op == 'const':
            key = ('const', float(value).hex())
        else:
            key = (op,) + tuple(a.id for a in args)
        if key is not None:
            hit = cls._table.get(key)
            if hit is not None:
                return hit
        self = object.__new__(cls)
        self.op = op
        self.args = args
        self.value = None if value is None else float(value)
        self.name = name
        self.id = Expr._count
        Expr._count += 1
        if key is not None:
            cls._table[key] = self
        return self

    # ------------------------------------------------------------------ construction helpers
    @staticmethod
    def const(v: Number) -> 'Expr':
        return Expr('const', (), value=float(v))

    @staticmethod
    def wrap(v) -> 'Expr':
        if isinstance(v, Expr):
            return v
        if isinstance(v, (int, float, np.integer, np.floating)):
            return Expr.const(float(v))
        if isinstance(v, np.ndarray) and v.size == 1:
            return Expr.wrap(v.reshape(-1)[0])
        raise TypeError(f"Cannot convert {type(v).__name__} to a symbolic expression")

    def is_const(self, v: Optional[float] = None) -> bool:
        return self.op == 'const' and (v is None or self.value == v)

    def is_symbolic(self) -> bool:
        return True

    # ------------------------------------------------------------------ arithmetic
    def __add__(self, o):
        if isinstance(o, np.ndarray):
            return NotImplemented
        return _add(self, Expr.wrap(o))

    def __radd__(self, o):
        return _add(Expr.wrap(o), self)

    def __sub__(self, o):
        if isinstance(o, np.ndarray):
            return NotImplemented
        return _sub(self, Expr.wrap(o))

    def __rsub__(self, o):
        return _sub(Expr.wrap(o), self)

    def __mul__(self, o):
        if isinstance(o, np.ndarray):
            return NotImplemented
        return _mul(self, Expr.wrap(o))

    def __rmul__(self, o):
        return _mul(Expr.wrap(o), self)

    def __truediv__(self, o):
        if isinstance(o, np.ndarray):
            return NotImplemented
        return _div(self, Expr.wrap(o))

    def __rtruediv__(self, o):
        return _div(Expr.wrap(o), self)

    def __pow__(self, o):
        if isinstance(o, np.ndarray):
            return NotImplemented
        return _pow(self, Expr.wrap(o))

    def __rpow__(self, o):
        return _pow(Expr.wrap(o), self)

    def __neg__(self):
        return _neg(self)

    def __pos__(self):
        return self

    def __abs__(self):
        return _unary('abs', self)

    def __float__(self):
        if self.op == 'const':
            return self.value
        raise TypeError("Symbolic expression cannot be converted to float")

    def __hash__(self):
        return self.id

    def __eq__(self, o):  # identity semantics (needed for dict keys); use ``is`` in library code
        return self is o

    def __repr__(self):
        return to_string(self)

    # numpy ufunc hooks so np.sin(expr_array) etc. work elementwise on object arrays
    def sin(self):
        return sin(self)

    def cos(self):
        return cos(self)

    def tan(self):
        return tan(self)

    def exp(self):
        return exp(self)

    def log(self):
        return log(self)

    def sqrt(self):
        return sqrt(self)

    def tanh(self):
        return tanh(self)

    def arctan(self):
        return atan(self)

    def arcsin(self):
        return _unary('asin', self)

    def arccos(self):
        return _unary('acos', self)

    def sinh(self):
        return _unary('sinh', self)

    def cosh(self):
        return _unary('cosh', self)


# ---------------------------------------------------------------------- node constructors with simplification
def _add(a: Expr, b: Expr) -> Expr:
    if a.op == 'const' and b.op == 'const':
        return Expr.const(a.value + b.value)
    if a.is_const(0.):
        return b
    if b.is_const(0.):
        return a
    if b.op == 'neg':
        return _sub(a, b.args[0])
    if a.op == 'neg':
        return _sub(b, a.args[0])
    return Expr('add', (a, b))


def _sub(a: Expr, b: Expr) -> Expr:
    if a.op == 'const' and b.op == 'const':
        return Expr.const(a.value - b.value)
    if b.is_const(0.):
        return a
    if a.is_const(0.):
        return _neg(b)
    if a is b:
        return Expr.const(0.)
    if b.op == 'neg':
        return _add(a, b.args[0])
    return Expr('sub', (a, b))


def _mul(a: Expr, b: Expr) -> Expr:
    if a.op == 'const' and b.op == 'const':
        return Expr.const(a.value * b.value)
    if a.is_const(0.) or b.is_const(0.):
        return Expr.const(0.)
    if a.is_const(1.):
        return b
    if b.is_const(1.):
        return a
    if a.is_const(-1.):
        return _neg(b)
    if b.is_const(-1.):
        return _neg(a)
    if a.op == 'neg' and b.op == 'neg':
        return _mul(a.args[0], b.args[0])
    if b.op == 'const' and a.op != 'const':  # canonical order: constant first
        a, b = b, a
    return Expr('mul', (a, b))


def _div(a: Expr, b: Expr) -> Expr:
    if b.op == 'const':
        if b.value == 1.:
            return a
        if a.op == 'const':
            return Expr.const(a.value / b.value)
    if a.is_const(0.):
        return Expr.const(0.)
    if a is b:
        return Expr.const(1.)
    return Expr('div', (a, b))


def _neg(a: Expr) -> Expr:
    if a.op == 'const':
        return Expr.const(-a.value)
    if a.op == 'neg':
        return a.args[0]
    return Expr('neg', (a,))


def _pow(a: Expr, b: Expr) -> Expr:
    if b.op == 'const':
        if b.value == 0.:
            return Expr.const(1.)
        if b.value == 1.:
            return a
        if a.op == 'const':
            return Expr.const(a.value ** b.value)
        if b.value == 2.:
            return _mul(a, a) if a.op in ('sym', 'const') else Expr('pow', (a, b))
        if b.value == .5:
            return sqrt(a)
    return Expr('pow', (a, b))


_CONST_FOLD = {
    'sin': math.sin, 'cos': math.cos, 'tan': math.tan, 'exp': math.exp, 'log': math.log, 'sqrt': math.sqrt,
    'tanh': math.tanh, 'atan': math.atan, 'asin': math.asin, 'acos': math.acos, 'sinh': math.sinh,
    'cosh': math.cosh, 'abs': abs, 'sign': lambda v: float(np.sign(v)), 'sigmoid': lambda v: 1. / (1. + math.exp(-v)),
}


def _unary(op: str, a) -> Expr:
    a = Expr.wrap(a)
    if a.op == 'const':
        try:
            return Expr.const(_CONST_FOLD[op](a.value))
        except (ValueError, OverflowError):
            pass
    return Expr(op, (a,))


def _elementwise(fun):
    def wrapper(a, *rest):
        if isinstance(a, np.ndarray):
            out = np.empty(a.shape, dtype=object)
            flat = out.reshape(-1)
            for k, v in enumerate(a.reshape(-1)):
                flat[k] = fun(v, *rest)
            return out
        return fun(a, *rest)
    wrapper.__name__ = fun.__name__
    return wrapper


@_elementwise
def sin(a): return _unary('sin', a)
@_elementwise
def cos(a): return _unary('cos', a)
@_elementwise
def tan(a): return _unary('tan', a)
@_elementwise
def exp(a): return _unary('exp', a)
@_elementwise
def log(a): return _unary('log', a)
@_elementwise
def sqrt(a): return _unary('sqrt', a)
@_elementwise
def tanh(a): return _unary('tanh', a)
@_elementwise
def atan(a): return _unary('atan', a)
@_elementwise
def asin(a): return _unary('asin', a)
@_elementwise
def acos(a): return _unary('acos', a)
@_elementwise
def sinh(a): return _unary('sinh', a)
@_elementwise
def cosh(a): return _unary('cosh', a)
@_elementwise
def fabs(a): return _unary('abs', a)
@_elementwise
def sign(a): return _unary('sign', a)
@_elementwise
def sigmoid(a): return _unary('sigmoid', a)


def atan2(a, b) -> Expr:
    return Expr('atan2', (Expr.wrap(a), Expr.wrap(b)))


def fmin(a, b) -> Expr:
    return Expr('fmin', (Expr.wrap(a), Expr.wrap(b)))


def fmax(a, b) -> Expr:
    return Expr('fmax', (Expr.wrap(a), Expr.wrap(b)))


# ---------------------------------------------------------------------- learned components as call nodes
# A Gaussian-process mean with hundreds of training points is not unrolled into the expression graph: it enters as a
# node ``gp:<id>:<what>`` whose arguments are the query coordinates; ``<what>`` = m (mean), g<j> (d/dx_j) or h<j>_<l>
# (d2/dx_j dx_l, j >= l).  Differentiation maps m -> g -> h; the emitters turn all nodes of one (GP, query) pair into ONE
# call of a loop over the training table (``hb::gp_se_eval`` in csrc/gp_device.cuh; a vectorised NumPy loop for the
# numeric evaluation used by the host-side tools and the oracle).
GP_TABLES: Dict[str, dict] = {}


def register_gp_table(Xs: np.ndarray, c: np.ndarray, inv_ls: np.ndarray) -> str:
    """``Xs (n, d)``: training inputs divided by the length scales, ``c (n,)``: ``alpha_i * sigma_f^2``,
    ``inv_ls (d,)``.  Returns the table id (content hash, so that equal GPs share a table)."""
    import hashlib
    Xs, c, inv_ls = (np.ascontiguousarray(v, dtype=float) for v in (Xs, c, inv_ls))
    gid = hashlib.sha256(Xs.tobytes() + c.tobytes() + inv_ls.tobytes()).hexdigest()[:10]
    GP_TABLES[gid] = dict(Xs=Xs, c=c, inv_ls=inv_ls, d=Xs.shape[1], n=Xs.shape[0])
    return gid


def gp_mean_node(gid: str, xq) -> 'Expr':
    args = tuple(Expr.wrap(v) for v in as_vector(xq))
    if len(args) != GP_TABLES[gid]['d']:
        raise ValueError("gp_mean_node: dimension of the query point does not match the table")
    return Expr(f'gp:{gid}:m', args)


MLP_TABLES: Dict[str, dict] = {}


def register_mlp_table(W1, b1, W2, b2, activation: str) -> str:
    """One-hidden-layer perceptron ``y = W2 act(W1 z + b1) + b2`` (scalings already folded in); ``activation`` 'tanh' or
    'sigmoid'.  Returns the table id."""
    import hashlib
    W1, b1, W2, b2 = (np.ascontiguousarray(v, dtype=float) for v in (W1, b1, W2, b2))
    act = {'tanh': 0, 'sigmoid': 1, 'logistic': 1}[activation]
    mid = hashlib.sha256(W1.tobytes() + b1.tobytes() + W2.tobytes() + b2.tobytes() + bytes([act])).hexdigest()[:10]
    MLP_TABLES[mid] = dict(W1=W1, b1=b1.reshape(-1), W2=W2, b2=b2.reshape(-1), act=act, nz=W1.shape[1], nh=W1.shape[0],
                           no=W2.shape[0])
    return mid


def mlp_output_nodes(mid: str, z) -> np.ndarray:
    args = tuple(Expr.wrap(v) for v in as_vector(z))
    t = MLP_TABLES[mid]
    if len(args) != t['nz']:
        raise ValueError("mlp_output_nodes: dimension of the input does not match the table")
    return np.array([Expr(f'mlp:{mid}:y{o}', args) for o in range(t['no'])], dtype=object)


def _mlp_out_index(what: str, t: dict) -> int:
    nz, no = t['nz'], t['no']
    nt = nz * (nz + 1) // 2
    if what[0] == 'y':
        return int(what[1:])
    if what[0] == 'j':
        o, i = (int(v) for v in what[1:].split('_'))
        return no + o * nz + i
    o, i, l = (int(v) for v in what[1:].split('_'))
    return no + no * nz + o * nt + i * (i + 1) // 2 + l


def _mlp_numpy(mid: str, order: int, *z):
    """Numeric evaluation of one MLP group: list of outputs in the order of ``hb::mlp1_eval``."""
    t = MLP_TABLES[mid]
    nz, nh, no = t['nz'], t['nh'], t['no']
    zz = [np.asarray(v) for v in z]
    shape = np.broadcast(*zz).shape
    a = sum(zz[i][..., None] * t['W1'][:, i] for i in range(nz)) + t['b1']          # (..., nh)
    if t['act'] == 0:
        h = np.tanh(a); d1 = 1. - h * h; d2 = -2. * h * d1
    else:
        h = 1. / (1. + np.exp(-a)); d1 = h * (1. - h); d2 = d1 * (1. - 2. * h)
    out = [(h * t['W2'][o]).sum(axis=-1) + t['b2'][o] for o in range(no)]
    if order >= 1:
        for o in range(no):
            for i in range(nz):
                out.append((d1 * t['W2'][o] * t['W1'][:, i]).sum(axis=-1))
    if order >= 2:
        for o in range(no):
            for i in range(nz):
                for l in range(i + 1):
                    out.append((d2 * t['W2'][o] * t['W1'][:, i] * t['W1'][:, l]).sum(axis=-1))
    return [np.broadcast_to(v, shape) for v in out]


def _gp_parts(op: str):
    _, gid, what = op.split(':')
    return gid, what


def _gp_out_index(what: str, d: int) -> int:
    if what == 'm':
        return 0
    if what[0] == 'g':
        return 1 + int(what[1:])
    j, l = (int(v) for v in what[1:].split('_'))
    return 1 + d + j * (j + 1) // 2 + l


def _gp_numpy(gid: str, order: int, *xq):
    """Numeric evaluation of one GP group: returns the list of outputs [m, g_0.., h_00, h_10, h_11, ...] (each shaped like
    the broadcast query coordinates; complex arguments allowed for complex-step checks)."""
    t = GP_TABLES[gid]
    d, Xs, c, il = t['d'], t['Xs'], t['c'], t['inv_ls']
    xs = [np.asarray(xq[j]) * il[j] for j in range(d)]
    shape = np.broadcast(*xs).shape
    r = [xs[j][..., None] - Xs[:, j] for j in range(d)]               # (..., n)
    ck = c * np.exp(-.5 * sum(rj * rj for rj in r))
    out = [ck.sum(axis=-1)]
    if order >= 1:
        out += [-(ck * r[j]).sum(axis=-1) * il[j] for j in range(d)]
    if order >= 2:
        for j in range(d):
            for l in range(j + 1):
                out.append((ck * (r[j] * r[l] - (1. if j == l else 0.))).sum(axis=-1) * il[j] * il[l])
    return [np.broadcast_to(o, shape) for o in out]


# ---------------------------------------------------------------------- vectors
def sym(name: str, n: Optional[int] = None, m: Optional[int] = None):
    """``sym('x')`` -> scalar symbol; ``sym('x', 3)`` -> (3,) object array ``x_0..x_2``; ``sym('P', 2, 2)`` matrix."""
    if n is None:
        return Expr('sym', (), name=name)
    if m is None:
        return np.array([Expr('sym', (), name=f'{name}_{k}') for k in range(n)], dtype=object)
    out = np.empty((n, m), dtype=object)
    for i in range(n):
        for j in range(m):
            out[i, j] = Expr('sym', (), name=f'{name}_{i}_{j}')
    return out


def named_syms(names: Sequence[str]) -> np.ndarray:
    return np.array([Expr('sym', (), name=str(k)) for k in names], dtype=object)


def as_vector(x) -> np.ndarray:
    """Coerce scalars / lists / arrays of numbers and expressions into a flat object array of ``Expr``."""
    if isinstance(x, Expr):
        return np.array([x], dtype=object)
    if x is None:
        return np.empty((0,), dtype=object)
    arr = np.asarray(x, dtype=object).reshape(-1)
    out = np.empty(arr.shape, dtype=object)
    for k, v in enumerate(arr):
        out[k] = Expr.wrap(v)
    return out


def is_symbolic(x) -> bool:
    if isinstance(x, Expr):
        return x.op != 'const'
    if isinstance(x, np.ndarray) and x.dtype == object:
        return any(isinstance(v, Expr) and v.op != 'const' for v in x.reshape(-1))
    if isinstance(x, (list, tuple)):
        return any(is_symbolic(v) for v in x)
    return False


# ---------------------------------------------------------------------- traversal
def topo_sort(outputs: Iterable[Expr]) -> List[Expr]:
    order: List[Expr] = []
    seen = set()
    for root in outputs:
        if root.id in seen:
            continue
        stack = [(root, 0)]
        while stack:
            node, idx = stack.pop()
            if idx == 0 and node.id in seen:
                continue
            if idx < len(node.args):
                stack.append((node, idx + 1))
                child = node.args[idx]
                if child.id not in seen:
                    stack.append((child, 0))
            else:
                if node.id not in seen:
                    seen.add(node.id)
                    order.append(node)
    return order


def symbols_of(outputs: Iterable[Expr]) -> List[Expr]:
    return [n for n in topo_sort(outputs) if n.op == 'sym']


def depends_on(expr, variables) -> bool:
    vs = {v.id for v in as_vector(variables)}
    return any(n.id in vs for n in topo_sort(as_vector(expr)))


def substitute(exprs, old, new):
    """Replace symbols ``old`` by expressions ``new`` (like ``ca.substitute``). Shape of ``exprs`` is preserved."""
    scalar = isinstance(exprs, Expr)
    arr = np.asarray(exprs, dtype=object) if not scalar else np.array([exprs], dtype=object)
    shape = arr.shape
    flat = [Expr.wrap(v) for v in arr.reshape(-1)]
    old_v = as_vector(old)
    new_v = as_vector(new)
    if len(old_v) != len(new_v):
        raise ValueError("substitute: dimension mismatch")
    memo: Dict[int, Expr] = {o.id: n for o, n in zip(old_v, new_v)}
    for node in topo_sort(flat):
        if node.id in memo:
            continue
        if not node.args:
            memo[node.id] = node
            continue
        new_args = tuple(memo[a.id] for a in node.args)
        if all(n is a for n, a in zip(new_args, node.args)):
            memo[node.id] = node
        else:
            memo[node.id] = _rebuild(node.op, new_args)
    out = np.array([memo[f.id] for f in flat], dtype=object).reshape(shape)
    return out[0] if scalar else out


def _rebuild(op: str, args: tuple) -> Expr:
    if op == 'add':
        return _add(*args)
    if op == 'sub':
        return _sub(*args)
    if op == 'mul':
        return _mul(*args)
    if op == 'div':
        return _div(*args)
    if op == 'pow':
        return _pow(*args)
    if op == 'neg':
        return _neg(*args)
    if op in ('atan2', 'fmin', 'fmax') or op.startswith('gp:') or op.startswith('mlp:'):
        return Expr(op, args)
    return _unary(op, args[0])


# ---------------------------------------------------------------------- differentiation
def _deriv_node(node: Expr, d: List[Optional[Expr]]) -> Optional[Expr]:
    """Forward tangent of ``node`` given tangents ``d`` of its args (``None`` means structurally zero)."""
    op, a = node.op, node.args
    if op == 'add':
        if d[0] is None:
            return d[1]
        if d[1] is None:
            return d[0]
        return _add(d[0], d[1])
    if op == 'sub':
        if d[1] is None:
            return d[0]
        if d[0] is None:
            return _neg(d[1])
        return _sub(d[0], d[1])
    if op == 'mul':
        t0 = None if d[0] is None else _mul(d[0], a[1])
        t1 = None if d[1] is None else _mul(a[0], d[1])
        if t0 is None:
            return t1
        if t1 is None:
            return t0
        return _add(t0, t1)
    if op == 'div':
        t0 = None if d[0] is None else _div(d[0], a[1])
        t1 = None if d[1] is None else _mul(_div(node, a[1]), d[1])
        if t1 is None:
            return t0
        if t0 is None:
            return _neg(t1)
        return _sub(t0, t1)
    if op == 'neg':
        return None if d[0] is None else _neg(d[0])
    if op == 'pow':
        base, ex = a
        t = None
        if d[0] is not None:
            if ex.op == 'const':
                t = _mul(_mul(ex, _pow(base, Expr.const(ex.value - 1.))), d[0])
            else:
                t = _mul(_mul(ex, _pow(base, _sub(ex, Expr.const(1.)))), d[0])
        if d[1] is not None:
            t1 = _mul(_mul(node, log(base)), d[1])
            t = t1 if t is None else _add(t, t1)
        return t
    if op == 'atan2':
        y, x = a
        den = _add(_mul(x, x), _mul(y, y))
        t = None
        if d[0] is not None:
            t = _mul(_div(x, den), d[0])
        if d[1] is not None:
            t1 = _mul(_div(y, den), d[1])
            t = _neg(t1) if t is None else _sub(t, t1)
        return t
    if op in ('fmin', 'fmax'):
        raise NotImplementedError("fmin/fmax are not differentiable in this layer")
    if op.startswith('mlp:'):
        mid, what = _gp_parts(op)
        t = None
        for i, di in enumerate(d):
            if di is None:
                continue
            if what[0] == 'y':
                part = Expr(f'mlp:{mid}:j{what[1:]}_{i}', a)
            elif what[0] == 'j':
                o, i0 = (int(v) for v in what[1:].split('_'))
                part = Expr(f'mlp:{mid}:h{o}_{max(i0, i)}_{min(i0, i)}', a)
            else:
                raise NotImplementedError("Third derivatives of an MLP call node are not available")
            term = _mul(part, di)
            t = term if t is None else _add(t, term)
        return t
    if op.startswith('gp:'):
        gid, what = _gp_parts(op)
        t = None
        for j, dj in enumerate(d):
            if dj is None:
                continue
            if what == 'm':
                part = Expr(f'gp:{gid}:g{j}', a)
            elif what[0] == 'g':
                i = int(what[1:])
                part = Expr(f'gp:{gid}:h{max(i, j)}_{min(i, j)}', a)
            else:
                raise NotImplementedError("Third derivatives of a GP mean are not available")
            term = _mul(part, dj)
            t = term if t is None else _add(t, term)
        return t
    if d[0] is None:
        return None
    x = a[0]
    if op == 'sin':
        g = cos(x)
    elif op == 'cos':
        g = _neg(sin(x))
    elif op == 'tan':
        g = _add(Expr.const(1.), _mul(node, node))
    elif op == 'exp':
        g = node
    elif op == 'log':
        g = _div(Expr.const(1.), x)
    elif op == 'sqrt':
        g = _div(Expr.const(.5), node)
    elif op == 'tanh':
        g = _sub(Expr.const(1.), _mul(node, node))
    elif op == 'atan':
        g = _div(Expr.const(1.), _add(Expr.const(1.), _mul(x, x)))
    elif op == 'asin':
        g = _div(Expr.const(1.), sqrt(_sub(Expr.const(1.), _mul(x, x))))
    elif op == 'acos':
        g = _neg(_div(Expr.const(1.), sqrt(_sub(Expr.const(1.), _mul(x, x)))))
    elif op == 'sinh':
        g = cosh(x)
    elif op == 'cosh':
        g = sinh(x)
    elif op == 'abs':
        g = sign(x)
    elif op == 'sign':
        return None
    elif op == 'sigmoid':
        g = _mul(node, _sub(Expr.const(1.), node))
    else:
        raise NotImplementedError(op)
    return _mul(g, d[0])


def jacobian(exprs, variables) -> np.ndarray:
    """Dense symbolic Jacobian ``d exprs / d variables`` as an ``(m, n)`` object array (forward mode, one sweep per
    variable over the nodes that depend on it)."""
    f = as_vector(exprs)
    v = as_vector(variables)
    order = topo_sort(f)
    J = np.empty((len(f), len(v)), dtype=object)
    zero = Expr.const(0.)
    for j, var in enumerate(v):
        tang: Dict[int, Optional[Expr]] = {}
        for node in order:
            if node is var:
                tang[node.id] = Expr.const(1.)
            elif not node.args:
                tang[node.id] = None
            else:
                ds = [tang[a.id] for a in node.args]
                tang[node.id] = None if all(t is None for t in ds) else _deriv_node(node, ds)
        for i, fi in enumerate(f):
            t = tang.get(fi.id)
            J[i, j] = zero if t is None else t
    return J


def gradient(expr, variables) -> np.ndarray:
    return jacobian(as_vector(expr)[:1], variables)[0]


def hessian(expr, variables) -> np.ndarray:
    g = gradient(expr, variables)
    H = jacobian(g, variables)
    # symmetrise structurally (identical by construction up to node ordering)
    n = H.shape[0]
    for i in range(n):
        for j in range(i):
            H[i, j] = H[j, i]
    return H


# ---------------------------------------------------------------------- printing / code generation
_INFIX = {'add': '+', 'sub': '-', 'mul': '*', 'div': '/'}
_C_UNARY = {'sin': 'sin', 'cos': 'cos', 'tan': 'tan', 'exp': 'exp', 'log': 'log', 'sqrt': 'sqrt', 'tanh': 'tanh',
            'atan': 'atan', 'asin': 'asin', 'acos': 'acos', 'sinh': 'sinh', 'cosh': 'cosh', 'abs': 'fabs'}
_NP_UNARY = {'sin': 'np.sin', 'cos': 'np.cos', 'tan': 'np.tan', 'exp': 'np.exp', 'log': 'np.log', 'sqrt': 'np.sqrt',
             'tanh': 'np.tanh', 'atan': 'np.arctan', 'asin': 'np.arcsin', 'acos': 'np.arccos', 'sinh': 'np.sinh',
             'cosh': 'np.cosh', 'abs': 'np.abs', 'sign': 'np.sign'}


def to_string(e: Expr, depth: int = 0) -> str:
    if depth > 6:
        return '…'
    if e.op == 'const':
        return repr(e.value)
    if e.op == 'sym':
        return e.name
    if e.op in _INFIX:
        return f'({to_string(e.args[0], depth + 1)} {_INFIX[e.op]} {to_string(e.args[1], depth + 1)})'
    if e.op == 'neg':
        return f'(-{to_string(e.args[0], depth + 1)})'
    return f'{e.op}(' + ', '.join(to_string(a, depth + 1) for a in e.args) + ')'


def _c_const(v: float) -> str:
    if math.isinf(v):
        return 'HUGE_VAL' if v > 0 else '-HUGE_VAL'
    if math.isnan(v):
        return 'NAN'
    r = repr(float(v))
    if 'e' not in r and '.' not in r:
        r += '.0'
    return r


def _cache_copy(var: str, off: int, nout: int, load: bool) -> str:
    """C statement that moves the outputs of one call group from / to the call cache ``cc`` (a rolled loop: the cache is
    meant to live in thread-local memory, not in registers, between the two functions)."""
    loop = f'for (int e_ = 0; e_ < {nout}; ++e_)'
    if load:
        return f'double {var}[{nout}]; {loop} {var}[e_] = cc[{off} + e_];'
    return f'{loop} cc[{off} + e_] = {var}[e_];'


def _emit(outputs: List[Expr], inputs: Dict[int, str], lang: str, tmp: str = 't',
          call_cache: Optional[dict] = None) -> Tuple[List[str], List[str]]:
    """Straight-line code for ``outputs``; returns (statements, output operand strings).

    ``call_cache`` (C only) shares the results of call nodes (GP / MLP loops) between two generated functions that are
    evaluated at the same point one after the other - Jacobian first, Lagrangian Hessian later: with ``mode='store'``
    every call group is evaluated at derivative order 2 whatever this function needs and its outputs are copied to the
    array ``cc`` (slots recorded in ``call_cache['slots']``, total in ``call_cache['size']``); with ``mode='load'`` groups
    found in the slots read ``cc`` instead of walking the table again."""
    order = topo_sort(outputs)
    uses: Dict[int, int] = {}
    for node in order:
        for a in node.args:
            uses[a.id] = uses.get(a.id, 0) + 1
    names: Dict[int, str] = {}
    lines: List[str] = []
    gp_groups: Dict[tuple, str] = {}
    order_nodes = order
    k = 0
    for node in order:
        if node.op == 'const':
            names[node.id] = _c_const(node.value) if lang == 'c' else repr(node.value)
            continue
        if node.op == 'sym':
            if node.id not in inputs:
                raise ValueError(f"Free symbol '{node.name}' is not an input of the function being generated")
            names[node.id] = inputs[node.id]
            continue
        a = [names[x.id] for x in node.args]
        op = node.op
        if op.startswith('mlp:'):
            mid, what = _gp_parts(op)
            key = ('mlp', mid) + tuple(x.id for x in node.args)
            t = MLP_TABLES[mid]
            if key not in gp_groups:
                order = 0
                for other in order_nodes:
                    if other.op.startswith(f'mlp:{mid}:') and tuple(x.id for x in other.args) == key[2:]:
                        w2 = _gp_parts(other.op)[1]
                        order = max(order, {'y': 0, 'j': 1, 'h': 2}[w2[0]])
                nz, nh, no = t['nz'], t['nh'], t['no']
                var = f'{tmp}mlp{len(gp_groups)}'
                if lang == 'c':
                    if call_cache is not None and call_cache['mode'] == 'store':
                        order = 2
                    nout = no + (no * nz if order >= 1 else 0) + (no * nz * (nz + 1) // 2 if order >= 2 else 0)
                    if call_cache is not None and call_cache['mode'] == 'load' and key in call_cache['slots']:
                        lines.append(_cache_copy(var, *call_cache['slots'][key], load=True))
                    else:
                        lines.append(f'double {var}[{nout}]; {{ const double {var}z[{nz}] = {{{", ".join(a)}}}; '
                                     f'hb::mlp1_eval<{nz}, {nh}, {no}, {order}, {t["act"]}>(hb_mlp_tab_{mid}, {var}z, {var}); }}')
                        if call_cache is not None and call_cache['mode'] == 'store':
                            call_cache['slots'][key] = (call_cache['size'], nout)
                            lines.append(_cache_copy(var, call_cache['size'], nout, load=False))
                            call_cache['size'] += nout
                else:
                    lines.append(f'{var} = _mlp_numpy({mid!r}, {order}, {", ".join(a)})')
                gp_groups[key] = var
            names[node.id] = f'{gp_groups[key]}[{_mlp_out_index(what, t)}]'
            continue
        if op.startswith('gp:'):
            gid, what = _gp_parts(op)
            key = (gid,) + tuple(x.id for x in node.args)
            if key not in gp_groups:
                # highest derivative order any node of this (GP, query) group asks for -> one call for all of them
                order = 0
                for other in order_nodes:
                    if other.op.startswith(f'gp:{gid}:') and tuple(x.id for x in other.args) == key[1:]:
                        w2 = _gp_parts(other.op)[1]
                        order = max(order, 0 if w2 == 'm' else (1 if w2[0] == 'g' else 2))
                dd = GP_TABLES[gid]['d']
                var = f'{tmp}gp{len(gp_groups)}'
                if lang == 'c':
                    if call_cache is not None and call_cache['mode'] == 'store':
                        order = 2
                    nout = 1 + (dd if order >= 1 else 0) + (dd * (dd + 1) // 2 if order >= 2 else 0)
                    if call_cache is not None and call_cache['mode'] == 'load' and key in call_cache['slots']:
                        lines.append(_cache_copy(var, *call_cache['slots'][key], load=True))
                    else:
                        lines.append(f'double {var}[{nout}]; {{ const double {var}x[{dd}] = {{{", ".join(a)}}}; '
                                     f'hb::gp_se_eval<{dd}, {order}>(hb_gp_tab_{gid}, {GP_TABLES[gid]["n"]}, hb_gp_ils_{gid}, '
                                     f'{var}x, {var}); }}')
                        if call_cache is not None and call_cache['mode'] == 'store':
                            call_cache['slots'][key] = (call_cache['size'], nout)
                            lines.append(_cache_copy(var, call_cache['size'], nout, load=False))
                            call_cache['size'] += nout
                else:
                    lines.append(f'{var} = _gp_numpy({gid!r}, {order}, {", ".join(a)})')
                gp_groups[key] = var
            names[node.id] = f'{gp_groups[key]}[{_gp_out_index(what, GP_TABLES[gid]["d"])}]'
            continue
        if op in _INFIX:
            rhs = f'{a[0]} {_INFIX[op]} {a[1]}'
        elif op == 'neg':
            rhs = f'-{a[0]}'
        elif op == 'pow':
            ex = node.args[1]
            if ex.op == 'const' and ex.value == int(ex.value) and 2 <= abs(ex.value) <= 4:
                p = '*'.join([a[0]] * int(abs(ex.value)))
                rhs = p if ex.value > 0 else f'1.0/({p})'
            else:
                rhs = f'pow({a[0]}, {a[1]})' if lang == 'c' else f'np.power({a[0]}, {a[1]})'
        elif op == 'atan2':
            rhs = f'atan2({a[0]}, {a[1]})' if lang == 'c' else f'np.arctan2({a[0]}, {a[1]})'
        elif op == 'fmin':
            rhs = f'fmin({a[0]}, {a[1]})' if lang == 'c' else f'np.minimum({a[0]}, {a[1]})'
        elif op == 'fmax':
            rhs = f'fmax({a[0]}, {a[1]})' if lang == 'c' else f'np.maximum({a[0]}, {a[1]})'
        elif op == 'sign':
            rhs = f'(({a[0]} > 0.0) - ({a[0]} < 0.0))' if lang == 'c' else f'np.sign({a[0]})'
        elif op == 'sigmoid':
            rhs = f'1.0/(1.0 + exp(-{a[0]}))' if lang == 'c' else f'1.0/(1.0 + np.exp(-{a[0]}))'
        else:
            rhs = f'{(_C_UNARY if lang == "c" else _NP_UNARY)[op]}({a[0]})'
        name = f'{tmp}{k}'
        k += 1
        names[node.id] = name
        lines.append(f'const double {name} = {rhs};' if lang == 'c' else f'{name} = {rhs}')
    return lines, [names[o.id] for o in outputs]


class Function:
    """Named multi-input/multi-output function of symbolic vectors (the role ``ca.Function`` plays in the
    reference, e.g. ``hilo_mpc/modules/base.py:320-360``).

    ``inputs``/``outputs`` are (name, array-of-Expr) pairs; arrays may be 1-D or 2-D (2-D outputs are emitted
    row-major).
    """

    def __init__(self, name: str, inputs: Sequence[Tuple[str, object]], outputs: Sequence[Tuple[str, object]]):
        self.name = name
        self.in_names = [n for n, _ in inputs]
        self.out_names = [n for n, _ in outputs]
        self.inputs = [as_vector(v) for _, v in inputs]
        self.out_shapes = [np.asarray(v, dtype=object).shape if not isinstance(v, Expr) else () for _, v in outputs]
        self.outputs = [as_vector(v) for _, v in outputs]
        self._np = None

    # -- numeric evaluation through generated NumPy source (vectorised over a trailing batch axis)
    def _build_numpy(self) -> Callable:
        in_map: Dict[int, str] = {}
        for iname, vec in zip(self.in_names, self.inputs):
            for k, s in enumerate(vec):
                if s.op != 'sym':
                    raise ValueError(f"Input '{iname}' must consist of pure symbols")
                in_map[s.id] = f'{iname}[{k}]'
        flat = [o for vec in self.outputs for o in vec]
        lines, outs = _emit(flat, in_map, 'py')
        src = [f"def {self.name}({', '.join(self.in_names)}):"]
        src += ['    ' + ln for ln in lines]
        pos = 0
        rets = []
        for oname, vec in zip(self.out_names, self.outputs):
            rets.append('[' + ', '.join(outs[pos:pos + len(vec)]) + ']')
            pos += len(vec)
        src.append('    return (' + ', '.join(rets) + ',)')
        ns = {'np': np, '_gp_numpy': _gp_numpy, '_mlp_numpy': _mlp_numpy}
        exec('\n'.join(src), ns)
        return ns[self.name]

    def __call__(self, *args):
        """Evaluate numerically. Each argument is array-like of shape ``(n_i,)`` or ``(n_i, batch)``; outputs are
        float (or complex, for complex-step differentiation) arrays shaped like the symbolic outputs, with the batch
        axis appended when present."""
        if self._np is None:
            self._np = self._build_numpy()
        cast = []
        for a, vec in zip(args, self.inputs):
            a = np.asarray(a)
            if a.ndim == 0:
                a = a.reshape(1)
            if a.shape[0] != len(vec):
                a = a.reshape(len(vec), -1) if a.size != len(vec) else a.reshape(len(vec))
            cast.append(a)
        res = self._np(*cast)
        outs = []
        for r, shp in zip(res, self.out_shapes):
            if len(r) == 0:
                outs.append(np.zeros(shp))
                continue
            r = np.array(np.broadcast_arrays(*[np.asarray(v) for v in r]))
            outs.append(r.reshape(shp + r.shape[1:]))
        return outs[0] if len(outs) == 1 else tuple(outs)

    # -- CUDA / C emission
    def c_source(self, qualifiers: str = '__host__ __device__ __forceinline__', real: str = 'double',
                 scalar_inputs: Sequence[str] = (), name: Optional[str] = None, call_cache: Optional[dict] = None) -> str:
        """Emit ``void name(const real* in0, ..., real* out0, ...)``; 2-D outputs are stored row-major. Inputs listed
        in ``scalar_inputs`` (length-1 vectors) are passed by value.  With ``call_cache`` (see ``_emit``) the function
        gets a last argument ``cc``: the call cache it fills (``mode='store'``) or reads (``mode='load'``)."""
        in_map: Dict[int, str] = {}
        sig = []
        for iname, vec in zip(self.in_names, self.inputs):
            if iname in scalar_inputs:
                if len(vec) != 1:
                    raise ValueError(f"Scalar input '{iname}' must have exactly one element")
                in_map[vec[0].id] = iname
                sig.append(f'const {real} {iname}')
            else:
                for k, s in enumerate(vec):
                    in_map[s.id] = f'{iname}[{k}]'
                sig.append(f'const {real}* __restrict__ {iname}')
        flat = [o for vec in self.outputs for o in vec]
        lines, outs = _emit(flat, in_map, 'c', call_cache=call_cache)
        sig += [f'{real}* __restrict__ {n}' for n in self.out_names]
        if call_cache is not None:
            sig.append(f'{real}* __restrict__ cc' if call_cache['mode'] == 'store' else f'const {real}* __restrict__ cc')
        src = [f"{qualifiers} void {name or self.name}({', '.join(sig)})", '{']
        src += ['    ' + ln for ln in lines]
        pos = 0
        for oname, vec in zip(self.out_names, self.outputs):
            for k in range(len(vec)):
                src.append(f'    {oname}[{k}] = {outs[pos + k]};')
            pos += len(vec)
        src.append('}')
        return '\n'.join(src)

    def n_nodes(self) -> int:
        return len(topo_sort([o for vec in self.outputs for o in vec]))


def gp_tables_used(exprs) -> List[str]:
    """Ids of the GP / MLP tables referenced by ``exprs`` (in first-use order; MLP ids carry the prefix 'mlp:')."""
    ids: List[str] = []
    for node in topo_sort(as_vector(exprs)):
        if node.op.startswith('gp:') or node.op.startswith('mlp:'):
            gid = _gp_parts(node.op)[0]
            gid = ('mlp:' + gid) if node.op.startswith('mlp:') else gid
            if gid not in ids:
                ids.append(gid)
    return ids


def gp_tables_c_source(ids: Sequence[str]) -> str:
    """Definitions of the training tables for a generated module (device constants + host copies for the CPU logic build)."""
    out = []
    for gid in ids:
        if gid.startswith('mlp:'):
            t = MLP_TABLES[gid[4:]]
            rows = np.concatenate([np.concatenate([t['W1'], t['b1'][:, None]], axis=1).reshape(-1), t['W2'].reshape(-1),
                                   t['b2']])
            body = ', '.join(_c_const(float(v)) for v in rows)
            out.append(f'HB_GP_TABLE double hb_mlp_tab_{gid[4:]}[{rows.size}] = {{{body}}};')
            continue
        t = GP_TABLES[gid]
        rows = np.concatenate([t['Xs'], t['c'][:, None]], axis=1).reshape(-1)
        body = ', '.join(_c_const(float(v)) for v in rows)
        ils = ', '.join(_c_const(float(v)) for v in t['inv_ls'])
        out.append(f'HB_GP_TABLE double hb_gp_tab_{gid}[{rows.size}] = {{{body}}};')
        out.append(f'HB_GP_TABLE double hb_gp_ils_{gid}[{t["d"]}] = {{{ils}}};')
    return '\n'.join(out)


def evaluate(exprs, symbols, values):
    """Numerically evaluate ``exprs`` with ``symbols := values`` (convenience for host-side setup code)."""
    f = Function('f', [('v', as_vector(symbols))], [('o', exprs)])
    return f(np.asarray(values, dtype=float).reshape(-1))


def dot(a, b):
    return np.dot(np.asarray(a, dtype=object), np.asarray(b, dtype=object))


def quad_form(r, W):
    """``r^T W r`` for an object vector ``r`` and numeric/square ``W``."""
    r = as_vector(r)
    W = np.asarray(W)
    out = Expr.const(0.)
    for i in range(len(r)):
        for j in range(len(r)):
            w = W[i, j]
            if isinstance(w, Expr) or w != 0.:
                out = out + r[i] * w * r[j]
    return out
