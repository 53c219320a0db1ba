"""Symbolic functions of the space variables (reference: hermipy/function.py:23-212).

A `Function` is a SymPy expression in the canonical variables x0..x4 together with the sorted list
`dirs` of the directions it is defined on.  The letters x, y, z, w, v are accepted as aliases of
x0..x4 on input.  Host-only glue: `ccode()` produces the C expression in v[i] that `core.discretize`
compiles with NVRTC and evaluates on the GPU grid; `split()` finds the separable structure that
`Quad.tensorize_at` exploits to replace one d-dimensional transform by lower-dimensional ones.
"""
import re

import sympy


def _as_terms(expr, kind):
    return list(expr.args) if isinstance(expr, kind) else [expr]


class Function():

    # letter notation and subscript notation of the same five variables
    xyz = list(sympy.symbols('x y z w v', real=True))
    x_sub = [sympy.symbols('x{}'.format(i), real=True) for i in range(len(xyz))]

    # name -> canonical symbol, for parsing strings and renaming foreign symbols
    conv = {}
    for _letter, _canonical in zip(xyz, x_sub):
        conv[str(_letter)] = _canonical
        conv[str(_canonical)] = _canonical
    del _letter, _canonical

    @staticmethod
    def tensorize(args):
        """Product of functions living on disjoint sets of directions."""
        seen, product = set(), sympy.Integer(1)
        for a in args:
            if not isinstance(a, Function) or seen & set(a.dirs):
                raise ValueError("Invalid arguments!")
            seen |= set(a.dirs)
            product = product * a.sym
        return Function(product, dirs=sorted(seen))

    def __init__(self, expr, dirs=None, dim=None):
        if isinstance(expr, Function):
            self.sym, self.dirs = expr.sym, expr.dirs
            return
        if not isinstance(expr, sympy.Basic):
            expr = sympy.sympify(expr, locals=self.conv)
        # symbols called x, y, x0, ... (whatever their assumptions) become the canonical ones
        renames = {s: self.conv[str(s)] for s in expr.free_symbols if str(s) in self.conv}
        if renames:
            expr = expr.subs(renames)
        self.sym = expr.expand()
        used = expr.free_symbols & set(self.x_sub)
        if dirs is not None:
            if dim is not None or list(dirs) != sorted(dirs):
                raise ValueError("Invalid arguments!")
            self.dirs = dirs
        elif dim is not None:
            self.dirs = list(range(dim))
        else:
            # the smallest leading block of directions containing every variable used
            top = max((self.x_sub.index(s) for s in used), default=-1)
            self.dirs = list(range(top + 1))
        assert used <= {self.x_sub[d] for d in self.dirs}

    def __eq__(self, other):
        return self.dirs == other.dirs and self.sym == other.sym

    def __repr__(self):
        return str([self.x_sub[d] for d in self.dirs]) + " --> " + str(self.sym)

    __str__ = __repr__

    def _binary(self, other, op):
        if not isinstance(other, Function):
            other = Function(other, dirs=self.dirs)
        if self.dirs != other.dirs:
            raise ValueError("Invalid argument!")
        return Function(op(self.sym, other.sym), dirs=self.dirs)

    def __mul__(self, other):
        return self._binary(other, lambda a, b: a * b)

    def __truediv__(self, other):
        return self._binary(other, lambda a, b: a / b)

    def __add__(self, other):
        return self._binary(other, lambda a, b: a + b)

    __rmul__ = __mul__
    __radd__ = __add__

    def as_xyz(self):
        return self.sym.subs(dict(zip(self.x_sub, self.xyz)))

    def ccode(self):
        """C expression in v[0..len(dirs)-1] (v[i] = coordinate along dirs[i])."""
        text = sympy.ccode(self.sym)
        for i, d in enumerate(self.dirs):
            text = re.sub(r'\bx{}\b'.format(d), 'v[{}]'.format(i), text)
        return text

    def split(self, legacy=True):
        """Separable structure of every additive term.

        legacy=False: one dict per term, {frozenset(directions): Function}; the key frozenset() holds the
        constant multiplier, every direction of `dirs` appears in exactly one key, and factors sharing a
        variable are merged into one group.  legacy=True: one list per term,
        [Function per direction..., multiplier] if the term is fully separable, else
        [Function(term/multiplier), multiplier]."""
        out = []
        for term in _as_terms(self.sym, sympy.Add):
            groups = {frozenset({d}): sympy.Integer(1) for d in self.dirs}
            groups[frozenset()] = sympy.Integer(1)
            for factor in _as_terms(term, sympy.Mul):
                key = frozenset(self.x_sub.index(s) for s in factor.free_symbols & set(self.x_sub))
                if key not in groups:
                    # merge every existing group that shares a direction with this factor
                    for other in [k for k in groups if k & key]:
                        factor = factor * groups.pop(other)
                        key = key | other
                    groups[key] = sympy.Integer(1)
                groups[key] = groups[key] * factor
            if legacy:
                constant = groups[frozenset()]
                if len(groups) == len(self.dirs) + 1:
                    parts = [Function(groups[frozenset({d})], dirs=[d]) for d in self.dirs]
                else:
                    parts = [Function(term / constant, dirs=self.dirs)]
                out.append(parts + [constant])
            else:
                out.append({key: Function(value, dirs=sorted(key)) for key, value in groups.items()})
        return out

    def project(self, dirs):
        """Restriction of a product-form function to `dirs`; the constant multiplier goes with the lowest
        direction of self.dirs."""
        dirs = dirs if isinstance(dirs, list) else [dirs]
        pieces = self.split(legacy=False)
        if len(pieces) != 1:
            raise ValueError("Invalid argument: function can't be split")
        groups = pieces[0]
        result = groups[frozenset()].sym if self.dirs[0] in dirs else sympy.Integer(1)
        for key, part in groups.items():
            if not key:
                continue
            if key <= set(dirs):
                result = result * part.sym
            elif dirs and set(dirs) <= key:
                raise ValueError("Can't project factor!")
        return Function(result, dirs=dirs)

    @staticmethod
    def sanitize(expr, max_denom=1e8):
        """Replace floating-point coefficients by nearby rationals (|c| < 1e-14 -> 0)."""
        total = sympy.Integer(0)
        for term in _as_terms((sympy.Integer(0) + expr).expand(), sympy.Add):
            product = sympy.Integer(1)
            for factor in _as_terms(term, sympy.Mul):
                if isinstance(factor, sympy.Float):
                    factor = 0 if abs(factor) < 1e-14 else factor
                    factor = sympy.Rational(factor).limit_denominator(max_denom)
                product = product * factor
            total = total + product
        return total
