"""Linear differential operators acting on an unknown f(x0, ...) (reference: hermipy/operator.py:26-146).

`split()` writes the operator as sum_m c_m(x) d^m f and returns {m: Function c_m}; `Quad.discretize_op`
turns every term into a varf (coefficient) followed by varfd column shifts (derivatives).  `map(factor)`
conjugates the operator, L -> factor^{-1} L(factor .), which is how a weighted basis enters."""
import itertools
import math

import sympy

from . import function as func

_Unknown = sympy.core.function.AppliedUndef


class Operator():

    f = sympy.Function('f')

    def __init__(self, expr, dirs=None):
        if isinstance(expr, Operator):
            self.sym, self.dirs = expr.sym, expr.dirs
            return
        parsed = func.Function(expr).sym
        if parsed == 0:
            if dirs is None:
                raise ValueError("Invalid argument!")
            self.sym, self.dirs = parsed, dirs
            return
        unknowns = list(parsed.atoms(_Unknown))
        if len(unknowns) != 1:
            raise TypeError("There should be exactly 1 functional argument!")
        self.sym = parsed.subs(unknowns[0].func, self.f)
        variables = list(self.sym.atoms(_Unknown))[0].args
        self.dirs = [func.Function.x_sub.index(v) for v in variables]
        assert len(self.dirs) == len(unknowns[0].args)

    def _variables(self):
        return [func.Function.x_sub[d] for d in self.dirs]

    def __eq__(self, other):
        return self.dirs == other.dirs and self.sym == other.sym

    def __repr__(self):
        return str(self.f(*self._variables())) + " --> " + str(self.sym)

    __str__ = __repr__

    def __mul__(self, other):
        if not isinstance(other, func.Function):
            other = func.Function(other, dirs=self.dirs)
        if self.dirs != other.dirs:
            raise ValueError("Invalid argument!")
        return Operator(self.sym * other.sym, dirs=self.dirs)

    def _combine(self, other, sign):
        if not isinstance(other, Operator):
            other = Operator(other, dirs=self.dirs)
        if self.dirs != other.dirs:
            raise ValueError("Invalid argument!")
        return Operator(self.sym + sign * other.sym, dirs=self.dirs)

    def __add__(self, other):
        return self._combine(other, 1)

    def __sub__(self, other):
        return self._combine(other, -1)

    def __neg__(self):
        return Operator(-self.sym, dirs=self.dirs)

    __radd__ = __add__
    __rmul__ = __mul__

    def __call__(self, arg):
        """Apply to a Function (-> Function) or compose with an Operator (-> Operator)."""
        if not isinstance(arg, (func.Function, Operator)):
            try:
                arg = Operator(arg, dirs=self.dirs)
            except TypeError:
                arg = func.Function(arg, dirs=self.dirs)
        applied = self.sym.subs(self.f(*self._variables()), arg.sym).doit()
        kind = Operator if isinstance(arg, Operator) else func.Function
        return kind(applied, dirs=self.dirs)

    def as_xyz(self):
        return self.sym.subs(dict(zip(func.Function.x_sub, func.Function.xyz)))

    def map(self, factor):
        if not isinstance(factor, func.Function):
            factor = func.Function(factor, dirs=self.dirs)
        if factor.dirs != self.dirs:
            raise ValueError("Directions differ")
        unknown = self.f(*self._variables())
        mapped = self.sym.subs(unknown, unknown * factor.sym).doit() / factor.sym
        return Operator(mapped.expand(), dirs=self.dirs)

    def split(self, max_order=5):
        """{derivative multi-order m: coefficient Function}.  The coefficient of d^m f is read off by
        applying the operator to the monomial x^m/m!, lowest orders first, so that lower-order terms are
        already removed from the remainder when a monomial is tried."""
        variables = self._variables()
        unknown = self.f(*variables)
        result, remainder = {}, self.sym.expand()
        orders = [m for m in itertools.product(range(max_order + 1), repeat=len(variables)) if sum(m) <= max_order]
        for m in orders:
            if remainder == 0:
                break
            monomial = sympy.Integer(1)
            for power, v in zip(m, variables):
                monomial = monomial * v ** power / math.factorial(power)
            coefficient, rest = sympy.Integer(0), sympy.Integer(0)
            for piece in (remainder.args if isinstance(remainder, sympy.Add) else [remainder]):
                value = piece.subs(unknown, monomial).doit()   # term by term: no cancellation noise
                if value == 0:
                    rest = rest + piece
                else:
                    coefficient = coefficient + value
            remainder = rest
            coefficient = sympy.simplify(coefficient)
            if coefficient != 0:
                result[tuple(m)] = func.Function(coefficient, dirs=self.dirs)
        if remainder != 0:
            raise ValueError("Nonzero remainder")
        return result
