"""Optional matplotlib front-ends of Quad.plot / Series.plot / Varf.plot (reference: quad.py:427-533,
series.py:213-257, varf.py:239-284).  matplotlib is imported lazily: the numerical package works
without it (the reference imports it at module top, series.py:27, varf.py:31)."""
import numpy as np


def _axes(ax):
    try:
        import matplotlib.pyplot as plt
    except ImportError as e:   # pragma: no cover
        raise ImportError("plotting needs matplotlib, which is not installed") from e
    if ax is None:
        return plt, plt.subplots(1)[1], True
    return plt, ax, False


def _finish(plt, ax, artist, own, colorbar):
    if not own:
        return artist
    if colorbar:
        plt.colorbar(artist, ax=ax)
    plt.show()


def _support_lines(ax, position, degrees):
    """+- sqrt(2) sqrt(2 k + 1) sigma around the mean: where the Hermite function of degree k lives."""
    for i, k in enumerate(degrees[:2]):
        half = np.sqrt(2) * np.sqrt(2 * k + 1) * np.sqrt(position.cov[i][i])
        line = ax.axvline if i == 0 else ax.axhline
        line(position.mean[i] - half)
        line(position.mean[i] + half)


def plot_quad(quad, arg, factor=None, ax=None, contours=0, bounds=False, title=None, **kwargs):
    import sympy as sym
    from .function import Function
    from .series import Series
    if not quad.position.is_diag:
        raise ValueError("Invalid argument: position must be diag!")
    plt, ax, own = _axes(ax)
    if isinstance(arg, sym.Basic):
        arg = Function(arg, dirs=quad.position.dirs)
    if isinstance(arg, Function):
        if factor is not None:
            raise ValueError("Invalid argument!")
        values = quad.discretize(arg)
    elif isinstance(arg, np.ndarray):
        values = arg
    elif isinstance(arg, Series):
        if np.iscomplexobj(arg.coeffs):
            arg.coeffs = np.real(arg.coeffs).copy(order='C')
        values = quad.eval(arg)
        if factor is not None:
            values = values * (factor if isinstance(factor, np.ndarray) else quad.discretize(factor))
        if bounds:
            _support_lines(ax, arg.position, [arg.degree] * arg.position.dim)
    else:
        raise TypeError("Unsupported type: " + str(type(arg)))
    axes_nodes = [quad.project(d).discretize(Function.xyz[d]) for d in quad.position.dirs]
    values = values.reshape(*[len(n) for n in quad.nodes]).T
    if quad.position.dim == 1:
        artist = ax.plot(*axes_nodes, values, **kwargs)
    else:
        artist = ax.contourf(*axes_nodes, values, 100, **kwargs)
        if contours > 0:
            ax.contour(*axes_nodes, values, levels=contours, colors='k')
    ax.set_title(title if title is not None else "Min: {:.3f}, Max: {:.3f}".format(np.min(values), np.max(values)))
    return _finish(plt, ax, artist, own, quad.position.dim == 2)


def plot_hermite_function(quad, multi_index, ax=None, bounds=True, **kwargs):
    from . import core
    if len(multi_index) != quad.position.dim:
        raise ValueError("Invalid argument!")
    plt, ax, own = _axes(ax)
    coeffs = np.zeros(core.iterator_size(len(multi_index), sum(multi_index), index_set="triangle"))
    coeffs[core.iterator_index(multi_index, index_set="triangle")] = 1
    artist = plot_quad(quad, quad.series(coeffs, index_set="triangle"), ax=ax, bounds=False, **kwargs)
    if bounds:
        _support_lines(ax, quad.position, list(multi_index))
    return _finish(plt, ax, artist, own, quad.position.dim == 2)


def plot_series(series, ax=None, title=None):
    plt, ax, own = _axes(ax)
    m = series.multi_indices()
    artist = None
    if series.position.dim == 1:
        artist = ax.bar(m[:, 0], series.coeffs)
    elif series.position.dim == 2:
        from matplotlib.ticker import MaxNLocator
        artist = ax.scatter(m[:, 0], m[:, 1], c=abs(series.coeffs), cmap='ocean_r', s=1e2, edgecolor='k')
        ax.xaxis.set_major_locator(MaxNLocator(integer=True))
        ax.yaxis.set_major_locator(MaxNLocator(integer=True))
    ax.set_title("Coefficients of the Hermite expansion" if title is None else title)
    return _finish(plt, ax, artist, own, series.position.dim == 2)


def plot_varf(varf, ax=None, lines=True):
    import matplotlib.colors
    plt, ax, own = _axes(ax)
    coo = varf.matrix.tocoo() if varf.is_sparse else None
    if coo is not None:
        rows, cols, data = coo.row, coo.col, abs(coo.data)
    else:
        rows, cols = np.nonzero(abs(varf.matrix) > 1e-9)
        data = abs(varf.matrix[rows, cols])
    artist = ax.scatter(rows, cols, c=data, cmap='ocean_r', norm=matplotlib.colors.LogNorm())
    ax.set_xlim([0, varf.matrix.shape[0]])
    ax.set_ylim([0, varf.matrix.shape[1]])
    if lines:   # separators between degrees (meaningful for the graded sets: triangle by sum, cube by max)
        grade, degree = (sum if varf.index_set == 'triangle' else max), 0
        for i, m in enumerate(varf.multi_indices()):
            if grade(m) > degree:
                degree += 1
                ax.axvline(x=i, ymin=0, ymax=i)
                ax.axhline(y=i, xmin=0, xmax=i)
    ax.invert_yaxis()
    return _finish(plt, ax, artist, own, True)
