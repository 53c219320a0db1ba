"""Library-wide options (reference: hermipy/settings.py:19-26, documented in docs/source/settings.rst).
Every option can be overridden per call through the keyword of the same name (`sparse=`, `tensorize=`,
`cache=`)."""

settings = {
    'cache': False,        # memoise core.* results as .npy/.npz files under `cachedir`
    'cachedir': '.cache',
    'debug': False,        # print name and arguments of every decorated call
    'sparse': False,       # varf / tensorize return scipy.sparse.csr_matrix
    'tensorize': True,     # split separable functions into lower-dimensional transforms (Quad.tensorize_at)
    'trails': False,       # print an indented call trail with timings
}
