"""Material parameter sets for the semi-affine fiber model.

Same call surface and values as ``/root/reference/jointforces/materials.py:1-41``: every material
is a plain dict ``{'K_0', 'D_0', 'L_S', 'D_S'}`` (linear stiffness, buckling coefficient, onset
strain of stiffening, stiffening coefficient; Steinwachs et al. 2016).  1e30 switches a regime
off; the kernels evaluate w, w', w'' without cancellation in that limit.
"""

_OFF = 1e30
_KEYS = ('K_0', 'D_0', 'L_S', 'D_S')


def custom(K_0, D_0, L_S, D_S):
    """Arbitrary nonlinear material (reference ``materials.py:11``)."""
    return dict(zip(_KEYS, (K_0, D_0, L_S, D_S)))


def linear(stiffness):
    """Linear elastic fibers of stiffness K_0 (reference ``materials.py:1``); E = K_0/6 at nu = 0.25."""
    return custom(stiffness, _OFF, _OFF, _OFF)


# Collagen G/R 1:1 at 0.6 / 1.2 / 2.4 mg/ml, fibrin 4.0 mg/ml, matrigel 10 mg/ml
# (reference ``materials.py:28-41``)
collagen06 = custom(447, 0.0008, 0.0075, 0.033)
collagen12 = custom(1645, 0.0008, 0.0075, 0.033)
collagen24 = custom(5208, 0.0008, 0.0075, 0.033)
fibrin40 = custom(2091, 0.002, _OFF, _OFF)
matrigel100 = custom(2364, _OFF, _OFF, _OFF)
