"""
TEST DOUBLE of the few `phi` 3.4 names that phiflow_b200/phi_cuda/flow.py touches - NOT PhiFlow.

PhiFlow 3.4 cannot be imported in this image (it needs phiml >= 1.14: `phiml.dataclasses`, top-level `from phiml import Tensor`;
the reference vendors PhiML 1.7.2 only), so the facade `phi_cuda/flow.py` could never execute.  This package gives it something to
execute against: Fields that hold REAL phiml Tensors / Extrapolations / Shapes under the attribute names of phi/field/_field.py:51-474
(`values`, `extrapolation`, `bounds`, `resolution`, `dx`, `is_grid`, `is_staggered`, `geometry`, `shape`, `with_values`,
`with_extrapolation`), the two private helpers of phi/physics/fluid.py the facade calls (`_get_obstacles_for` :85-91,
`_pressure_extrapolation` :264-274), and "stock" functions that only record that they were called (STOCK_CALLS) so the tests can see
a fall-through.  It is on sys.path only inside tests/test_phi_cuda_facade.py.
"""
__version__ = '0.0-test-double'
STOCK_CALLS = []          # (module.function, args) appended by every stock function of this double
