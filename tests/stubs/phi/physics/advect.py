"""phi.physics.advect of the test double (signatures: phi/physics/advect.py:20-24, 156-159, 182-186)."""
from .. import STOCK_CALLS


def euler(*args, **kwargs):
    raise NotImplementedError("integrator object of the test double; only its identity is used")


def rk4(*args, **kwargs):
    raise NotImplementedError


def semi_lagrangian(field, velocity, dt, integrator=euler):
    STOCK_CALLS.append(('advect.semi_lagrangian', (field, velocity, dt, integrator)))
    return 'stock semi_lagrangian'


def mac_cormack(field, velocity, dt, correction_strength=1.0, integrator=euler):
    STOCK_CALLS.append(('advect.mac_cormack', (field, velocity, dt, correction_strength, integrator)))
    return 'stock mac_cormack'


def advect(field, velocity, dt, integrator=euler):
    STOCK_CALLS.append(('advect.advect', (field, velocity, dt, integrator)))
    return 'stock advect'
