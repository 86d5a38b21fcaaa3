"""Traction constants, API of reference `src/phasefieldx/Loading/loading_functions.py:21-111`."""
from phasefieldx_b200 import fem


def loading_Tx(msh, value=0.0):
    return fem.Constant(msh, (value))


def loading_Txy(msh, value_x=0.0, value_y=0.0):
    return fem.Constant(msh, (value_x, value_y))


def loading_Txyz(msh, value_x=0.0, value_y=0.0, value_z=0.0):
    return fem.Constant(msh, (value_x, value_y, value_z))
