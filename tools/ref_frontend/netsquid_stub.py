"""Kept for the fixture tooling's import path: the NetSquid name shim lives in the product package now
(``dqc_simulator_b200/compat/netsquid_shim.py``) so that ``compat.frontend.run_qasm`` can use it."""
from dqc_simulator_b200.compat.netsquid_shim import *  # noqa: F401,F403
from dqc_simulator_b200.compat.netsquid_shim import install  # noqa: F401
