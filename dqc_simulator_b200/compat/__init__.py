"""Glue between the reference package and the engine:

* ``netsquid_shim`` -- the NetSquid *names* the reference's front end needs to import (no arithmetic);
* ``frontend``      -- ``run_qasm``: QASM file + hardware spec -> reference front end (unmodified, in-process)
                       -> executor -> CUDA engine;
* ``qrepr``         -- the level-1 boundary of SURVEY.md 8b: a NetSquid ``QRepr``-shaped state object whose
                       ``operate / measure / apply_pauli_noise / depolarize / reduced_dm / .dm`` land in the C-ABI.
"""
